#!/bin/bash
# ncu --set full capture of one kernel.  usage: gpu_ncu.sh <tag> <kernel-regex> <skip> <count> <cmd...>
tag=$1; kern=$2; skip=$3; cnt=$4; shift 4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c $cnt -f -o gpurun_out/$tag "$@" > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
ls -la gpurun_out/$tag.ncu-rep
