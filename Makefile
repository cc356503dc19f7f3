# Build the CUDA engine (sm_100a only) and the CPU oracle.
NVCC      ?= /usr/local/cuda/bin/nvcc
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall -Xptxas -v
CSRC      := bayadera_b200/csrc
LIB       := bayadera_b200/libbayadera_b200.so

all: $(LIB) oracle

$(LIB): $(wildcard $(CSRC)/*.cu $(CSRC)/*.cuh $(CSRC)/*.inc) include/bayadera_b200.h
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC)/engine.cu -ldl 2> $(CSRC)/ptxas.log || (cat $(CSRC)/ptxas.log; exit 1)

oracle:
	$(MAKE) -C oracle
	$(MAKE) -C oracle ref

clean:
	rm -f $(LIB) $(CSRC)/ptxas.log
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
