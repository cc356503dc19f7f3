"""Runs tests/multigpu_check.py under torchrun on 2 GPUs when the box has them (skipped on a 1-GPU box; the
host-side protocol is covered on CPU by tests/test_distributed_cpu.py)."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _gpus() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpus() < 2, reason="needs 2 GPUs")
def test_multi_gpu_modes_a_and_b():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTIGPU OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
