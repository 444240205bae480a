"""Multi-GPU checks (one process per GPU under torchrun, NCCL): collected by pytest, skipped on
boxes with fewer than two GPUs.  Each script prints "... PASS" from rank 0."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def _torchrun(script, args, nproc=2, timeout=900):
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, script)] + args
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    return out.returncode, out.stdout + out.stderr


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("script,args,token", [
    ("mgpu_sharded_check.py", ["16"], "PASS"),
    ("mgpu_hierarchy_check.py", ["32", "3"], "MGPU_HIERARCHY PASS"),
    ("mgpu_solve_check.py", ["32", "3"], "PASS"),
])
def test_two_gpu_checks(script, args, token):
    rc, out = _torchrun(script, args)
    assert rc == 0, out[-3000:]
    assert token in out and "FAIL" not in out, out[-3000:]
