"""Multi-GPU parity (needs >= 2 GPUs; skipped on the single-GPU box): row-sharded RHS and integrator vs the
single-GPU result, through torchrun + scripts/multi_gpu_check.py."""
import ctypes
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu(oq):
    n = ctypes.c_int(0)
    oq._lib.load().oq_device_count(ctypes.byref(n))
    return n.value


@pytest.mark.parametrize("form", ["dense", "classes"])
@pytest.mark.parametrize("world", [2])
def test_sharded_equals_single(gpu, world, form):
    if _ngpu(gpu) < world:
        pytest.skip(f"needs {world} GPUs")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", "29561" if form == "dense" else "29562",
                          os.path.join(ROOT, "scripts", "multi_gpu_check.py"), form],
                         capture_output=True, text=True, timeout=400, cwd=ROOT)
    assert res.returncode == 0, (res.stdout[-1500:], res.stderr[-1500:])
    assert res.stdout.count("ok=True") == 2 * world and "ok=False" not in res.stdout   # Tsit5 and VCABM5 per rank
