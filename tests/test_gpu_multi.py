"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a 1-GPU box): the slab-decomposed Poisson solve and the
slab-decomposed time stepping must be bit-identical to the single-GPU path and the oracle."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import fluid_dynamics1_b200 as fd
        return fd.lib().cnv_device_count()
    except Exception:
        return 0


def _torchrun(script, args, world, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist", script)] + [str(a) for a in args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_poisson_bitwise(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    assert "SLAB CHECK PASSED" in _torchrun("slab_gpu_check.py", [world * 40, 96, 4], world, 29700 + world)
    assert "SLAB CHECK PASSED" in _torchrun("slab_gpu_check.py", [1024, 1024, 8], world, 29710 + world)


@pytest.mark.parametrize("world", [2, 4])
def test_slab_time_stepping_bitwise(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    assert "SLAB SIM CHECK PASSED" in _torchrun("slab_sim_gpu_check.py", [128, 4], world, 29720 + world)
