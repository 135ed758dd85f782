"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: slab layout, halo exchange pattern,
norm all-reduce and the stopping state machine, with the CUDA pass replaced by the schedule emulator."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from fluid_dynamics1_b200.parallel import slab_bounds, slab_layout


def _run(world, rows, cols, T, tol, itmax, port, extra_env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist", "slab_cpu_worker.py"), str(rows), str(cols), str(T),
           str(tol), str(itmax)]
    env = dict(os.environ, OMP_NUM_THREADS="1", **(extra_env or {}))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok=True" in r.stdout


def test_slab_bounds_partition():
    for rows in (17, 64, 4096, 16384):
        for world in (1, 2, 3, 4, 8):
            b = [slab_bounds(rows, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == rows
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(y - x for x, y in b) - min(y - x for x, y in b) <= 1
    g0, n, lo, hi, hlo, hhi = slab_layout(4096, 8, 3, 4)
    assert (g0, n, lo, hi, hlo, hhi) == (1536 - 8, 512 + 16, 8, 520, 8, 8)
    assert slab_layout(4096, 8, 0, 4)[:4] == (0, 512 + 8, 0, 512)
    assert slab_layout(4096, 8, 7, 4)[:4] == (3584 - 8, 512 + 8, 8, 520)


def test_slab_protocol_gloo_trimmed_boundary_chunks():
    """Three slabs whose boundary chunks are shorter than the interior ones (PassGeom::trim_lo/hi; the middle rank is
    trimmed at both ends): same field and sweep count as the single-domain oracle."""
    _run(3, 180, 40, 2, 1e-3, 5000, 29690, dict(CNV_POISSON_TRIM="4", CNV_TEST_CHUNKS="4"))


@pytest.mark.parametrize("world,T,tol,itmax", [(2, 4, 1e-3, 5000), (2, 2, 0.0, 9), (3, 1, 1e-2, 5000)])
def test_slab_protocol_gloo(world, T, tol, itmax):
    _run(world, 60, 40, T, tol, itmax, 29600 + world * 10 + T)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_c_driver_bootstrap_layer(world):
    """The fork + socket-pair bootstrap of `CNV_GPUS=N ./cnavier_b200` (csrc/driver.cc Boot: all-gather of the 256-byte IPC
    records, broadcast of the NCCL id, barrier, agreement on a failure) with `world` real processes and no GPU.  Run in a
    child interpreter: the self-test forks, which a pytest process full of threads should not do."""
    import subprocess
    import sys

    code = f"import fluid_dynamics1_b200 as fd, sys; sys.exit(fd.lib().cnv_boot_selftest({world}))"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
