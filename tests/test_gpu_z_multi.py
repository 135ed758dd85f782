"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a 1-GPU box): the slab-decomposed Poisson solve and the
slab-decomposed time stepping must be bit-identical to the single-GPU path and the oracle.
(The file name sorts after test_gpu_parity.py on purpose: under `pytest -x` the single-GPU parity suite runs first.)"""
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


BACKENDS = ["peer", "nccl", "torch"]  # CNV_DIST_BACKEND: in-kernel NVLink peer stores | library NCCL group | torch p2p


def _torchrun(script, args, world, port, backend="peer", extra_env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist", script)] + [str(a) for a in args]
    env = dict(os.environ, CNV_DIST_BACKEND=backend, **(extra_env or {}))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_poisson_bitwise(world, backend):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    assert "SLAB CHECK PASSED" in _torchrun("slab_gpu_check.py", [world * 40, 96, 4], world, 29700 + world, backend)
    assert "SLAB CHECK PASSED" in _torchrun("slab_gpu_check.py", [1024, 1024, 8], world, 29710 + world, backend)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("world", [2, 4])
def test_slab_time_stepping_bitwise(world, backend):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    assert "SLAB SIM CHECK PASSED" in _torchrun("slab_sim_gpu_check.py", [128, 4], world, 29720 + world, backend)


@pytest.mark.parametrize("world", [2, 4])
def test_slab_peer_repeated_solves(world):
    """Back-to-back solves on the peer-memory path: per-sweep residual history and field identical to the single-GPU
    solve every time (guards the epoch handshake between re-initialising the iterate and the neighbours' pushes)."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    assert "STRESS PASSED" in _torchrun("slab_stress.py", [world * 100, 96, 4, 12], world, 29730 + world, "peer")
    assert "STRESS PASSED" in _torchrun("slab_stress.py", [1024, 1024, 8, 4], world, 29740 + world, "peer")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_c_driver_multi_gpu_matches_single_gpu(world, tmp_path):
    """`CNV_GPUS=N cnavier_b200 cfg run` (csrc/driver.cc: one forked process per GPU, NCCL id and CUDA IPC handles exchanged in
    C over socket pairs, cnv_sim_step_slab) against the same executable on one GPU: identical Poisson and continuity log lines
    and byte-identical VTK files (config_high_re.txt grid, 128^2; both Poisson exchanges)."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import api
    import fluid_dynamics1_b200 as fd
    exe = os.path.join(os.path.dirname(fd._lib.LIB_PATH), "cnavier_b200")
    cfg = dict(api.CONFIG_HIGH_RE, tf=(6 + 0.5) * 0.001, output_interval=3)       # 6 steps, dumps at t = 0, 3

    def run(tag, env):
        d = tmp_path / tag
        d.mkdir()
        api.write_config(cfg, str(d / "cfg.txt"))
        r = subprocess.run([exe, "cfg.txt", "run"], cwd=d, capture_output=True, text=True, timeout=240, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
        log = (d / "output" / "logs" / "run.txt").read_text()
        lines = [ln.split(" | Elapsed")[0] for ln in log.splitlines() if ln.startswith(("Poisson equation", "Continuity max"))]
        vtk = {n: (d / "output" / "run" / n).read_bytes() for n in sorted(os.listdir(d / "output" / "run"))}
        return r.stdout, lines, vtk

    _, lines1, vtk1 = run("one", {})
    assert len(lines1) == 12 and len(vtk1) == 8
    for backend in ("peer", "nccl"):
        out, lines, vtk = run(f"{backend}{world}", dict(CNV_GPUS=str(world), CNV_DIST_BACKEND=backend, CNV_PEER_TIMEOUT_MS="30000"))
        assert f"Slab decomposition: {world} ranks" in out and ("peer memory" in out) == (backend == "peer"), out[-1500:]
        assert lines == lines1, backend
        assert vtk == vtk1, backend
