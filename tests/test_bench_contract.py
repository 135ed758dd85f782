"""Host-side logic of bench.py (no GPU): the workload each N measures, the synthetic input every rank can regenerate for any
row range, the CPU arm's thread environment, and the behaviour of the reference arm's non-zero ranks under torchrun."""
import argparse
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    d = dict(gpus=1, steps=20, warmup=3, impl="b200", series="auto", n=4096, sweeps=1024, T=0, scaling="weak", no_cpu=False,
             cpu_seconds=15.0)
    d.update(kw)
    return argparse.Namespace(**d)


def test_headline_workloads_are_the_baseline_configs():
    """N = 1: BASELINE config 4 (4096^2).  N > 1: the weak series on config 5's per-GPU shape -- exactly 16384^2 at 8 GPUs."""
    rows, cols, per, scaling, desc = bench.headline_shape(_args(), 1)
    assert (rows, cols, per, scaling) == (4096, 4096, 4096, "weak") and "config 4" in desc
    for world in (2, 4, 8):
        rows, cols, per, scaling, desc = bench.headline_shape(_args(gpus=world), world)
        assert (rows, cols, per, scaling) == (2048 * world, 16384, 2048, "weak") and "config 5" in desc
    assert bench.headline_shape(_args(gpus=8), 8)[:2] == (16384, 16384)
    # --series headline measures the same workload without the side figures
    assert bench.headline_shape(_args(gpus=8, series="headline"), 8) == bench.headline_shape(_args(gpus=8), 8)
    # --series single: the grid given on the command line, weak or strong
    assert bench.headline_shape(_args(gpus=4, series="single", n=4096, scaling="strong"), 4)[:4] == (4096, 4096, 1024, "strong")
    assert bench.headline_shape(_args(gpus=4, series="single", n=4096, scaling="weak"), 4)[:4] == (16384, 4096, 4096, "weak")


def test_both_arms_print_the_same_config_object():
    """The driver divides the two arms' values: they must describe the same workload (identical configs)."""
    for world in (1, 2, 8):
        a = bench.workload_config(_args(gpus=world, impl="b200"), world)
        b = bench.workload_config(_args(gpus=world, impl="reference"), world)
        assert a == b and a["gpus"] == world and a["sweeps_per_step"] == 1024 and "workload" in a and "l2" in a
    assert "larger than L2" in bench.workload_config(_args(), 1)["l2"]


def test_synthetic_input_is_the_same_for_every_decomposition():
    """A slab (plus any neighbouring rows) regenerates exactly the rows of the whole field: the parity check of an N-GPU run
    compares against an oracle run on those rows."""
    total, cols = 1000, 96
    whole = bench.synthetic_vorticity(total, cols)
    assert whole.shape == (total, cols) and whole.flags.c_contiguous and np.all(np.isfinite(whole))
    for row0, n in ((0, 250), (250, 250), (240, 300), (700, 300), (255, 2), (511, 2)):
        part = bench.synthetic_vorticity(n, cols, row0=row0, total_rows=total)
        assert part.tobytes() == whole[row0:row0 + n].tobytes(), (row0, n)
    assert not np.array_equal(bench.synthetic_vorticity(64, cols, seed=1), bench.synthetic_vorticity(64, cols, seed=2))


def test_cpu_arm_uses_every_host_core(monkeypatch):
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must not inherit that (a 1-thread reference
    inflated the ratio 12 x)."""
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    monkeypatch.setenv("OMP_PLACES", "cores")
    env = bench.cpu_env()
    assert int(env["OMP_NUM_THREADS"]) == bench.host_cores() >= 1 and "OMP_PLACES" not in env


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without output and without work."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_roofline_traffic_only_for_captured_shapes():
    """`roofline.traffic` is the ncu figure of exactly the shape timed, else null (a constant was reported for
    shapes it was never measured on)."""
    assert bench.measured_traffic(4096, 4096, 8) and bench.measured_traffic(4096, 4096, 8) > 3e8
    assert bench.measured_traffic(2080, 16384, 8) is None and bench.measured_traffic(4096, 4096, 4) is None
