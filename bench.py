#!/usr/bin/env python
"""bench.py -- headline benchmark of the cnavier hot path on B200.

Metric (BASELINE.json): Poisson cell-updates/s (and sweeps/s) against the HBM roofline.
Workload at N GPUs: the 4096^2 Re=1000 lid-driven cavity grid (BASELINE config 4, the configuration
the north_star target is quoted on; it fits one GPU), slab-decomposed over N GPUs by rows with a
fixed 4096 x 4096 slab per GPU (weak scaling; `--scaling strong` keeps the total at 4096^2).
A "step" = one fixed-sweep red-black SOR solve (`--sweeps`, default 1024) of lap(psi) = -w from a zero
initial guess on a synthetic cavity-like vorticity field (zero guess + state reset are inside the
timed region; no convergence exit: tol = 0, so no work is ever skipped).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                     # the reference's own CPU code (oracle/_ref)

One JSON line on stdout (rank 0).  See DESIGN.md section "Measurement" for the definitions.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALGO_BYTES_PER_CELL_SWEEP = 24.0  # read psi + read f + write psi


def synthetic_vorticity(nrows, ncols, row0=0, total_rows=None, seed=1234):
    """Deterministic cavity-like vorticity: strong near the moving lid (last row), smooth bulk, plus
    small-scale noise so that no cell is trivially zero."""
    total_rows = total_rows or nrows
    rng = np.random.default_rng(seed + row0)
    i = (np.arange(row0, row0 + nrows) / total_rows)[:, None]
    j = (np.arange(ncols) / ncols)[None, :]
    w = -40.0 * np.exp(-((1 - i) * 30.0)) * np.sin(np.pi * j) + 4.0 * np.sin(2 * np.pi * i) * np.sin(3 * np.pi * j)
    w += 0.05 * rng.standard_normal((nrows, ncols))
    return np.ascontiguousarray(w)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  The sampler is
    started before the warm-up (nvidia-smi needs a moment to produce its first line); samples are selected by
    their timestamps against the wall-clock window of the timed region."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.proc = None
        self.device = device
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for stamp, ln in self.lines:
            if t0 is not None and not (t0 - 0.03 <= stamp <= t1 + 0.08):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 10:
                continue
            try:
                sm.append(float(f[2])); mx.append(float(f[3])); power.append(float(f[4]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(power)) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def _config_name(n):
    return {1024: "BASELINE config 3", 4096: "BASELINE config 4", 16384: "BASELINE config 5"}.get(n, "custom")


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import fluid_dynamics1_b200 as fd
    from fluid_dynamics1_b200 import parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    fd.require_gpu()
    torch.cuda.set_device(local_rank)
    L = fd.lib()
    L.cnv_set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ncols = args.n
    if args.scaling == "weak":
        total_rows, rows_per = args.n * world, args.n
    else:
        total_rows, rows_per = args.n, args.n // world
    T, S = args.T, args.sweeps
    dx = dy = 1.0 / ncols
    beta = fd.sor_beta(ncols, ncols)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    slab = parallel.SlabPoisson(total_rows, ncols, T, rank, world, stream=sp) if world > 1 else None
    if slab is None:
        solver = fd.PoissonSolver(rows_per, ncols, T)
        solver.set_consts(dx, dy, beta)
        w_host = synthetic_vorticity(rows_per, ncols)
        solver.upload(w_host, -1.0, sp)  # f = -w (src/main.c:348)
        T = solver.T
        plan = solver.plan
    else:
        slab.set_consts(dx, dy, beta)
        w_host = synthetic_vorticity(slab.own_rows, ncols, slab.row0, total_rows)
        slab.upload_owned(w_host, -1.0)
        T = slab.T
        plan = slab.solver.plan
    npass = (S + T - 1) // T
    interior_cells = (total_rows - 2) * (ncols - 2)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step(pass_events=None):
        """resident-input step: zero guess, reset state machine, S sweeps."""
        if slab is None:
            solver.L.cnv_poisson_prepare(solver.h, None, 0, 1.0, sp)  # f == NULL: only zero the iterate buffers
            solver.reset(S, 0.0, sp)
            if pass_events is not None:
                pass_events[0].record(stream)
            solver.enqueue(npass, sp)
            if pass_events is not None:
                pass_events[1].record(stream)
        else:
            slab.zero_iterate()
            slab.reset(S, 0.0)
            if pass_events is not None:
                pass_events[0].record(stream)
            slab.enqueue(npass)
            if pass_events is not None:
                pass_events[1].record(stream)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step()
    barrier()
    launches0 = L.cnv_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pass_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.time()
    ev0.record(stream)
    for k in range(args.steps):
        one_step(pass_ev[k])
    ev1.record(stream)
    barrier()
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    launches = L.cnv_launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    pass_ms = sum(a.elapsed_time(b) for a, b in pass_ev)
    # verify the work really happened
    st = solver.state(sp) if slab is None else slab.state()
    assert st["sweeps"] == S and st["state"] == 2, st

    # ---- e2e: host buffers in, host psi out, through the C ABI, copies inside the timed region ----
    pin_in = torch.from_numpy(w_host).pin_memory()
    pin_out = torch.empty_like(pin_in).pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    in_np, out_np = pin_in.numpy(), pin_out.numpy()

    # Single GPU: two solver objects on two streams keep two solves in flight, so the host copies of one solve
    # overlap the sweeps of the other (every solve still pays its own H2D, preparation, sweeps and D2H inside the
    # timed region).  Slabs: one solve at a time.
    if slab is None:
        stream2 = torch.cuda.Stream()
        sp2 = C.c_void_p(stream2.cuda_stream)
        solver2 = fd.PoissonSolver(rows_per, ncols, T)
        solver2.set_consts(dx, dy, beta)
        pin_out2 = torch.empty_like(pin_in).pin_memory()
        lanes = [(solver, sp, out_np), (solver2, sp2, pin_out2.numpy())]
    e2e_count = [0]

    def e2e_step():
        if slab is None:
            sv, spx, outx = lanes[e2e_count[0] & 1]
            e2e_count[0] += 1
            sv.upload(in_np, -1.0, spx)               # H2D + rhs preparation + zero guess
            sv.reset(S, 0.0, spx)
            sv.enqueue(npass, spx)
            sv.L.cnv_poisson_download_async(sv.h, npass & 1, outx, spx)   # D2H of psi
        else:
            slab.upload_owned(in_np, -1.0)
            slab.reset(S, 0.0)
            slab.enqueue(npass)
            slab.download_owned(slab.buf_after(npass), out_np)

    e2e_step()
    if slab is None:
        e2e_step()
    barrier()
    e0.record(stream)
    if slab is None:
        stream2.wait_stream(stream)                  # both lanes start after e0
    for _ in range(args.steps):
        e2e_step()
    if slab is None:
        stream.wait_stream(stream2)                  # e1 after the last solve of either lane
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    e2e_seq_ms = None
    if slab is None:
        # for transparency also one solve at a time (a single lane, D2H synchronised before the next H2D)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nseq = max(2, args.steps // 4)
        s0.record(stream)
        for _ in range(nseq):
            solver.upload(in_np, -1.0, sp)
            solver.reset(S, 0.0, sp)
            solver.enqueue(npass, sp)
            solver.L.cnv_poisson_download(solver.h, npass & 1, out_np, sp)
        s1.record(stream)
        torch.cuda.synchronize()
        e2e_seq_ms = s0.elapsed_time(s1) / nseq
        # both lanes produced the field of the resident-input run (same input, same sweeps)
        chk = solver.download(npass & 1)
        assert np.array_equal(out_np, chk) and np.array_equal(lanes[1][2], chk), "e2e result differs from the device-resident run"
        solver2.close()

    if world > 1:
        t = torch.tensor([ms, pass_ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, pass_ms, e2e_ms = t.tolist()
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())

    if rank == 0:
        peak, peak_src = peaks()
        value = interior_cells * S * args.steps / (ms * 1e-3)
        e2e_val = interior_cells * S * args.steps / (e2e_ms * 1e-3)
        # dominant kernel: k_poisson_pass<T>.  Algorithmic bytes per launch = 24 B x interior cells of this
        # GPU's slab x sweeps per launch (T); duration = CUDA-event time around the pass launches / count.
        per_gpu_cells = interior_cells / world
        launch_s = pass_ms * 1e-3 / (npass * args.steps)
        achieved = ALGO_BYTES_PER_CELL_SWEEP * per_gpu_cells * T / launch_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "poisson_pass_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        out = {
            "metric": "poisson_cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{total_rows}x{ncols} Re={5000 if ncols == 16384 else 1000} lid-driven cavity, red-black SOR Poisson solve "
                                   f"({_config_name(args.n)} grid{' per GPU' if args.scaling == 'weak' and world > 1 else ''})",
                       "grid": [total_rows, ncols], "slab_rows_per_gpu": rows_per, "sweeps_per_step": S,
                       "temporal_block_T": T, "strip_width": plan["WS"], "rows_per_chunk": plan["Hout"],
                       "ctas": plan["nstrips"] * plan["nchunks"], "arith_path": "pow2-exact" if plan["pow2"] else "general",
                       "parallelism": f"slab{world}" if world > 1 else "single",
                       **({"exchange": ("peer" if slab.peer else "nccl" if slab.comm else "torch") +
                                       ("+lagged-decision" if slab.peer and len(slab.bufs) == 3 else "") +
                                       ("+tile-kernel" if plan.get("tiled") else "")} if slab is not None else {}),
                       "l2": ("inputs larger than L2 (3 x %.0f MB resident arrays per GPU vs 126 MB L2), no flush" if
                              3 * rows_per * ncols * 8 > 126e6 else
                              "arrays fit L2 at this slab size (3 x %.0f MB per GPU vs 126 MB L2): L2-resident run, no flush") %
                             (rows_per * ncols * 8 / 1e6)},
            "sweeps_per_s": S * args.steps / (ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic if not plan.get("tiled") else None, "peak_source": peak_src,
                         "kernel": (f"k_poisson_tile<M={plan['M']}> (T={T})" if plan.get("tiled") else f"k_poisson_pass<T={T}>"),
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CELL_SWEEP * per_gpu_cells * T,
                         "launch_us": launch_s * 1e6,
                         "note": "temporal blocking: T sweeps per HBM pass, so algorithmic GB/s may exceed the HBM peak"},
            "e2e": {"value": e2e_val, "unit": "cell-updates/s", "h2d_bytes_per_step": int(w_host.nbytes * world),
                    "d2h_bytes_per_step": int(w_host.nbytes * world), "ms_per_step": e2e_ms / args.steps,
                    "sequential_value": (interior_cells * S / (e2e_seq_ms * 1e-3) if e2e_seq_ms else None),
                    "pipeline": ("2 solves in flight: two solver objects on two streams, pinned host buffers; each solve = H2D of w, "
                                 "rhs preparation, sweeps, D2H of psi" if world == 1 else "one solve at a time per slab")},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1:
            out["stencil_phase"] = stencil_phase(fd, torch, args.n, sp, stream, peak)
            out["timestep_1024"] = timestep_1024(fd, peak, cpu=not args.no_cpu)
            if args.n == 4096:
                out["timestep_4096"] = timestep_4096(fd, peak)
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(args.n, args.cpu_seconds)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def stencil_phase(fd, torch, n, sp, stream, peak, reps=20):
    """The explicit-stencil part of one time step (BCs + wall vorticity, fused derivatives + Euler, velocity
    recovery, continuity diagnostic) on the same grid: 72 algorithmic bytes per cell per time step
    (SURVEY.md section 8d).  Reported beside the Poisson number; it is ~0.1 % of a time step at this size."""
    cfg = fd.config_from_dict(dict(nx=n, ny=n, Re=1000.0, dt=5e-6, poisson_max_it=100000))
    sim = fd.Simulation(cfg)
    L = fd.lib()
    L.cnv_sim_stencil_phase(sim.h, 3, sp)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    L.cnv_sim_stencil_phase(sim.h, reps, sp)
    e1.record(stream)
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / reps
    sim.close()
    gbs = 72.0 * n * n / t / 1e9
    return {"cell_updates_per_s": n * n / t, "us_per_step": t * 1e6, "algorithmic_bytes_per_cell": 72, "achieved_gbs": gbs,
            "frac_of_hbm_peak": gbs / peak, "kernels": ["k_ring_bc_vorticity", "k_euler_fused", "k_velocity", "k_continuity"]}


def timestep_1024(fd, peak, steps=4, cpu=True):
    """BASELINE config 3 (1024^2, Re 1000): WHOLE time steps (BCs + wall vorticity, derivatives + Euler, Poisson solve
    to the reference's tolerance with its exact sweep count, velocities, continuity) on the device-resident path.
    cell-steps/s = N^2 x steps / t; bytes per cell-step = 72 + 24 x sweeps."""
    n = 1024
    cfg = dict(nx=n, ny=n, Re=1000.0, dt=1e-4, poisson_max_it=20000, poisson_tol=1e-3)
    sim = fd.Simulation(cfg)
    sim.step(2)                                   # warm-up (also fills the pass-count predictor)
    t0 = time.perf_counter()
    r = sim.step(steps)
    dt = (time.perf_counter() - t0) / steps       # cnv_sim_step synchronises once per step
    sweeps = float(np.mean(r["k"])) + 1
    sim.close()
    out = {"grid": [n, n], "steps": steps, "ms_per_step": dt * 1e3, "cell_steps_per_s": n * n / dt,
           "poisson_sweeps_per_step": sweeps, "algorithmic_gbs": (72 + 24 * sweeps) * n * n / dt / 1e9,
           "frac_of_hbm_peak": (72 + 24 * sweeps) * n * n / dt / 1e9 / peak,
           "note": "8 MB fields: L2-resident, launch/latency bound rather than HBM bound"}
    if cpu:
        from oracle import api
        port = api.port()
        t0 = time.perf_counter()
        ref = port.run(dict(api.CONFIG_DEFAULT, **cfg), 1, redblack=True)      # matrix-free port, OpenMP
        tc = time.perf_counter() - t0
        out["cpu_port"] = {"ms_per_step": tc * 1e3, "cell_steps_per_s": n * n / tc, "threads": port.max_threads(),
                           "sweeps": int(ref["k"][0]) + 1, "kind": "port (the reference executable cannot run beyond 128^2)"}
    return out


def timestep_4096(fd, peak, steps=2):
    """BASELINE config 4 (4096^2, Re 1000) as WHOLE time steps on one GPU: every step solves the streamfunction
    Poisson equation to the reference's tolerance (about 6 k sweeps per step at the start of the run) between the
    stencil phases.  cell-steps/s = N^2 x steps / t; bytes per cell-step = 72 + 24 x sweeps."""
    n = 4096
    cfg = dict(nx=n, ny=n, Re=1000.0, dt=5e-6, poisson_max_it=100000, poisson_tol=1e-3)
    sim = fd.Simulation(cfg)
    sim.step(1)                                   # warm-up (fills the pass-count predictor)
    t0 = time.perf_counter()
    r = sim.step(steps)
    dt = (time.perf_counter() - t0) / steps       # cnv_sim_step synchronises once per step
    sweeps = float(np.mean(r["k"])) + 1
    sim.close()
    gbs = (72 + 24 * sweeps) * n * n / dt / 1e9
    return {"grid": [n, n], "steps": steps, "ms_per_step": dt * 1e3, "cell_steps_per_s": n * n / dt,
            "poisson_sweeps_per_step": sweeps, "poisson_cell_updates_per_s": (n - 2) ** 2 * sweeps / dt,
            "algorithmic_gbs": gbs, "frac_of_hbm_peak": gbs / peak}


# ------------------------------------------------------------------------------------------------
_CHILD = r"""
import sys, time, numpy as np
sys.path.insert(0, {root!r})
from oracle import api
R = api.ref()
n, k = {n}, {k}
from bench import synthetic_vorticity
f = -synthetic_vorticity(n, n)
beta = api.port().beta(n, n)
print("START", time.time(), flush=True)
R.poisson(f, 1.0 / n, 1.0 / n, k, 0.0, beta, sor=True)   # never converges (tol = 0): k sweeps, then exit(1)
"""


def ref_sweeps_time(n, k):
    """Wall time of k red-black sweeps of the UNMODIFIED poisson_SOR_log (oracle/_ref): run in a child
    process with itmax = k, tol = 0 -- it performs exactly k sweeps and then exit(1)s (src/poisson.c:280-284)."""
    with tempfile.TemporaryDirectory() as td:
        p = subprocess.Popen([sys.executable, "-c", _CHILD.format(root=ROOT, n=n, k=k)], stdout=subprocess.PIPE, text=True, cwd=td)
        start = None
        for line in p.stdout:
            if line.startswith("START"):
                start = float(line.split()[1])
        p.wait()
        end = time.time()
    if start is None or p.returncode != 1:
        raise RuntimeError("reference child did not run to its itmax exit")
    return end - start


def cpu_sweep_rate(n, seconds, prefer_ref=True):
    """(cell-updates/s, kind, cores, sample description) of the CPU reference path on this box."""
    from oracle import api
    cores = os.cpu_count() or 1
    port = api.port()
    f = -synthetic_vorticity(n, n)
    beta = port.beta(n, n)
    t0 = time.time()
    port.poisson_sweeps(f, 1.0 / n, 1.0 / n, 2, beta)
    per_sweep = (time.time() - t0) / 2
    k = int(max(2, min(2000, seconds / max(per_sweep, 1e-6))))
    if prefer_ref and api.ref() is not None:
        t = ref_sweeps_time(n, k)
        kind = "reference"
        what = "unmodified poisson_SOR_log (oracle/_ref, -O2 -fopenmp -DOPENMP_ENABLED, red-black)"
    else:
        t0 = time.time()
        port.poisson_sweeps(f, 1.0 / n, 1.0 / n, k, beta)
        t = time.time() - t0
        kind = "port"
        what = "oracle/cnavier_oracle.c orc_poisson_sweeps (OpenMP red-black)"
    threads = port.max_threads()
    return (n - 2) ** 2 * k / t, kind, threads, f"{k} sweeps of the {n}x{n} grid, {what}, {threads} OpenMP threads on {cores} host cores"


def cpu_baseline(n, seconds):
    v, kind, threads, sample = cpu_sweep_rate(n, seconds)
    return {"value": v, "unit": "cell-updates/s", "cores": threads, "kind": kind, "sample": sample}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    kind = threads = sample = None
    for i in range(args.warmup + args.steps):
        v, kind, threads, sample = cpu_sweep_rate(args.n, per_step)
        if i >= args.warmup:
            vals.append(v)
    value = float(len(vals) / sum(1.0 / v for v in vals))  # total work / total time
    total_rows = args.n * args.gpus if args.scaling == "weak" else args.n
    # time one full step (args.sweeps sweeps of the whole grid) would take at the sampled rate
    ms_equiv = (total_rows - 2) * (args.n - 2) * args.sweeps / value * 1e3
    out = {"impl": "reference", "metric": "poisson_cell_updates_per_s", "value": value, "unit": "cell-updates/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_equiv, "higher_is_better": True,
           "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{total_rows}x{args.n} Re={5000 if args.n == 16384 else 1000} lid-driven cavity, red-black SOR Poisson solve ({_config_name(args.n)} grid)",
                      "note": "CPU arm: each step is a bounded sample of sweeps on the 4096x4096 grid; rate is size-independent per cell; "
                              "ms_per_step is the time one full step would take at that rate"},
           "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": threads, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--grid", dest="n", type=int, default=4096,
                    help="grid columns (and rows per GPU under weak scaling); use --grid under torchrun, whose own parser trips over --n")
    ap.add_argument("--sweeps", type=int, default=1024, help="red-black SOR sweeps per step")
    ap.add_argument("--T", type=int, default=0, help="temporal block depth (0 = library default)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 3)   # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
