#!/usr/bin/env python
"""bench.py -- headline benchmark of the cnavier hot path on B200.

Metric (BASELINE.json): Poisson cell-updates/s (and sweeps/s) against the HBM roofline at 1/2/4/8 B200.
A "step" = one fixed-sweep red-black SOR solve (`--sweeps`, default 1024) of lap(psi) = -w from a zero initial guess on a
synthetic cavity-like vorticity field (zero guess + state reset are inside the timed region; tol = 0, so no work is ever
skipped).

Workloads (`--series auto`, the default):
  N = 1   the 4096^2 Re=1000 cavity grid (BASELINE config 4, the configuration the north_star target is quoted on).
  N > 1   headline `value`: the WEAK series on BASELINE config 5's per-GPU shape -- 2048 x 16384 rows per GPU, i.e. a
          (2048 N) x 16384 grid that is exactly 16384^2 at N = 8;
          `strong`: BASELINE config 4 itself, 4096^2 split over the N GPUs (strong scaling);
          `weak_4096_rows_per_gpu`: a 4096 x 4096 slab per GPU (the series round 1 reported), side key only;
          `weak_base_1gpu`: every rank alone on a 2048 x 16384 grid in the same run (the weak series' denominator).
Every workload is verified after its timed loop (`parity_check`): a 24-sweep solve (3 passes) on the same data, every
rank's owned rows bitwise against the oracle (oracle/liboracle.so, test infrastructure, outside every timed region).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (N > 1: under torch.distributed.run)
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
NOISE_BLOCK = 256                 # rows per independently seeded noise block (any rank can regenerate any row range)
PARITY_SWEEPS = 24


def synthetic_vorticity(nrows, ncols, row0=0, total_rows=None, seed=1234):
    """Deterministic cavity-like vorticity for global rows [row0, row0 + nrows) of a total_rows x ncols grid: strong near
    the moving lid (last row), smooth bulk, plus small-scale noise so that no cell is trivially zero.  The noise is drawn
    per block of 256 global rows, so a slab (or a slab plus neighbouring rows) sees exactly the values of the whole grid."""
    total_rows = total_rows or nrows
    i = (np.arange(row0, row0 + nrows) / total_rows)[:, None]
    j = (np.arange(ncols) / ncols)[None, :]
    w = -40.0 * np.exp(-((1 - i) * 30.0)) * np.sin(np.pi * j) + 4.0 * np.sin(2 * np.pi * i) * np.sin(3 * np.pi * j)
    for b in range(row0 // NOISE_BLOCK, (row0 + nrows - 1) // NOISE_BLOCK + 1):
        blk = np.random.default_rng([seed, b, ncols]).standard_normal((NOISE_BLOCK, ncols))
        lo, hi = max(row0, b * NOISE_BLOCK), min(row0 + nrows, (b + 1) * NOISE_BLOCK)
        w[lo - row0:hi - row0] += 0.05 * blk[lo - b * NOISE_BLOCK:hi - b * NOISE_BLOCK]
    return np.ascontiguousarray(w)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def measured_traffic(nrows, ncols, T):
    """DRAM bytes per pass launch from the committed ncu capture of exactly this shape and depth, else None."""
    tp = os.path.join(ROOT, "profiles", "poisson_pass_traffic.json")
    try:
        d = json.load(open(tp))
        return d.get("by_shape", {}).get(f"{nrows}x{ncols}:T{T}")
    except Exception:
        return None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


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
def headline_shape(args, world):
    """(total_rows, ncols, rows_per_gpu, scaling, description) of the workload `value` is measured on."""
    if args.series in ("auto", "headline") and world > 1:
        return 2048 * world, 16384, 2048, "weak", (
            f"{2048 * world}x16384 Re=5000 lid-driven cavity, red-black SOR Poisson solve (BASELINE config 5 grid: 2048 x 16384 rows "
            f"per GPU, 16384^2 at 8 GPUs)")
    n = args.n
    name = {1024: "BASELINE config 3", 4096: "BASELINE config 4", 16384: "BASELINE config 5"}.get(n, "custom")
    if args.scaling == "weak":
        return n * world, n, n, "weak", (f"{n * world}x{n} Re={5000 if n == 16384 else 1000} lid-driven cavity, red-black SOR Poisson "
                                         f"solve ({name} grid{' per GPU' if world > 1 else ''})")
    return n, n, n // world, "strong", f"{n}x{n} Re={5000 if n == 16384 else 1000} lid-driven cavity, red-black SOR Poisson solve ({name} grid)"


def workload_config(args, world):
    """The `config` object of the JSON line: identical for both arms (--impl b200 / reference)."""
    total_rows, ncols, rows_per, scaling, desc = headline_shape(args, world)
    return {"workload": desc, "grid": [total_rows, ncols], "rows_per_gpu": rows_per, "sweeps_per_step": args.sweeps,
            "gpus": world, "scaling": scaling,
            "l2": ("inputs larger than L2 (3 x %.0f MB resident arrays per GPU vs 126 MB L2), no flush" if
                   3 * rows_per * ncols * 8 > 126e6 else
                   "arrays fit L2 at this slab size (3 x %.0f MB per GPU vs 126 MB L2): L2-resident run, no flush") %
                  (rows_per * ncols * 8 / 1e6)}


class Ctx:
    """Per-process plumbing: torch for streams / events / torch.distributed, ctypes for the C ABI."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import fluid_dynamics1_b200 as fd
        from fluid_dynamics1_b200 import parallel
        self.torch, self.dist, self.fd, self.parallel = torch, dist, fd, parallel
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
        fd.require_gpu()
        torch.cuda.set_device(self.local_rank)
        self.L = fd.lib()
        self.L.cnv_set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.stream = torch.cuda.current_stream()
        self.sp = C.c_void_p(self.stream.cuda_stream)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def max_over_ranks(self, vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def gather_floats(self, v):
        if self.world == 1:
            return [float(v)]
        t = self.torch.zeros(self.world, device="cuda", dtype=self.torch.float64)
        t[self.rank] = v
        self.dist.all_reduce(t)
        return t.tolist()

    def gather_ints(self, v):
        return [int(round(x)) for x in self.gather_floats(float(v))]

    def all_ok(self, ok):
        if self.world == 1:
            return bool(ok)
        t = self.torch.tensor([1 if ok else 0], device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())


class PoissonRun:
    """One Poisson workload: total_rows x ncols, slab-decomposed over ctx.world GPUs (world = 1: the whole grid), or --
    `single=True` -- this rank alone on the whole grid while the other ranks do the same (weak-scaling denominator)."""

    def __init__(self, ctx, total_rows, ncols, T, single=False):
        fd = ctx.fd
        self.ctx, self.total_rows, self.ncols = ctx, total_rows, ncols
        self.world = 1 if single else ctx.world
        self.dx = self.dy = 1.0 / ncols
        self.beta = fd.sor_beta(ncols, ncols)
        if self.world > 1:
            self.slab = ctx.parallel.SlabPoisson(total_rows, ncols, T, ctx.rank, ctx.world, stream=ctx.sp)
            self.slab.set_consts(self.dx, self.dy, self.beta)
            self.solver = self.slab.solver
            self.row0, self.own_rows = self.slab.row0, self.slab.own_rows
        else:
            self.slab = None
            self.solver = fd.PoissonSolver(total_rows, ncols, T)
            self.solver.set_consts(self.dx, self.dy, self.beta)
            self.row0, self.own_rows = 0, total_rows
        self.T = self.solver.T
        self.plan = dict(self.solver.plan)
        self.w_host = synthetic_vorticity(self.own_rows, ncols, self.row0, total_rows)
        self.interior_cells = (total_rows - 2) * (ncols - 2)
        self.pin_in = self.pin_out = None

    # ---- resident-input steps (the `value` measurement) ----
    def upload(self, host):
        if self.slab is None:
            self.solver.upload(host, -1.0, self.ctx.sp)  # f = -w (src/main.c:348)
        else:
            self.slab.upload_owned(host, -1.0)

    def zero_and_run(self, sweeps, events=None):
        """zero guess, reset the state machine, `sweeps` sweeps (tol = 0)."""
        npass = (sweeps + self.T - 1) // self.T
        sp, stream = self.ctx.sp, self.ctx.stream
        if self.slab is None:
            self.solver.L.cnv_poisson_prepare(self.solver.h, None, 0, 1.0, sp)  # f == NULL: only zero the iterate buffers
            self.solver.reset(sweeps, 0.0, sp)
        else:
            self.slab.zero_iterate()
            self.slab.reset(sweeps, 0.0)
        if events is not None:
            events[0].record(stream)
        self.enqueue(npass)
        if events is not None:
            events[1].record(stream)
        return npass

    def enqueue(self, npass):
        if self.slab is None:
            self.solver.enqueue(npass, self.ctx.sp)
        else:
            self.slab.enqueue(npass)

    def state(self):
        return self.solver.state(self.ctx.sp) if self.slab is None else self.slab.state()

    def result_buf(self, npass):
        return npass & 1 if self.slab is None else self.slab.buf_after(npass)

    def download_owned(self, which, out):
        if self.slab is None:
            self.solver.L.cnv_poisson_download(self.solver.h, which, out, self.ctx.sp)
        else:
            self.slab.download_owned(which, out)
        return out

    def timed(self, sweeps, steps, warmup, clocks=None):
        """-> dict(ms, pass_ms, npass, launches, clocks).  CUDA events on the launching stream, barrier + synchronize on both
        sides, max over ranks."""
        ctx = self.ctx
        self.upload(self.w_host)
        for _ in range(warmup):
            self.zero_and_run(sweeps)
        ctx.barrier()
        launches0 = ctx.L.cnv_launch_count()
        ev0, ev1 = ctx.event(), ctx.event()
        pass_ev = [(ctx.event(), ctx.event()) for _ in range(steps)]
        ctx.barrier()
        wall0 = time.time()
        ev0.record(ctx.stream)
        npass = 0
        for k in range(steps):
            npass = self.zero_and_run(sweeps, pass_ev[k])
        ev1.record(ctx.stream)
        ctx.barrier()
        wall1 = time.time()
        launches = ctx.L.cnv_launch_count() - launches0
        ms = ev0.elapsed_time(ev1)
        pass_ms = sum(a.elapsed_time(b) for a, b in pass_ev)
        st = self.state()   # verify the work really happened
        assert st["sweeps"] == sweeps and st["state"] == 2, st
        if self.world > 1:
            ms, pass_ms = ctx.max_over_ranks([ms, pass_ms])
        return dict(ms=ms, pass_ms=pass_ms, npass=npass, launches=launches, wall=(wall0, wall1))

    # ---- parity: the workload's own data, every rank's owned rows bitwise against the oracle ----
    def parity_check(self, sweeps=PARITY_SWEEPS):
        from oracle import api   # the checker; never inside a timed region
        port = api.port()
        ctx = self.ctx
        self.upload(self.w_host)
        npass = self.zero_and_run(sweeps)
        st = self.state()
        got = np.empty((self.own_rows, self.ncols))
        self.download_owned(self.result_buf(npass), got)
        # oracle on the owned rows plus enough neighbouring rows that the artificial zero ring cannot reach them: a sweep
        # moves information by 2 rows, and the first row must keep the global (i + j) colouring (even row offset)
        reach = 2 * sweeps + 2
        g0 = max(0, self.row0 - reach)
        g0 -= g0 & 1
        g1 = min(self.total_rows, self.row0 + self.own_rows + reach)
        f = -synthetic_vorticity(g1 - g0, self.ncols, g0, self.total_rows)
        want, norms = port.poisson_sweeps(f, self.dx, self.dy, sweeps, self.beta)
        want = want[self.row0 - g0:self.row0 - g0 + self.own_rows]
        ok = st["sweeps"] == sweeps and got.tobytes() == want.tobytes()
        return ctx.all_ok(ok) if self.world > 1 else ok

    # ---- end to end: host buffers in, host psi out, through the C ABI, copies inside the timed region ----
    def e2e(self, sweeps, steps):
        ctx, torch = self.ctx, self.ctx.torch
        if self.pin_in is None:   # the library's own host allocator: page-locked, on the GPU's NUMA node (what allocm() hands out)
            self.pin_in = ctx.fd.host_empty(self.w_host.shape)
            self.pin_in[...] = self.w_host
            self.pin_out = ctx.fd.host_empty(self.w_host.shape)
        in_np, out_np = self.pin_in, self.pin_out
        npass = (sweeps + self.T - 1) // self.T

        def step():
            self.upload(in_np)                                  # H2D of w + rhs preparation (+ halo exchange) + zero guess
            if self.slab is None:
                self.solver.reset(sweeps, 0.0, ctx.sp)
            else:
                self.slab.reset(sweeps, 0.0)
            self.enqueue(npass)
            self.download_owned(self.result_buf(npass), out_np)  # D2H of psi, synchronised: one solve at a time

        step()
        ctx.barrier()
        e0, e1 = ctx.event(), ctx.event()
        e0.record(ctx.stream)
        for _ in range(steps):
            step()
        e1.record(ctx.stream)
        ctx.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            ms = ctx.max_over_ranks([ms])[0]
        self.host_numa_node = int(ctx.fd.lib().cnv_host_numa_node())
        return ms, out_np

    def close(self):
        if self.slab is not None:
            self.slab.close()
        else:
            self.solver.close()


def pipelined_e2e(ctx, run, sweeps, steps):
    """Single GPU, side figure: two solver objects on two streams keep two INDEPENDENT solves in flight, so the host copies
    of one overlap the sweeps of the other (every solve still pays its own H2D, preparation, sweeps and D2H inside the
    timed region).  A time-stepping caller has one dependent solve per step and sees the sequential figure."""
    torch, fd = ctx.torch, ctx.fd
    stream2 = torch.cuda.Stream()
    sp2 = C.c_void_p(stream2.cuda_stream)
    solver2 = fd.PoissonSolver(run.total_rows, run.ncols, run.T)
    solver2.set_consts(run.dx, run.dy, run.beta)
    in_np = run.pin_in
    pin_out2 = fd.host_empty(run.pin_in.shape)
    lanes = [(run.solver, ctx.sp, run.pin_out), (solver2, sp2, pin_out2)]
    npass = (sweeps + run.T - 1) // run.T
    count = [0]

    def step():
        sv, spx, outx = lanes[count[0] & 1]
        count[0] += 1
        sv.upload(in_np, -1.0, spx)
        sv.reset(sweeps, 0.0, spx)
        sv.enqueue(npass, spx)
        sv.L.cnv_poisson_download_async(sv.h, npass & 1, outx, spx)

    step(); step()
    ctx.barrier()
    e0, e1 = ctx.event(), ctx.event()
    e0.record(ctx.stream)
    stream2.wait_stream(ctx.stream)                  # both lanes start after e0
    for _ in range(steps):
        step()
    ctx.stream.wait_stream(stream2)                  # e1 after the last solve of either lane
    e1.record(ctx.stream)
    ctx.barrier()
    ms = e0.elapsed_time(e1)
    same = np.array_equal(lanes[0][2], lanes[1][2])
    solver2.close()
    return ms, same


def dropin_e2e(ctx, run, reps=3):
    """The call a C user of the reference makes: poisson_SOR_log(mtrx f, dx, dy, itmax, tol, beta, FILE*) of
    libcnavier_dropin.so (include/poisson.h:16 of the reference) on the 4096^2 grid, host mtrx in, caller-owned host mtrx
    out, solved to the reference's tolerance (1e-3).  Timed with the host clock around the call; beside it the same solve
    through the pinned-buffer solver object (upload + solve + download), one at a time."""
    fd = ctx.fd
    D = fd.dropin()
    n, m = run.total_rows, run.ncols
    tol, itmax = 1e-3, 400000
    f = -run.w_host
    F = D.initm(n, m)
    C.memmove(F.M[0], f.ctypes.data, f.nbytes)       # the drop-in allocm backs a mtrx with one contiguous (page-locked) block
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    out = {}
    with tempfile.TemporaryDirectory() as td:
        logp = os.path.join(td, "log.txt")
        times, k = [], None
        U = None
        for i in range(reps + 1):
            fh = libc.fopen(logp.encode(), b"w")
            t0 = time.perf_counter()
            if U is not None:
                D.freem(U)                       # like main.c: psi.M = freem(psi) before psi = poisson_SOR_log(...) (src/main.c:347-355)
            U = D.poisson_SOR_log(F, run.dx, run.dy, itmax, tol, run.beta, fh)
            t1 = time.perf_counter()
            libc.fclose(fh)
            if i > 0:
                times.append(t1 - t0)
            k = int(open(logp).read().split()[4])
        psi_dropin = np.ctypeslib.as_array(U.M[0], shape=(n * m,)).reshape(n, m).copy()
        D.freem(U)
        D.freem(F)
    t_dropin = float(np.mean(times))
    # the same solve through the solver object with page-locked buffers
    s = run.solver
    pin_f = ctx.fd.host_empty(f.shape)
    pin_f[...] = f
    pin_u = ctx.fd.host_empty(f.shape)
    times2, k2 = [], None
    for i in range(reps + 1):
        t0 = time.perf_counter()
        s.upload(pin_f, 1.0, ctx.sp)
        r = s.solve(itmax, tol, ctx.sp)
        s.L.cnv_poisson_download(s.h, r["buf"], pin_u, ctx.sp)
        t1 = time.perf_counter()
        if i > 0:
            times2.append(t1 - t0)
        k2 = r["k"]
    t_obj = float(np.mean(times2))
    cells = (n - 2) * (m - 2)
    out = {"call": "poisson_SOR_log(mtrx f, dx, dy, itmax, tol = 1e-3, beta, FILE*) of libcnavier_dropin.so, host mtrx in / caller-owned host mtrx out",
           "sweeps": k + 1, "ms_per_call": t_dropin * 1e3, "value": cells * (k + 1) / t_dropin, "unit": "cell-updates/s",
           "solver_object_pinned_ms_per_call": t_obj * 1e3, "solver_object_pinned_value": cells * (k2 + 1) / t_obj,
           "dropin_over_solver_object_time": t_dropin / t_obj,
           "same_result": bool(k == k2 and np.array_equal(psi_dropin, pin_u))}
    return out


def series_entry(ctx, run, args, steps, peak, with_e2e=True):
    """Measure one workload: timed loop, parity check, (optionally) e2e."""
    t = run.timed(args.sweeps, steps, args.warmup)
    parity = run.parity_check()
    per_gpu_cells = run.interior_cells / run.world
    launch_s = t["pass_ms"] * 1e-3 / (t["npass"] * steps)
    achieved = ALGO_BYTES_PER_CELL_SWEEP * per_gpu_cells * run.T / launch_s / 1e9
    out = {"grid": [run.total_rows, run.ncols], "value": run.interior_cells * args.sweeps * steps / (t["ms"] * 1e-3),
           "unit": "cell-updates/s", "ms_per_step": t["ms"] / steps, "steps": steps,
           "sweeps_per_s": args.sweeps * steps / (t["ms"] * 1e-3), "pass_launch_us": launch_s * 1e6,
           "frac": achieved / peak, "achieved_gbs_per_gpu": achieved,
           "parity_check": "bitwise-ok" if parity else "MISMATCH",
           "plan": {"temporal_block_T": run.T, "strip_width": run.plan["WS"], "rows_per_chunk": run.plan["Hout"],
                    "ctas": run.plan["nstrips"] * run.plan["nchunks"], "arith_path": "pow2-exact" if run.plan["pow2"] else "general"}}
    if run.slab is not None:
        out["plan"]["exchange"] = "peer" if run.slab.peer else "nccl" if run.slab.comm else "torch"
    if with_e2e:
        ms, _ = run.e2e(args.sweeps, steps)
        out["e2e"] = {"value": run.interior_cells * args.sweeps * steps / (ms * 1e-3), "ms_per_step": ms / steps}
    out["_timed"] = t
    return out


def run_gpu(args):
    ctx = Ctx(args)
    rank, world = ctx.rank, ctx.world
    # the oracle is only the checker of parity_check(); give it this rank's share of the host cores (torchrun exports
    # OMP_NUM_THREADS=1)
    os.environ["OMP_NUM_THREADS"] = str(max(1, host_cores() // world))
    peak, peak_src = peaks()
    total_rows, ncols, rows_per, scaling, _ = headline_shape(args, world)
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()

    run = PoissonRun(ctx, total_rows, ncols, args.T)
    head = series_entry(ctx, run, args, args.steps, peak, with_e2e=True)
    t = head.pop("_timed")
    clocks = sampler.stop(*t["wall"]) if rank == 0 else None
    launches = t["launches"]
    if world > 1:
        lt = ctx.torch.tensor([launches], device="cuda", dtype=ctx.torch.int64)
        ctx.dist.all_reduce(lt)
        launches = int(lt.item())
    T = run.T
    per_gpu_cells = run.interior_cells / world
    out = {
        "metric": "poisson_cell_updates_per_s", "value": head["value"], "unit": "cell-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "plan": head["plan"],
        "sweeps_per_s": head["sweeps_per_s"],
        "parity_check": head["parity_check"],
        "parity_check_what": f"{PARITY_SWEEPS} sweeps ({(PARITY_SWEEPS + T - 1) // T} passes) on the benchmark's own data after the timed loop: every "
                             f"rank's owned rows bitwise against oracle/liboracle.so (red-black restatement of src/poisson.c:238-262)",
        "roofline": {"bound": "hbm", "achieved": head["achieved_gbs_per_gpu"], "peak": peak, "unit": "GB/s", "frac": head["frac"],
                     "traffic": measured_traffic(run.solver.nrows, ncols, T), "peak_source": peak_src,
                     "kernel": f"k_poisson_pass<T={T}>",
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CELL_SWEEP * per_gpu_cells * T,
                     "launch_us": head["pass_launch_us"],
                     "note": "per GPU; temporal blocking: T sweeps per HBM pass, so algorithmic GB/s may exceed the HBM peak; traffic = ncu "
                             "dram bytes per launch for exactly this slab shape (null where not captured)"},
        "e2e": {"value": head["e2e"]["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": int(run.w_host.nbytes * world),
                "d2h_bytes_per_step": int(run.w_host.nbytes * world), "ms_per_step": head["e2e"]["ms_per_step"],
                "pipeline": "one solve at a time per GPU through the C ABI: pinned host w -> H2D + rhs preparation"
                            + (" + halo exchange" if world > 1 else "") + " -> sweeps -> D2H of psi, synchronised before the next solve",
                "host_buffers": "cnv_host_alloc (page-locked, pooled; placed on the GPU's NUMA node where sysfs names it)",
                "host_numa_nodes": ctx.gather_ints(getattr(run, "host_numa_node", -1)) if world > 1 else [getattr(run, "host_numa_node", -1)]},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    ok = head["parity_check"] == "bitwise-ok"

    if world == 1:
        # side figures on one GPU
        ms, same = pipelined_e2e(ctx, run, args.sweeps, args.steps)
        out["e2e"]["pipelined_value"] = run.interior_cells * args.sweeps * args.steps / (ms * 1e-3)
        out["e2e"]["pipelined_note"] = ("two independent solves in flight (two solver objects, two streams): copies of one overlap the "
                                        "sweeps of the other; not available to a time-stepping caller")
        ok = ok and same
        if args.series == "auto":
            out["e2e"]["dropin"] = dropin_e2e(ctx, run)
            ok = ok and out["e2e"]["dropin"]["same_result"]
        run.close()
        if args.series == "auto":
            base = PoissonRun(ctx, 2048, 16384, args.T)
            e = series_entry(ctx, base, args, max(3, args.steps // 2), peak, with_e2e=False)
            e.pop("_timed")
            out["weak_base_config5_shape"] = e
            ok = ok and e["parity_check"] == "bitwise-ok"
            base.close()
        out["stencil_phase"] = stencil_phase(ctx.fd, ctx.torch, args.n, ctx.sp, ctx.stream, peak)
        out["timestep_1024"] = timestep_1024(ctx.fd, peak, cpu=not args.no_cpu)
        if args.n == 4096:
            out["timestep_4096"] = timestep_4096(ctx.fd, peak)
        if not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(args.n, args.cpu_seconds)
    else:
        run.close()
        if args.series == "auto":
            side_steps = max(3, args.steps // 2)
            # the weak series' denominator: every rank alone on one config-5 slab shape, all GPUs of the box busy at once
            base = PoissonRun(ctx, 2048, 16384, args.T, single=True)
            tb = base.timed(args.sweeps, side_steps, args.warmup)
            vals = ctx.gather_floats(base.interior_cells * args.sweeps * side_steps / (tb["ms"] * 1e-3))
            pb = ctx.all_ok(base.parity_check())
            base.close()
            out["weak_base_1gpu"] = {"grid": [2048, 16384], "value_mean": float(np.mean(vals)), "value_min": float(np.min(vals)),
                                     "value_max": float(np.max(vals)), "parity_check": "bitwise-ok" if pb else "MISMATCH",
                                     "note": "each of the N ranks alone on a 2048 x 16384 grid, all N GPUs running at the same time"}
            out["weak_efficiency_vs_base"] = out["value"] / (world * float(np.mean(vals)))
            ok = ok and pb
            # BASELINE config 4 under strong scaling
            srun = PoissonRun(ctx, 4096, 4096, args.T)
            e = series_entry(ctx, srun, args, args.steps, peak, with_e2e=True)
            e.pop("_timed")
            srun.close()
            e["workload"] = "4096x4096 Re=1000 lid-driven cavity (BASELINE config 4), slab-decomposed over the N GPUs: strong scaling"
            out["strong"] = e
            ok = ok and e["parity_check"] == "bitwise-ok"
            # round 1's series
            wrun = PoissonRun(ctx, 4096 * world, 4096, args.T)
            e = series_entry(ctx, wrun, args, side_steps, peak, with_e2e=False)
            e.pop("_timed")
            wrun.close()
            out["weak_4096_rows_per_gpu"] = e
            ok = ok and e["parity_check"] == "bitwise-ok"
            out["timestep_slab"] = timestep_slab(ctx, peak)
            ok = ok and out["timestep_slab"].get("parity_check", "bitwise-ok") == "bitwise-ok"

    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        ctx.dist.destroy_process_group()
    if not ok:
        raise SystemExit("bench.py: a parity check FAILED (see parity_check keys)")


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
           "poisson_sweeps_per_step": sweeps, "us_per_sweep": dt * 1e6 / sweeps, "algorithmic_gbs": (72 + 24 * sweeps) * n * n / dt / 1e9,
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


def timestep_slab(ctx, peak, steps=2):
    """BASELINE config 4 (4096^2, Re 1000) as WHOLE time steps slab-decomposed over the N GPUs (SlabSimulation: stencil phases
    with 3-row halo exchanges, distributed Poisson solve to the reference's tolerance, continuity all-reduce).  The sweep
    counts of the steps must equal those of the single-GPU run (they are a function of the bits of the fields)."""
    n = 4096
    cfg = dict(nx=n, ny=n, Re=1000.0, dt=5e-6, poisson_max_it=100000, poisson_tol=1e-3)
    sim = ctx.parallel.SlabSimulation(cfg, ctx.rank, ctx.world, stream=ctx.sp)
    sim.step(1)
    ctx.barrier()
    t0 = time.perf_counter()
    r = sim.step(steps)
    ctx.barrier()
    dt = ctx.max_over_ranks([(time.perf_counter() - t0) / steps])[0]
    sweeps = float(np.mean(r["k"])) + 1
    sim.close()
    gbs = (72 + 24 * sweeps) * n * n / dt / 1e9
    return {"grid": [n, n], "gpus": ctx.world, "steps": steps, "ms_per_step": dt * 1e3, "cell_steps_per_s": n * n / dt,
            "poisson_sweeps_per_step": sweeps, "poisson_k": [int(k) for k in r["k"]],
            "poisson_cell_updates_per_s": (n - 2) ** 2 * sweeps / dt, "algorithmic_gbs": gbs, "frac_of_hbm_peak_per_gpu": gbs / peak / ctx.world}


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


def cpu_env():
    """Environment of the CPU arm: all host cores (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(host_cores())
    env["OMP_PROC_BIND"] = "false"
    env.pop("OMP_PLACES", None)
    return env


def ref_sweeps_time(n, k):
    """Wall time of k red-black sweeps of the UNMODIFIED poisson_SOR_log (oracle/_ref): run in a child
    process with itmax = k, tol = 0 -- it performs exactly k sweeps and then exit(1)s (src/poisson.c:280-284)."""
    with tempfile.TemporaryDirectory() as td:
        p = subprocess.Popen([sys.executable, "-c", _CHILD.format(root=ROOT, n=n, k=k)], stdout=subprocess.PIPE, text=True, cwd=td,
                             env=cpu_env())
        start = None
        for line in p.stdout:
            if line.startswith("START"):
                start = float(line.split()[1])
        p.wait()
        end = time.time()
    if start is None or p.returncode != 1:
        raise RuntimeError("reference child did not run to its itmax exit")
    return end - start


_CALIB = r"""
import sys, time, numpy as np
sys.path.insert(0, {root!r})
from oracle import api
from bench import synthetic_vorticity
n, k = {n}, {k}
port = api.port()
f = -synthetic_vorticity(n, n)
t0 = time.time()
port.poisson_sweeps(f, 1.0 / n, 1.0 / n, k, port.beta(n, n))
print("PER_SWEEP", (time.time() - t0) / k, port.max_threads(), flush=True)
"""


def cpu_sweep_rate(n, seconds, prefer_ref=True):
    """(cell-updates/s, kind, threads, sample description, sample seconds) of the CPU reference path on this box."""
    from oracle import api
    cores = host_cores()
    # calibrate the sample size in a child as well, so that it runs with all host cores whatever this process' OpenMP state
    c = subprocess.run([sys.executable, "-c", _CALIB.format(root=ROOT, n=n, k=2)], capture_output=True, text=True, env=cpu_env())
    per_sweep, threads = 1.0, cores
    for line in c.stdout.splitlines():
        if line.startswith("PER_SWEEP"):
            per_sweep, threads = float(line.split()[1]), int(line.split()[2])
    k = int(max(2, min(2000, seconds / max(per_sweep, 1e-6))))
    if prefer_ref and api.ref() is not None:
        t = ref_sweeps_time(n, k)
        kind = "reference"
        what = "unmodified poisson_SOR_log (oracle/_ref, -O2 -fopenmp -DOPENMP_ENABLED, red-black)"
    else:
        c = subprocess.run([sys.executable, "-c", _CALIB.format(root=ROOT, n=n, k=k)], capture_output=True, text=True, env=cpu_env())
        t = [float(x.split()[1]) for x in c.stdout.splitlines() if x.startswith("PER_SWEEP")][0] * k
        kind = "port"
        what = "oracle/cnavier_oracle.c orc_poisson_sweeps (OpenMP red-black)"
    return ((n - 2) ** 2 * k / t, kind, threads,
            f"{k} sweeps of the {n}x{n} grid, {what}, {threads} OpenMP threads on {cores} host cores", t)


def cpu_baseline(n, seconds):
    v, kind, threads, sample, _ = cpu_sweep_rate(n, seconds)
    return {"value": v, "unit": "cell-updates/s", "cores": threads, "kind": kind, "sample": sample}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores.  Each step is a bounded
    sample of sweeps (the rate per cell is size independent; the sample runs on the 4096-wide BASELINE config 4 grid)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    per_step = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals, secs = [], []
    kind = threads = sample = None
    for i in range(args.warmup + args.steps):
        v, kind, threads, sample, t = cpu_sweep_rate(args.n, per_step)
        if i >= args.warmup:
            vals.append(v); secs.append(t)
    value = float(len(vals) / sum(1.0 / v for v in vals))  # total work / total time
    total_rows, ncols, _, scaling, _ = headline_shape(args, world)
    out = {"impl": "reference", "metric": "poisson_cell_updates_per_s", "value": value, "unit": "cell-updates/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True,
           "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, world),
           "step_is": "a bounded sample of sweeps of the same solve (see cpu_baseline.sample); ms_per_step is its measured wall time",
           "ms_per_full_step_equiv": (total_rows - 2) * (ncols - 2) * args.sweeps / value * 1e3,
           "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": threads, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--series", default="auto", choices=["auto", "headline", "single"],
                    help="auto: the BASELINE workloads (see the module docstring); single: only the grid given by --grid / --scaling")
    ap.add_argument("--n", "--grid", dest="n", type=int, default=4096,
                    help="grid columns (and rows per GPU under weak scaling) of --series single; use --grid under torchrun, whose own "
                         "parser trips over --n")
    ap.add_argument("--sweeps", type=int, default=1024, help="red-black SOR sweeps per step")
    ap.add_argument("--T", type=int, default=0, help="temporal block depth (0 = library default)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="--series single at N > 1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.n != 4096 or args.scaling != "weak":
        args.series = "single"
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 3)   # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
