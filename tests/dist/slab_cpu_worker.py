"""gloo worker (CPU): the multi-GPU protocol of fluid_dynamics1_b200/parallel.py with the kernel replaced by
the CPU schedule emulator (tests/emul) -- slab layout, 2T-row halo exchange, norm all-reduce, device state
machine (decide) -- checked against the single-domain oracle.  Spawned by tests/test_distributed_cpu.py."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from fluid_dynamics1_b200.parallel import slab_bounds, slab_layout
from oracle import api

dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def main():
    rows, cols, T, tol, itmax = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    E = C.CDLL(os.path.join(ROOT, "tests", "emul", "libstream_emul.so"))
    E.emul_pass.argtypes = [C.c_int] * 10 + [C.c_double] * 3 + [C.c_int, dp, dp, dp, C.c_int, dp]
    ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    E.emul_decide.argtypes = [ip, dp, dp, C.c_int, C.c_void_p]
    E.emul_pass_sweeps.argtypes = [C.c_int] * 4

    rng = np.random.default_rng(11)
    f = rng.standard_normal((rows, cols))
    dx, dy = 1.0 / cols, 1.0 / rows
    port = api.port()
    beta = port.beta(rows, cols)

    grow0, nloc, own_lo, own_hi, hlo, hhi = slab_layout(rows, world, rank, T)
    H = 2 * T
    ld = (cols + 15) // 16 * 16
    floc = np.zeros((nloc, ld))
    floc[:, :cols] = f[grow0:grow0 + nloc]          # rhs incl. halo rows (the GPU path exchanges them once)
    bufs = [np.zeros((nloc, ld)), np.zeros((nloc, ld))]
    ints = np.array([0, 0, 0, 0, itmax, -1, 0], dtype=np.int32)   # state, cur, sweeps, redo, itmax, k, passes
    dbls = np.array([tol, 0.0, 0.0])

    def exchange(a):
        t = torch.from_numpy(a)
        reqs = []
        if rank > 0:
            reqs.append(dist.isend(t[own_lo:own_lo + H].clone(), rank - 1))
            lo = torch.empty((H, ld), dtype=torch.float64)
            reqs.append(dist.irecv(lo, rank - 1))
        if rank < world - 1:
            reqs.append(dist.isend(t[own_hi - H:own_hi].clone(), rank + 1))
            hi = torch.empty((H, ld), dtype=torch.float64)
            reqs.append(dist.irecv(hi, rank + 1))
        for r in reqs:
            r.wait()
        if rank > 0:
            t[0:H] = lo
        if rank < world - 1:
            t[own_hi:own_hi + H] = hi

    p = 0
    while ints[0] == 0:
        nsw = E.emul_pass_sweeps(int(ints[2]), int(ints[3]), itmax, T)
        cur = int(ints[1])
        norms = np.zeros(8)
        rc = E.emul_pass(T, nloc, cols, ld, grow0, rows, own_lo, own_hi, 0, int(os.environ.get("CNV_TEST_CHUNKS", "0")), dx, dy, beta, 0,
                         bufs[cur], floc, bufs[cur ^ 1], nsw, norms)
        assert rc == 0
        # same static pattern as the GPU host loop: exchange the buffer pass p is assumed to have written
        exchange(bufs[(p + 1) & 1])
        tn = torch.from_numpy(norms)
        dist.all_reduce(tn)
        E.emul_decide(ints, dbls, norms, nsw, None)
        p += 1
        assert p < itmax + 4
    # gather owned rows on rank 0
    res = bufs[int(ints[1])][own_lo:own_hi, :cols].copy()
    parts = [None] * world
    dist.all_gather_object(parts, res)
    ok = True
    if rank == 0:
        full = np.concatenate(parts)
        want = port.poisson(f, dx, dy, itmax, tol, beta, redblack=True)
        status = 0 if ints[0] == 1 else 1
        ok = (status == want["status"] and int(ints[5]) == want["k"] and full.tobytes() == want["u"].tobytes()
              and abs(dbls[1] - want["e"]) <= 1e-12 * max(want["e"], 1e-300))
        sizes = [slab_bounds(rows, world, r) for r in range(world)]
        ok = ok and sizes[0][0] == 0 and sizes[-1][1] == rows and all(sizes[i][1] == sizes[i + 1][0] for i in range(world - 1))
        print(f"slab cpu check world={world} {rows}x{cols} T={T}: k={ints[5]} want {want['k']} passes={p} ok={ok}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
