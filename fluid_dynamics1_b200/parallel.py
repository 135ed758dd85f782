"""Slab-decomposed Poisson solve over the GPUs of one node: one process per GPU (torch.distributed,
NCCL over NVLink/NVSwitch for the plumbing), rows split into contiguous slabs.

The reference has no distributed code; this is the multi-GPU layer of
BASELINE.json's north_star: per-pass halo exchange of 2T rows with the two slab neighbours and ONE
all-reduce of the T per-sweep L1 update norms per pass, after which every rank takes the same stopping
decision on the device (k_decide).  Red-black colouring uses the GLOBAL (i + j) parity, so the fields are
bit-identical to the single-GPU result for any number of slabs.

Per pass p (device-side state machine, no host synchronisation):
    k_poisson_pass (T sweeps on the slab, out of place)         -> local norms[T]
    send/recv 2T boundary rows of the output buffer to/from both neighbours
    all_reduce(norms, SUM)
    k_decide
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .solver import PoissonSolver


def slab_bounds(total_rows: int, world: int, rank: int):
    """Rows [r0, r1) owned by `rank`: contiguous, sizes differing by at most one row."""
    base, rem = divmod(total_rows, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def slab_layout(total_rows: int, world: int, rank: int, T: int):
    """Local array layout of one slab: (grow0, nrows_local, own_lo, own_hi, halo_lo, halo_hi).
    Interior slab edges carry 2T halo rows (the dependency reach of T red-black sweeps)."""
    r0, r1 = slab_bounds(total_rows, world, rank)
    hlo = 2 * T if rank > 0 else 0
    hhi = 2 * T if rank < world - 1 else 0
    return r0 - hlo, (r1 - r0) + hlo + hhi, hlo, hlo + (r1 - r0), hlo, hhi


class _DevView:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class SlabPoisson:
    def __init__(self, total_rows: int, ncols: int, T: int, rank: int, world: int, stream=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = rank, world
        self.total_rows, self.ncols = total_rows, ncols
        T = T or 8
        for r in range(world):
            a, b = slab_bounds(total_rows, world, r)
            if b - a < 2 * T:
                raise ValueError(f"slab of rank {r} has {b - a} rows < halo depth {2 * T}")
        self.grow0, self.nrows, self.own_lo, self.own_hi, self.hlo, self.hhi = slab_layout(total_rows, world, rank, T)
        self.row0 = self.grow0 + self.own_lo
        self.own_rows = self.own_hi - self.own_lo
        self.solver = PoissonSolver(self.nrows, ncols, T, slab=(self.grow0, total_rows, self.own_lo, self.own_hi))
        self.T = self.solver.T
        self.H = 2 * self.T
        self.L = self.solver.L
        self.h = self.solver.h
        self.ld = self.solver.ld
        self.stream = stream
        self.L.cnv_poisson_set_distributed(self.h, 1)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.bufs = [torch.as_tensor(_DevView(self.L.cnv_poisson_buf_ptr(self.h, i), (self.nrows, self.ld)), device=dev)
                     for i in range(2)]
        self.rhs = torch.as_tensor(_DevView(self.L.cnv_poisson_rhs_ptr(self.h), (self.nrows, self.ld)), device=dev)
        self.norms = torch.as_tensor(_DevView(self.L.cnv_poisson_norms_ptr(self.h), (8,)), device=dev)
        self.passes_enqueued = 0

    # ---- data movement ----------------------------------------------------------------------
    def set_consts(self, dx, dy, beta):
        self.solver.set_consts(dx, dy, beta)

    def exchange_halos(self, t):
        """2T boundary rows of tensor `t` (local array, pitch ld) to/from both slab neighbours."""
        dist, H = self.dist, self.H
        ops = []
        if self.rank > 0:  # lower neighbour: my first owned rows -> its high halo; its last owned rows -> my low halo
            ops.append(dist.P2POp(dist.isend, t[self.own_lo:self.own_lo + H], self.rank - 1))
            ops.append(dist.P2POp(dist.irecv, t[0:H], self.rank - 1))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, t[self.own_hi - H:self.own_hi], self.rank + 1))
            ops.append(dist.P2POp(dist.irecv, t[self.own_hi:self.own_hi + H], self.rank + 1))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def upload_owned(self, f_owned, fsign=1.0):
        """Owned rows of the right-hand side from a host array; halo rows come from the neighbours."""
        torch = self.torch
        f_owned = np.ascontiguousarray(f_owned, dtype=np.float64)
        assert f_owned.shape == (self.own_rows, self.ncols)
        src = torch.from_numpy(f_owned)
        stage = torch.empty((self.nrows, self.ncols), dtype=torch.float64, device=self.rhs.device)
        stage.zero_()
        stage[self.own_lo:self.own_hi].copy_(src, non_blocking=True)
        self.exchange_halos(stage)
        self.L.cnv_poisson_prepare(self.h, C.c_void_p(stage.data_ptr()), self.ncols, fsign, self.stream)
        self._stage = stage  # keep alive until the stream has consumed it

    def zero_iterate(self):
        self.L.cnv_poisson_prepare(self.h, None, 0, 1.0, self.stream)

    def download_owned(self, which, out):
        t = self.bufs[which & 1][self.own_lo:self.own_hi, :self.ncols]
        self.torch.from_numpy(out).copy_(t)  # synchronising D2H
        return out

    # ---- solve ----------------------------------------------------------------------------------
    def reset(self, itmax, tol):
        self.solver.reset(itmax, tol, self.stream)
        self.passes_enqueued = 0

    def enqueue(self, npasses):
        for _ in range(npasses):
            self.L.cnv_poisson_enqueue(self.h, 1, self.stream)
            # static host pattern: pass p writes buffer (p+1)&1 (a "redo" pass breaks the alternation only
            # for the final pass, whose halos are never read again)
            self.exchange_halos(self.bufs[(self.passes_enqueued + 1) & 1])
            self.dist.all_reduce(self.norms)
            self.L.cnv_poisson_enqueue_decide(self.h, self.stream)
            self.passes_enqueued += 1

    def state(self):
        return self.solver.state(self.stream)

    def solve(self, itmax, tol, first_batch=16):
        """Run to convergence / itmax with the reference's stopping rule; every rank returns the same dict."""
        self.zero_iterate()
        self.reset(itmax, tol)
        max_passes = (itmax + self.T - 1) // self.T + 2
        batch = first_batch
        while True:
            self.enqueue(min(batch, max_passes + 1 - self.passes_enqueued))
            st = self.state()
            if st["state"] != 0:
                break
            if self.passes_enqueued > max_passes:
                raise RuntimeError("Poisson state machine did not terminate")
            batch = 4
        return dict(status=0 if st["state"] == 1 else 1, k=st["k"], e=st["e"] if st["state"] == 1 else st["last_e"],
                    sweeps=st["sweeps"], passes=st["passes"], buf=st["cur"])

    def gather_result(self, which):
        """Full field on rank 0 (tests / small grids)."""
        torch, dist = self.torch, self.dist
        mine = self.bufs[which & 1][self.own_lo:self.own_hi, :self.ncols].contiguous()
        sizes = [slab_bounds(self.total_rows, self.world, r) for r in range(self.world)]
        if self.rank == 0:
            parts = [torch.empty((b - a, self.ncols), dtype=torch.float64, device=mine.device) for a, b in sizes]
            parts[0].copy_(mine)
            for r in range(1, self.world):
                dist.recv(parts[r], r)
            return torch.cat(parts).cpu().numpy()
        dist.send(mine, 0)
        return None
