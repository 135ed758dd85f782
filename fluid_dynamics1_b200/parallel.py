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
import os

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


_native_comm = {}  # (rank, world) -> cnv_comm handle (one NCCL communicator per process)


def native_comm(L, dist, torch, rank, world):
    """The library's own NCCL communicator (cnv_comm_*), bootstrapped through torch.distributed: rank 0 creates the
    NCCL unique id, it is broadcast as a Python object, every rank joins.  Returns None when disabled
    (CNV_DIST_BACKEND=torch) or unavailable on any rank; the callers then fall back to torch.distributed p2p."""
    key = (rank, world)
    if key in _native_comm:
        return _native_comm[key]
    handle = None
    if os.environ.get("CNV_DIST_BACKEND", "nccl") != "torch":
        ids = [None]
        if rank == 0:
            buf = C.create_string_buffer(128)
            if L.cnv_comm_unique_id(buf) == 0:
                ids = [buf.raw]
        dist.broadcast_object_list(ids, src=0)
        if ids[0] is not None:
            handle = L.cnv_comm_create(rank, world, ids[0])
        ok = torch.tensor([1 if handle else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            handle = None
    _native_comm[key] = handle
    return handle


class SlabPoisson:
    def __init__(self, total_rows: int, ncols: int, T: int, rank: int, world: int, stream=None, handle=None):
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
        self.solver = PoissonSolver(self.nrows, ncols, T, slab=(self.grow0, total_rows, self.own_lo, self.own_hi), handle=handle)
        self.T = self.solver.T
        self.H = 2 * self.T
        self.L = self.solver.L
        self.h = self.solver.h
        self.ld = self.solver.ld
        self.stream = stream
        self.L.cnv_poisson_set_distributed(self.h, 1)
        dev = torch.device("cuda", torch.cuda.current_device())
        self._dev = dev
        self._map_bufs()
        self.rhs = torch.as_tensor(_DevView(self.L.cnv_poisson_rhs_ptr(self.h), (self.nrows, self.ld)), device=dev)
        self.norms = torch.as_tensor(_DevView(self.L.cnv_poisson_norms_ptr(self.h), (8,)), device=dev)
        self.passes_enqueued = 0
        # per-pass exchange, best first (CNV_DIST_BACKEND = peer | nccl | torch):
        #   peer   boundary rows stored into the neighbours' halos by the pass kernel itself over NVLink (CUDA IPC),
        #          norms published in every rank's mailbox: one kernel per pass, no collective launch
        #   nccl   the library's own NCCL group on the compute stream (halos + norm all-gather) + decide kernel
        #   torch  torch.distributed batch_isend_irecv + all_reduce
        backend = os.environ.get("CNV_DIST_BACKEND", "peer")
        self.comm = native_comm(self.L, dist, torch, rank, world)
        if self.comm:
            self.L.cnv_poisson_attach_comm(self.h, self.comm)
        self.peer = backend == "peer" and world <= 8 and self._setup_peer()
        self._map_bufs()
        self.solver.refresh_plan()

    def _map_bufs(self):
        n = self.L.cnv_poisson_num_buffers(self.h)
        self.bufs = [self.torch.as_tensor(_DevView(self.L.cnv_poisson_buf_ptr(self.h, i), (self.nrows, self.ld)), device=self._dev)
                     for i in range(n)]

    def _setup_peer(self):
        """Exchange CUDA-IPC handles / push counts and map the neighbours' buffers; all ranks or none."""
        L, dist, torch = self.L, self.dist, self.torch
        buf = C.create_string_buffer(256)
        L.cnv_poisson_peer_export(self.h, buf)
        lo, hi = C.c_longlong(), C.c_longlong()
        L.cnv_poisson_peer_push_counts(self.h, self.rank, self.world, C.byref(lo), C.byref(hi))
        mine = (buf.raw, [self.own_lo, self.own_hi, int(lo.value), int(hi.value)])
        allr = [None] * self.world
        dist.all_gather_object(allr, mine)
        handles = b"".join(a[0] for a in allr)
        layout = (C.c_int * (4 * self.world))(*[x for a in allr for x in a[1]])
        rc = L.cnv_poisson_peer_import(self.h, self.rank, self.world, handles, layout)
        ok = torch.tensor([1 if rc == 0 else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            L.cnv_poisson_peer_disable(self.h)
            return False
        return True

    # ---- data movement ----------------------------------------------------------------------
    def set_consts(self, dx, dy, beta):
        self.solver.set_consts(dx, dy, beta)

    def exchange_halos(self, t, depth=None):
        """`depth` (default 2T) boundary rows of tensor `t` (local array, pitch ld) to/from both slab neighbours."""
        dist, H = self.dist, self.H
        d = H if depth is None else depth
        if self.comm and tuple(t.shape) == (self.nrows, self.ld) and t.is_contiguous():
            self.L.cnv_poisson_exchange_halos(self.h, C.c_void_p(t.data_ptr()), d, self.stream)
            return
        ops = []
        if self.rank > 0:  # lower neighbour: my first owned rows -> its high halo; its last owned rows -> my low halo
            ops.append(dist.P2POp(dist.isend, t[self.own_lo:self.own_lo + d], self.rank - 1))
            ops.append(dist.P2POp(dist.irecv, t[self.own_lo - d:self.own_lo], self.rank - 1))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, t[self.own_hi - d:self.own_hi], self.rank + 1))
            ops.append(dist.P2POp(dist.irecv, t[self.own_hi:self.own_hi + d], self.rank + 1))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def upload_owned(self, f_owned, fsign=1.0):
        """Owned rows of the right-hand side from a host array (page-locked for a true DMA); halo rows come from the
        neighbours.  One library call: pitched H2D straight into the solver's right-hand-side array, halo exchange on
        the library's communicator, scaling in place, zero iterate -- no staging array, nothing synchronises."""
        f_owned = np.ascontiguousarray(f_owned, dtype=np.float64)
        assert f_owned.shape == (self.own_rows, self.ncols)
        if self.comm or self.world == 1:
            if self.L.cnv_poisson_upload_owned(self.h, f_owned, fsign, self.stream) == 0:
                self._src = f_owned  # keep alive until the stream has consumed it
                return
        torch = self.torch       # CNV_DIST_BACKEND=torch (no native communicator): stage through a device array
        stage = torch.zeros((self.nrows, self.ncols), dtype=torch.float64, device=self.rhs.device)
        stage[self.own_lo:self.own_hi].copy_(torch.from_numpy(f_owned), non_blocking=True)
        self.exchange_halos(stage)
        self.L.cnv_poisson_prepare(self.h, C.c_void_p(stage.data_ptr()), self.ncols, fsign, self.stream)
        self._stage = stage  # keep alive until the stream has consumed it

    def zero_iterate(self):
        self.L.cnv_poisson_prepare(self.h, None, 0, 1.0, self.stream)

    def buf_after(self, npasses):
        """Index of the buffer holding the iterate after `npasses` full passes of a solve that did not stop early."""
        return npasses % len(self.bufs)

    def download_owned(self, which, out, sync=True):
        """Owned rows of iterate buffer `which` into the host array `out` (own_rows x ncols, page-locked for a true DMA)."""
        assert out.shape == (self.own_rows, self.ncols) and out.flags.c_contiguous
        self.L.cnv_poisson_download_owned_async(self.h, which % len(self.bufs), out, self.stream)
        if sync:
            self.torch.cuda.synchronize()
        return out

    def close(self):
        """Tear-down in the order the peer path needs (include/cnavier_b200.h): quiesce, synchronise, barrier of all
        ranks, unmap the neighbours' buffers, barrier, and only then free this rank's own buffers."""
        if getattr(self, "h", None) is None:
            return
        if self.peer:
            self.L.cnv_poisson_peer_quiesce(self.h, self.stream)
        self.torch.cuda.synchronize()
        if self.world > 1 and self.dist.is_initialized():
            self.dist.barrier()
            self.L.cnv_poisson_peer_close(self.h)
            self.torch.cuda.synchronize()
            self.dist.barrier()
        self.peer = False
        self.bufs, self.rhs, self.norms = [], None, None
        self.solver.close()
        self.h = None

    # ---- solve ----------------------------------------------------------------------------------
    def reset(self, itmax, tol):
        self.solver.reset(itmax, tol, self.stream)
        self.passes_enqueued = 0

    def enqueue(self, npasses):
        if self.peer:  # one kernel per pass: the exchange is fused into it (peer stores over NVLink)
            self.L.cnv_poisson_enqueue(self.h, npasses, self.stream)
            self.passes_enqueued += npasses
            return
        if self.comm:  # one NCCL group + decide kernel per pass, enqueued by the library (no Python per pass)
            self.L.cnv_poisson_enqueue_dist(self.h, npasses, self.stream)
            self.passes_enqueued += npasses
            return
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

    def solve(self, itmax, tol, first_batch=16, rendezvous=False):
        """Run to convergence / itmax with the reference's stopping rule; every rank returns the same dict."""
        if self.peer and rendezvous:
            # host-level rendezvous before the first pass: the in-kernel waits of the peer path then only ever span the skew
            # WITHIN a solve, not whatever the hosts did in between (output files, garbage collection, ...)
            self.dist.barrier()
        self.zero_iterate()
        self.reset(itmax, tol)
        max_passes = (itmax + self.T - 1) // self.T + 2   # (+2: the redo pass and the slack of the batching)
        # like PoissonSolver::solve: sweep counts drift slowly from one time step to the next, so the first batch covers the
        # previous solve's pass count (+ the redo pass) and is followed by small ones
        # -- every read-back of the state is a host synchronisation and, on the peer path, a rendezvous of all ranks
        predicted = getattr(self, "_predicted_passes", 0)
        batch = predicted + 1 if predicted > 0 else first_batch
        if not self.peer:
            batch = min(batch, 32)   # a surplus (no-op) pass still costs its halo exchange + norm collective on these paths
        while True:
            self.enqueue(min(batch, max_passes + 1 - self.passes_enqueued))
            st = self.state()
            if st["state"] != 0:
                break
            if self.passes_enqueued > max_passes:
                raise RuntimeError("Poisson state machine did not terminate")
            batch = 4
        self._predicted_passes = st["passes"]
        return dict(status=0 if st["state"] == 1 else 1, k=st["k"], e=st["e"] if st["state"] == 1 else st["last_e"],
                    sweeps=st["sweeps"], passes=st["passes"], buf=st["cur"])

    def gather_result(self, which):
        """Full field on rank 0 (tests / small grids)."""
        torch, dist = self.torch, self.dist
        mine = self.bufs[which % len(self.bufs)][self.own_lo:self.own_hi, :self.ncols].contiguous()
        sizes = [slab_bounds(self.total_rows, self.world, r) for r in range(self.world)]
        if self.rank == 0:
            parts = [torch.empty((b - a, self.ncols), dtype=torch.float64, device=mine.device) for a, b in sizes]
            parts[0].copy_(mine)
            for r in range(1, self.world):
                dist.recv(parts[r], r)
            return torch.cat(parts).cpu().numpy()
        dist.send(mine, 0)
        return None


class SlabSimulation:
    """Slab-decomposed time stepping (the loop body of src/main.c:283-395) over the GPUs of one node.

    Each rank owns a contiguous block of rows of every field (u, v, w, psi, rhs) plus 2T halo rows per interior
    edge; closures and boundary conditions use GLOBAL row indices, so the fields are bit-identical to the
    single-GPU run.  Per step: phase 0 (BCs + wall vorticity) -> exchange w (3 rows) -> phase 1 (derivatives +
    Euler + rhs) -> exchange rhs (2T rows) -> distributed Poisson solve -> exchange psi (3 rows) -> phase 2
    (velocities) -> exchange u, v (3 rows) -> phase 3 (continuity) -> all-reduce(max) / all-reduce(min)."""

    STENCIL_HALO = 3  # reach of the order-6 stencils

    def __init__(self, cfg, rank: int, world: int, T: int = 0, stream=None):
        import torch
        import torch.distributed as dist
        from .config import config_from_dict
        self.torch, self.dist = torch, dist
        _lib.require_gpu()
        self.L = _lib.lib()
        if isinstance(cfg, dict):
            cfg = config_from_dict(cfg)
        self.cfg, self.rank, self.world, self.stream = cfg, rank, world, stream
        self.h = self.L.cnv_sim_create_slab(C.byref(cfg), T, rank, world)
        lay = (C.c_int * 8)()
        self.L.cnv_sim_layout(self.h, lay)
        self.grow0, self.nloc, self.own_lo, self.own_hi, self.ld, self.ncols, self.T, self.gnrows = list(lay)
        self.poisson = SlabPoisson(self.gnrows, self.ncols, self.T, rank, world, stream=stream,
                                   handle=self.L.cnv_sim_poisson(self.h))
        assert (self.poisson.grow0, self.poisson.nrows, self.poisson.own_lo, self.poisson.own_hi) == \
               (self.grow0, self.nloc, self.own_lo, self.own_hi)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.cont = torch.as_tensor(_DevView(self.L.cnv_sim_field_ptr(self.h, 3), (2,)), device=dev)
        self.psi_buf = 0

    def close(self):
        if self.h:
            self.poisson.close()
            self.L.cnv_sim_destroy(self.h)
            self.h = None

    __del__ = close

    def _field(self, which):
        # w's buffer swaps every step, so views are taken fresh
        return self.torch.as_tensor(_DevView(self.L.cnv_sim_field_ptr(self.h, which), (self.nloc, self.ld)), device=self.dev)

    def step(self, nsteps=1, diagnostics=True):
        L, P, H3 = self.L, self.poisson, self.STENCIL_HALO
        if P.comm:
            # the library's own slab step (csrc/capi.cu cnv_sim_step_slab: stencil phases, halo exchanges on its NCCL communicator,
            # distributed Poisson solve, continuity all-reduce) -- what the C driver runs under CNV_GPUS=N
            k = np.zeros(nsteps, dtype=np.int32)
            e, cmax, cmin = np.zeros(nsteps), np.zeros(nsteps), np.zeros(nsteps)
            failed = L.cnv_sim_step_slab(self.h, nsteps, k.ctypes.data, e.ctypes.data,
                                         cmax.ctypes.data if diagnostics else None, cmin.ctypes.data if diagnostics else None)
            if failed < 0:
                raise RuntimeError("cnv_sim_step_slab: no communicator")
            n = failed if failed else nsteps
            self.psi_buf = None
            return dict(failed_step=failed, k=list(k[:n]), e=list(e[:n]), cont_max=list(cmax[:n]) if diagnostics else [],
                        cont_min=list(cmin[:n]) if diagnostics else [])
        # CNV_DIST_BACKEND=torch (no native communicator): the same step orchestrated through torch.distributed
        ks, es, cmax, cmin = [], [], [], []
        for _ in range(nsteps):
            L.cnv_sim_phase(self.h, 0, self.stream)
            P.exchange_halos(self._field(2), H3)
            L.cnv_sim_phase(self.h, 1, self.stream)
            P.exchange_halos(P.rhs)                       # 2T rows: the solver recomputes its halo
            r = P.solve(self.cfg.poisson_max_it, self.cfg.poisson_tol)
            ks.append(r["k"]); es.append(r["e"])
            if r["status"] != 0:
                return dict(failed_step=len(ks), k=ks, e=es, cont_max=cmax, cont_min=cmin)
            self.psi_buf = r["buf"]
            L.cnv_sim_set_psi_buf(self.h, self.psi_buf)
            P.exchange_halos(P.bufs[self.psi_buf], H3)    # (a "redo" pass leaves the result's halos stale)
            L.cnv_sim_phase(self.h, 2, self.stream)
            P.exchange_halos(self._field(0), H3)
            P.exchange_halos(self._field(1), H3)
            if diagnostics:
                L.cnv_sim_phase(self.h, 3, self.stream)
                mx, mn = self.cont[0:1].clone(), self.cont[1:2].clone()
                self.dist.all_reduce(mx, op=self.dist.ReduceOp.MAX)
                self.dist.all_reduce(mn, op=self.dist.ReduceOp.MIN)
                cmax.append(float(mx.item())); cmin.append(float(mn.item()))
        return dict(failed_step=0, k=ks, e=es, cont_max=cmax, cont_min=cmin)

    def gather_fields(self):
        """psi, w, u, v of the whole grid on rank 0 (None elsewhere)."""
        if self.poisson.comm and self.psi_buf is None:
            shape = (self.gnrows, self.ncols)
            out = {n: np.empty(shape) for n in ("psi", "w", "u", "v")} if self.rank == 0 else None
            ptr = (lambda n: out[n].ctypes.data) if self.rank == 0 else (lambda n: None)
            self.L.cnv_sim_gather_fields_slab(self.h, ptr("psi"), ptr("w"), ptr("u"), ptr("v"))
            return out
        out = {}
        srcs = {"psi": self.poisson.bufs[self.psi_buf], "w": self._field(2), "u": self._field(0), "v": self._field(1)}
        for name, t in srcs.items():
            mine = t[self.own_lo:self.own_hi, :self.ncols].contiguous()
            sizes = [slab_bounds(self.gnrows, self.world, r) for r in range(self.world)]
            if self.rank == 0:
                parts = [self.torch.empty((b - a, self.ncols), dtype=self.torch.float64, device=self.dev) for a, b in sizes]
                parts[0].copy_(mine)
                for r in range(1, self.world):
                    self.dist.recv(parts[r], r)
                out[name] = self.torch.cat(parts).cpu().numpy()
            else:
                self.dist.send(mine, 0)
        return out if self.rank == 0 else None
