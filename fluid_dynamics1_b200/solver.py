"""numpy-facing host layer over the C ABI: the same operations the reference exposes through
include/{finitediff,fluiddyn,poisson}.h, on dense row-major float64 arrays ``A[i, j]``
(``i`` = first index of the reference's ``mtrx``).  Every call runs on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Config


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def sor_beta(nx: int, ny: int) -> float:
    """SOR relaxation factor of the driver (src/main.c:134, truncated PI)."""
    return float(_lib.lib().cnv_sor_beta(nx, ny))


def num_steps(tf: float, dt: float) -> int:
    """Number of time steps the reference loop executes (src/main.c:162, :276)."""
    return int(_lib.lib().cnv_num_steps(tf, dt))


def diff_matrix(n: int, order: int, deriv: int, h: float) -> np.ndarray:
    """Dense Diff1 / Diff2 matrix (src/finitediff.c:51, :178).  Host only."""
    D = np.zeros((n, n))
    if _lib.lib().cnv_diff_dense(n, order, deriv, h, D):
        raise ValueError("** Error: valid orders are 2, 4 or 6 **")
    return D


def apply_operator(A, axis: int, deriv: int, order: int, h: float) -> np.ndarray:
    """DX (axis=1) / DY (axis=0) first or second derivative, matrix-free on the GPU."""
    _lib.require_gpu()
    A = _c(A)
    out = np.empty_like(A)
    if _lib.lib().cnv_apply_host(A, A.shape[0], A.shape[1], axis, deriv, order, h, out):
        raise ValueError("** Error: valid orders are 2, 4 or 6 **")
    return out


def euler(w, dwdx, dwdy, d2wdx2, d2wdy2, u, v, Re: float, dt: float) -> np.ndarray:
    """src/fluiddyn.c:71-102; returns the advanced vorticity (the reference updates ``w`` in place)."""
    _lib.require_gpu()
    w = np.array(w, dtype=np.float64, order="C")
    _lib.lib().cnv_euler_host(w, _c(dwdx), _c(dwdy), _c(d2wdx2), _c(d2wdy2), _c(u), _c(v), w.shape[0], w.shape[1], Re, dt)
    return w


def continuity(dudx, dvdy) -> np.ndarray:
    _lib.require_gpu()
    a = _c(dudx)
    out = np.empty_like(a)
    _lib.lib().cnv_continuity_host(a, _c(dvdy), a.shape[0], a.shape[1], out)
    return out


def vorticity(first, second) -> np.ndarray:
    """Returns ``second - first`` like the reference (src/fluiddyn.c:194)."""
    _lib.require_gpu()
    a = _c(first)
    out = np.empty_like(a)
    _lib.lib().cnv_vorticity_host(a, _c(second), a.shape[0], a.shape[1], out)
    return out


def pressure_rhs(u, v, order: int, dx: float, dy: float) -> np.ndarray:
    """f = dudx**2 + dvdy**2 + 2*dudy*dvdx, the right-hand side of the pressure recipe the reference leaves commented
    out (src/main.c:421-427); p = poisson(-f)."""
    _lib.require_gpu()
    a = _c(u)
    out = np.empty_like(a)
    if _lib.lib().cnv_pressure_rhs_host(a, _c(v), a.shape[0], a.shape[1], order, dx, dy, out):
        raise ValueError("valid orders are 2, 4 or 6 (and the grid must be at least that large)")
    return out


def error(a, b) -> float:
    """sum |a - b| (src/poisson.c:34-60)."""
    _lib.require_gpu()
    a = _c(a)
    return float(_lib.lib().cnv_error_host(a, _c(b), a.shape[0], a.shape[1]))


class PoissonNotConverged(RuntimeError):
    """The reference prints 'Error: maximum number of iterations achieved for Poisson equation.' and exits."""


def poisson_sor(f, dx: float, dy: float, itmax: int, tol: float, beta: float = 1.0, T: int = 0, history: bool = False,
                raise_on_itmax: bool = True) -> dict:
    """poisson_SOR_log / poisson_log semantics (src/poisson.c:176-285), red-black ordering.
    Returns dict(u, k, e[, history]); k is the reference's logged iteration number (sweeps-1)."""
    _lib.require_gpu()
    f = _c(f)
    u = np.empty_like(f)
    k, e = C.c_int(), C.c_double()
    hist = np.zeros(itmax) if history else None
    st = _lib.lib().cnv_poisson_host(f, f.shape[0], f.shape[1], dx, dy, itmax, tol, beta, T, u, C.byref(k), C.byref(e),
                                     hist.ctypes.data if history else None)
    if st and raise_on_itmax:
        raise PoissonNotConverged("Error: maximum number of iterations achieved for Poisson equation.")
    out = dict(u=u, k=k.value, e=e.value, status=st)
    if history:
        out["history"] = hist[:k.value + 1]
    return out


class PoissonSolver:
    """Device-resident solver object (cnv_poisson_*): reusable buffers, asynchronous passes."""

    def __init__(self, nrows: int, ncols: int, T: int = 0, slab=None, handle=None):
        self.h = None
        self.owned = handle is None
        _lib.require_gpu()
        self.L = _lib.lib()
        if handle is not None:            # view of a solver owned by a cnv_sim
            self.h = handle
        elif slab is None:
            self.h = self.L.cnv_poisson_create(nrows, ncols, T)
        else:
            grow0, gnrows, own_lo, own_hi = slab
            self.h = self.L.cnv_poisson_create_slab(nrows, ncols, T, grow0, gnrows, own_lo, own_hi)
        self.nrows, self.ncols = nrows, ncols
        self.refresh_plan()
        self.T = self.plan["T"]
        self.ld = self.L.cnv_poisson_ld(self.h)

    def refresh_plan(self):
        """(Re-)read the launch plan: which pass kernel runs can change when the multi-GPU exchange is set up."""
        info = (C.c_longlong * 18)()
        self.L.cnv_poisson_plan_info(self.h, info)
        keys = ("WS", "HX", "Wout", "Hout", "nstrips", "nchunks", "threads", "smem", "T", "pow2",
                "onchip", "oc_T", "oc_ntx", "oc_nty", "oc_NPX", "oc_NPY", "oc_OW", "oc_OH")
        self.plan = dict(zip(keys, list(info)))

    def close(self):
        if self.h and self.owned:
            self.L.cnv_poisson_destroy(self.h)
        self.h = None

    __del__ = close

    def set_consts(self, dx, dy, beta):
        self.L.cnv_poisson_set_consts(self.h, dx, dy, beta)
        info = (C.c_longlong * 18)()
        self.L.cnv_poisson_plan_info(self.h, info)
        self.plan["pow2"] = info[9]

    def upload(self, f, fsign=1.0, stream=None):
        self.L.cnv_poisson_upload(self.h, _c(f), fsign, stream)

    def solve(self, itmax, tol, stream=None):
        k, e, sw, ps, rb = C.c_int(), C.c_double(), C.c_int(), C.c_int(), C.c_int()
        st = self.L.cnv_poisson_solve(self.h, itmax, tol, stream, C.byref(k), C.byref(e), C.byref(sw), C.byref(ps), C.byref(rb))
        return dict(status=st, k=k.value, e=e.value, sweeps=sw.value, passes=ps.value, buf=rb.value)

    def reset(self, itmax, tol, stream=None):
        self.L.cnv_poisson_reset(self.h, itmax, tol, stream)

    def enqueue(self, npasses, stream=None):
        self.L.cnv_poisson_enqueue(self.h, npasses, stream)

    def state(self, stream=None):
        st, e = (C.c_int * 6)(), (C.c_double * 2)()
        self.L.cnv_poisson_state(self.h, stream, st, e)
        return dict(state=st[0], cur=st[1], sweeps=st[2], passes=st[3], k=st[4], redo=st[5], e=e[0], last_e=e[1])

    def download(self, which, stream=None):
        u = np.empty((self.nrows, self.ncols))
        self.L.cnv_poisson_download(self.h, which, u, stream)
        return u


class Simulation:
    """Device-resident time stepping: the loop body of the reference driver (src/main.c:283-395)."""

    def __init__(self, cfg, T: int = 0):
        self.h = None
        _lib.require_gpu()
        self.L = _lib.lib()
        if isinstance(cfg, dict):
            from .config import config_from_dict
            cfg = config_from_dict(cfg)
        self.cfg = cfg
        self.h = self.L.cnv_sim_create(C.byref(cfg), T)
        self.shape = (cfg.nx, cfg.ny)

    def close(self):
        if self.h:
            self.L.cnv_sim_destroy(self.h)
            self.h = None

    __del__ = close

    def step(self, nsteps=1, diagnostics=True):
        k = np.zeros(nsteps, dtype=np.int32)
        e = np.zeros(nsteps)
        cmax, cmin = np.zeros(nsteps), np.zeros(nsteps)
        failed = self.L.cnv_sim_step(self.h, nsteps, k.ctypes.data, e.ctypes.data,
                                     cmax.ctypes.data if diagnostics else None, cmin.ctypes.data if diagnostics else None)
        return dict(failed_step=failed, k=k, e=e, cont_max=cmax, cont_min=cmin)

    def pressure(self, itmax: int = 0, tol: float = 0.0):
        """Pressure of the current velocity field: p = poisson(-f) with the configured Poisson variant
        (itmax / tol default to the configuration's)."""
        p = np.empty(self.shape)
        k, e = C.c_int(), C.c_double()
        st = self.L.cnv_sim_pressure(self.h, itmax, tol, p.ctypes.data, C.byref(k), C.byref(e))
        if st < 0:
            raise RuntimeError("pressure is available on single-GPU simulations only")
        return dict(status=st, p=p, k=k.value, e=e.value)

    def fields(self):
        out = {n: np.empty(self.shape) for n in ("psi", "w", "u", "v")}
        self.L.cnv_sim_get_fields(self.h, *[out[n].ctypes.data for n in ("psi", "w", "u", "v")])
        return out

    def counters(self):
        c = (C.c_longlong * 3)()
        self.L.cnv_sim_counters(self.h, c)
        return dict(sweeps=c[0], passes=c[1], steps=c[2])


class _HostBlock:
    """A block of cnv_host_alloc memory (page-locked, on the current device's NUMA node, recycled through the library's pool)."""

    def __init__(self, nbytes):
        self.L = _lib.lib()
        self.ptr = self.L.cnv_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError(f"cnv_host_alloc({nbytes})")

    def __del__(self):
        try:
            self.L.cnv_host_free(self.ptr)
        except Exception:
            pass


def host_empty(shape, dtype=np.float64):
    """numpy array over the library's host allocator (csrc/capi.cu cnv_host_alloc): what allocm() of the drop-in library hands
    to main.c -- blocks of >= 1 MiB are page-locked, so the host-buffer entry points copy them with one DMA each way."""
    shape = tuple(int(x) for x in (shape if hasattr(shape, "__len__") else (shape,)))
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    blk = _HostBlock(max(nbytes, 8))
    buf = (C.c_char * max(nbytes, 8)).from_address(blk.ptr)
    buf._block = blk  # the array keeps the buffer alive, the buffer keeps the block alive
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
