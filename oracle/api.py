"""ctypes front-end of the parity oracle.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (``cpu_baseline`` leg and
``--impl reference``) may import this module.  The product package
``fluid_dynamics1_b200`` never does (tests/test_no_oracle_in_product.py enforces it).

Two back-ends:

* ``port()``  -- ``oracle/liboracle.so``: our C restatement (``cnavier_oracle.c``).
* ``ref()``   -- ``oracle/_ref/libcnavier_ref.so``: the UNMODIFIED reference sources
  (compiled by ``oracle/Makefile`` from ``/root/reference/src``) behind flat-array shims;
  ``ref(serial=True)`` is the same built without ``-DOPENMP_ENABLED`` (lexicographic SOR).
  Returns ``None`` when the prebuilt files are absent.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-C", HERE, "liboracle.so"] + (["ref"] if ref else []), check=True,
                   stdout=subprocess.DEVNULL)


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("Re", "dt", "dx", "dy", "beta", "poisson_tol")] + \
               [(n, C.c_int) for n in ("nx", "ny", "order", "poisson_max_it", "poisson_type", "redblack")] + \
               [(n, C.c_double) for n in ("u1", "u2", "u3", "u4", "v1", "v2", "v3", "v4")]


class Port:
    """The C restatement (oracle/cnavier_oracle.c)."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.L = C.CDLL(path)
        L.orc_diff_row.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), _dp, C.POINTER(C.c_int)]
        L.orc_diff_dense.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, _dp]
        L.orc_apply.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp]
        L.orc_poisson.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                  C.c_int, C.c_int, _dp, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p]
        L.orc_poisson_sweeps.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, _dp, C.c_void_p]
        L.orc_euler.argtypes = [_dp] * 7 + [C.c_double, C.c_double, C.c_size_t]
        L.orc_continuity.argtypes = [_dp, _dp, _dp, C.c_size_t]
        L.orc_vorticity.argtypes = [_dp, _dp, _dp, C.c_size_t]
        L.orc_pressure_rhs.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _dp]
        L.orc_beta.restype = C.c_double
        L.orc_beta.argtypes = [C.c_int, C.c_int]
        L.orc_num_steps.argtypes = [C.c_double, C.c_double]
        L.orc_step.argtypes = [C.POINTER(OrcParams), _dp, _dp, _dp, _dp, C.POINTER(C.c_int), C.POINTER(C.c_double),
                               C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_run.argtypes = [C.POINTER(OrcParams), C.c_double, C.c_double, C.c_int, _dp, _dp, _dp, _dp, _ip, _dp]

    def max_threads(self) -> int:
        return int(self.L.orc_max_threads())

    def diff_row(self, n, order, deriv, h, i):
        s, c = C.c_int(), C.c_int()
        co = np.zeros(7)
        if self.L.orc_diff_row(n, order, deriv, h, i, C.byref(s), co, C.byref(c)):
            raise ValueError("valid orders are 2, 4 or 6")
        return s.value, co[:c.value].copy()

    def diff_dense(self, n, order, deriv, h):
        D = np.zeros((n, n))
        if self.L.orc_diff_dense(n, order, deriv, h, D):
            raise ValueError("valid orders are 2, 4 or 6")
        return D

    def apply(self, A, axis, deriv, order, h):
        A = np.ascontiguousarray(A, dtype=np.float64)
        out = np.empty_like(A)
        rc = self.L.orc_apply(A, A.shape[0], A.shape[1], axis, deriv, order, h, out)
        if rc:
            raise ValueError("orc_apply failed (%d)" % rc)
        return out

    def poisson(self, f, dx, dy, itmax, tol, beta=1.0, redblack=True, sor=True, history=False):
        f = np.ascontiguousarray(f, dtype=np.float64)
        u = np.zeros_like(f)
        k, e = C.c_int(), C.c_double()
        hist = np.zeros(itmax) if history else None
        st = self.L.orc_poisson(f, f.shape[0], f.shape[1], dx, dy, itmax, tol, beta, int(redblack), int(sor), u,
                                C.byref(k), C.byref(e), hist.ctypes.data if history else None)
        out = dict(u=u, k=k.value, e=e.value, status=st)
        if history:
            out["history"] = hist[:k.value + 1]
        return out

    def poisson_sweeps(self, f, dx, dy, nsweeps, beta, u=None):
        f = np.ascontiguousarray(f, dtype=np.float64)
        u = np.zeros_like(f) if u is None else np.ascontiguousarray(u, dtype=np.float64)
        norms = np.zeros(nsweeps)
        self.L.orc_poisson_sweeps(f, f.shape[0], f.shape[1], dx, dy, nsweeps, beta, u, norms.ctypes.data)
        return u, norms

    def euler(self, w, dwdx, dwdy, d2wdx2, d2wdy2, u, v, Re, dt):
        w = np.array(w, dtype=np.float64, order="C")
        self.L.orc_euler(w, *[np.ascontiguousarray(a, dtype=np.float64) for a in (dwdx, dwdy, d2wdx2, d2wdy2, u, v)],
                         Re, dt, w.size)
        return w

    def pressure_rhs(self, u, v, order, dx, dy):
        u = np.ascontiguousarray(u, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty_like(u)
        if self.L.orc_pressure_rhs(u, v, u.shape[0], u.shape[1], order, dx, dy, out):
            raise ValueError("orc_pressure_rhs failed")
        return out

    def beta(self, nx, ny):
        return float(self.L.orc_beta(nx, ny))

    def num_steps(self, tf, dt):
        return int(self.L.orc_num_steps(tf, dt))

    def params(self, cfg: dict, redblack=True) -> OrcParams:
        P = OrcParams()
        P.Re, P.dt = cfg["Re"], cfg["dt"]
        P.nx, P.ny, P.order = cfg["nx"], cfg["ny"], cfg["order"]
        P.dx, P.dy = float(cfg["Lx"]) / cfg["nx"], float(cfg["Ly"]) / cfg["ny"]
        P.beta = self.beta(cfg["nx"], cfg["ny"])
        P.poisson_tol, P.poisson_max_it, P.poisson_type = cfg["poisson_tol"], cfg["poisson_max_it"], cfg["poisson_type"]
        P.redblack = int(redblack)
        for k in ("u1", "u2", "u3", "u4", "v1", "v2", "v3", "v4"):
            setattr(P, k, cfg[k])
        return P

    def run(self, cfg: dict, nsteps: int, redblack=True):
        """nsteps steps of the reference time loop from its initial condition."""
        P = self.params(cfg, redblack)
        shp = (cfg["nx"], cfg["ny"])
        u, v, w, psi = (np.zeros(shp) for _ in range(4))
        ks, es = np.zeros(nsteps, dtype=np.int32), np.zeros(nsteps)
        fail = self.L.orc_run(C.byref(P), cfg["ui"], cfg["vi"], nsteps, u, v, w, psi, ks, es)
        return dict(u=u, v=v, w=w, psi=psi, k=ks, e=es, failed_step=fail)

    def run_diag(self, cfg: dict, nsteps: int, redblack=True):
        """run() one orc_step at a time, also returning the per-step continuity max / min (src/main.c:387-408)."""
        shp = (cfg["nx"], cfg["ny"])
        u, v, w, psi = (np.zeros(shp) for _ in range(4))
        u[1:-1, 1:-1] = cfg["ui"]
        v[1:-1, 1:-1] = cfg["vi"]
        out = dict(k=[], e=[], cont_max=[], cont_min=[], failed_step=0)
        for t in range(nsteps):
            r = self.step(cfg, u, v, w, psi, redblack)
            for key in ("k", "e", "cont_max", "cont_min"):
                out[key].append(r[key])
            if r["status"]:
                out["failed_step"] = t + 1
                break
        out.update(u=u, v=v, w=w, psi=psi)
        return out

    def step(self, cfg: dict, u, v, w, psi, redblack=True):
        P = self.params(cfg, redblack)
        k, e, mx, mn = C.c_int(), C.c_double(), C.c_double(), C.c_double()
        st = self.L.orc_step(C.byref(P), u, v, w, psi, C.byref(k), C.byref(e), C.byref(mx), C.byref(mn))
        return dict(status=st, k=k.value, e=e.value, cont_max=mx.value, cont_min=mn.value)


class Ref:
    """The unmodified reference functions (oracle/_ref/libcnavier_ref*.so)."""

    def __init__(self, path, openmp):
        L = self.L = C.CDLL(path)
        self.openmp = openmp
        L.ref_diff.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, _dp]
        L.ref_apply_dense.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp]
        L.ref_poisson.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int,
                                  _dp, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.ref_error.restype = C.c_double
        L.ref_error.argtypes = [_dp, _dp, C.c_int, C.c_int]
        L.ref_euler.argtypes = [_dp] * 7 + [C.c_int, C.c_int, C.c_double, C.c_double]
        L.ref_continuity.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp]
        L.ref_vorticity.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp]
        L.ref_load_default_config.argtypes = [C.c_void_p]
        L.ref_load_config_from_file.argtypes = [C.c_char_p, C.c_void_p]
        L.ref_set_openmp(1 if openmp else 0)

    def diff(self, n, order, deriv, h):
        D = np.zeros((n, n))
        self.L.ref_diff(n, order, deriv, h, D)
        return D

    def apply_dense(self, A, axis, deriv, order, h):
        A = np.ascontiguousarray(A, dtype=np.float64)
        assert A.shape[0] == A.shape[1], "the reference's Kronecker route only works on square grids"
        out = np.empty_like(A)
        self.L.ref_apply_dense(A, A.shape[0], axis, deriv, order, h, out)
        return out

    def poisson(self, f, dx, dy, itmax, tol, beta=1.0, sor=True):
        f = np.ascontiguousarray(f, dtype=np.float64)
        u = np.zeros_like(f)
        k, e = C.c_int(), C.c_double()
        st = self.L.ref_poisson(f, f.shape[0], f.shape[1], dx, dy, itmax, tol, beta, 2 if sor else 1, u, C.byref(k), C.byref(e))
        return dict(u=u, k=k.value, e=e.value, status=st)

    def error(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64)
        return float(self.L.ref_error(a, np.ascontiguousarray(b, dtype=np.float64), a.shape[0], a.shape[1]))

    def euler(self, w, dwdx, dwdy, d2wdx2, d2wdy2, u, v, Re, dt):
        w = np.array(w, dtype=np.float64, order="C")
        self.L.ref_euler(w, *[np.ascontiguousarray(a, dtype=np.float64) for a in (dwdx, dwdy, d2wdx2, d2wdy2, u, v)],
                         w.shape[0], w.shape[1], Re, dt)
        return w

    def continuity(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64)
        out = np.empty_like(a)
        self.L.ref_continuity(a, np.ascontiguousarray(b, dtype=np.float64), a.shape[0], a.shape[1], out)
        return out

    def vorticity(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64)
        out = np.empty_like(a)
        self.L.ref_vorticity(a, np.ascontiguousarray(b, dtype=np.float64), a.shape[0], a.shape[1], out)
        return out


_cache: dict = {}


def port() -> Port:
    if "port" not in _cache:
        _cache["port"] = Port()
    return _cache["port"]


def ref(serial: bool = False):
    key = "ref_ser" if serial else "ref"
    if key not in _cache:
        path = os.path.join(HERE, "_ref", "libcnavier_ref_ser.so" if serial else "libcnavier_ref.so")
        _cache[key] = Ref(path, openmp=not serial) if os.path.exists(path) else None
    return _cache[key]


def ref_binary(serial: bool = False):
    path = os.path.join(HERE, "_ref", "cnavier_ser" if serial else "cnavier_omp")
    return path if os.path.exists(path) else None


# ---- configs (the reference's shipped files, restated as dicts; config_default.txt:7-54,
# ---- config_high_re.txt:5-40) --------------------------------------------------------------
CONFIG_DEFAULT = dict(Re=1000.0, Lx=1, Ly=1, nx=64, ny=64, dt=0.005, tf=20.0, max_co=1.0, order=6,
                      poisson_max_it=10000, poisson_tol=1e-3, poisson_type=2, openmp_enabled=1, output_interval=10,
                      ui=0.0, vi=0.0, u1=0.0, u2=0.0, u3=0.0, u4=1.0, v1=0.0, v2=0.0, v3=0.0, v4=0.0)
CONFIG_HIGH_RE = dict(CONFIG_DEFAULT, Re=5000.0, nx=128, ny=128, dt=0.001, tf=10.0, max_co=0.5,
                      poisson_max_it=15000, poisson_tol=5e-4, output_interval=5)


def write_config(cfg: dict, path: str) -> None:
    with open(path, "w") as f:
        f.write("# generated by oracle/api.py\n")
        for k, v in cfg.items():
            f.write(f"{k} = {v!r}\n")


def run_reference_binary(cfg: dict, workdir: str, serial: bool = False, threads: int | None = None, timeout=3600):
    """Run the real reference executable on ``cfg``; returns (log_text, dumps) where dumps maps
    field name -> array [ndumps, nx, ny] of the raw fp64 fields at every printvtk call."""
    exe = ref_binary(serial)
    if exe is None:
        raise FileNotFoundError("oracle/_ref not built")
    os.makedirs(workdir, exist_ok=True)
    dump = os.path.join(workdir, "dump")
    os.makedirs(dump, exist_ok=True)
    for fn in os.listdir(dump):
        os.remove(os.path.join(dump, fn))
    cfgp = os.path.join(workdir, "cfg.txt")
    write_config(cfg, cfgp)
    env = dict(os.environ, CNAVIER_DUMP_DIR=dump, CNAVIER_DUMP_ONLY="1")
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    subprocess.run([exe, cfgp, "run"], cwd=workdir, env=env, check=True, stdout=subprocess.DEVNULL, timeout=timeout)
    with open(os.path.join(workdir, "output", "logs", "run.txt")) as f:
        log = f.read()
    names = {"stream-function": "psi", "vorticity": "w", "x-velocity": "u", "y-velocity": "v"}
    dumps = {}
    for title, key in names.items():
        p = os.path.join(dump, title + ".f64")
        if os.path.exists(p):
            dumps[key] = np.fromfile(p, dtype=np.float64).reshape(-1, cfg["nx"], cfg["ny"])
    return log, dumps


def parse_poisson_log(text: str):
    """[(k, residual_string)] from 'Poisson equation solved with k iterations - ... error: X' lines."""
    out = []
    for line in text.splitlines():
        if line.startswith("Poisson equation solved with"):
            parts = line.split()
            out.append((int(parts[4]), parts[-1]))
    return out


def parse_continuity_log(text: str):
    """[(max_string, min_string)] from the 'Continuity max: X | Continuity min: Y | ...' lines (src/main.c:400-408)."""
    out = []
    for line in text.splitlines():
        if line.startswith("Continuity max:"):
            parts = line.split()
            out.append((parts[2], parts[6]))
    return out
