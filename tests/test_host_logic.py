"""CPU tests of the host-side logic and of the kernel's schedule (through the CPU emulator that compiles
the kernel's own per-thread step function, tests/emul/stream_emul.cc).  No GPU needed."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


_EMUL = {}


def _load_emul():
    """The schedule emulator, compiled from the kernel's own header (csrc/poisson_stream.h)."""
    if "e" in _EMUL:
        return _EMUL["e"]
    so = os.path.join(ROOT, "tests", "emul", "libstream_emul.so")
    src = os.path.join(ROOT, "tests", "emul", "stream_emul.cc")
    hdrs = [os.path.join(ROOT, "fluid_dynamics1_b200", "csrc", h) for h in ("poisson_stream.h", "poisson_plan.h", "exact.h")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(x) for x in [src] + hdrs):
        subprocess.run(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", src, "-o", so], check=True)
    E = C.CDLL(so)
    _EMUL["e"] = E
    return E


@pytest.fixture(scope="module")
def emul():
    E = _load_emul()
    E.emul_pass.argtypes = [C.c_int] * 10 + [C.c_double] * 3 + [C.c_int, dp, dp, dp, C.c_int, dp]
    E.emul_plan.argtypes = [C.c_int] * 10 + [np.ctypeslib.ndpointer(dtype=np.int64)]
    E.emul_decide.argtypes = [ip, dp, dp, C.c_int, C.c_void_p]
    E.emul_pass_sweeps.argtypes = [C.c_int] * 4
    E.emul_onchip_solve.argtypes = [C.c_int] * 7 + [C.c_double] * 3 + [C.c_int, dp, dp, C.c_int, C.c_double, ip, dp, C.c_void_p,
                                                                        np.ctypeslib.ndpointer(dtype=np.int64)]
    E.emul_check_div.restype = C.c_long
    E.emul_check_div.argtypes = [C.c_double, dp, C.c_long]
    return E


def _emul_sweeps(E, port, n, m, T, npass, mode, ws=0, ch=0, dx=None, dy=None, seed=0):
    rng = np.random.default_rng(seed)
    dx = dx or 1.0 / n
    dy = dy or 1.0 / m
    beta = port.beta(n, m)
    f = rng.standard_normal((n, m))
    ld = (m + 15) // 16 * 16
    fp = np.zeros((n, ld))
    fp[:, :m] = f
    a, b = np.zeros((n, ld)), np.zeros((n, ld))
    got_norms = []
    for _ in range(npass):
        norms = np.zeros(T)
        assert E.emul_pass(T, n, m, ld, 0, n, 0, n, ws, ch, dx, dy, beta, mode, a, fp, b, T, norms) == 0
        a, b = b, a
        got_norms += list(norms)
    u, onorms = port.poisson_sweeps(f, dx, dy, T * npass, beta)
    return a[:, :m], u, np.array(got_norms), onorms


@pytest.mark.parametrize("T", [1, 2, 4, 6, 8])
@pytest.mark.parametrize("shape", [(64, 64), (40, 72), (100, 100), (33, 47)])
def test_stream_schedule_bitwise_vs_oracle(emul, port, shape, T):
    """T temporally blocked sweeps per pass == T plain red-black sweeps of the oracle, bit for bit:
    ring buffer, stage skew, register history, halos, strip/chunk decomposition, both arithmetic paths."""
    n, m = shape
    for mode in (0, 1):                               # 0: pow2 fast path where exact, 1: literal general sequence
        for ws, ch in ((0, 0), (64, 2), (96, 3)):
            if ws and ws - 2 * max(8, 2 * T) < 8:
                continue
            got, want, gn, on = _emul_sweeps(emul, port, n, m, T, 3, mode, ws, ch)
            assert got.tobytes() == want.tobytes(), (shape, T, mode, ws, ch)
            np.testing.assert_allclose(gn, on, rtol=1e-13)


# ---- the persistent on-chip kernel (csrc/poisson_onchip.h): its per-thread code, tile geometry, masks and lagged stop
# machine run as a whole solve on the CPU (tests/emul/stream_emul.cc::run_onchip) ----
def _onchip(E, port, n, m, T, ntx, nty, itmax, tol, mode=0, dx=None, dy=None, sms=148, seed=0, beta=None, u0=None):
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((n, m))
    dx, dy = dx or 1.0 / m, dy or 1.0 / n
    beta = beta or port.beta(n, m)
    ld = (m + 15) // 16 * 16
    fp = np.zeros((n, ld))
    fp[:, :m] = f
    bufs = np.zeros((3, n, ld))
    ints, dbls, plan = np.zeros(8, dtype=np.int32), np.zeros(4), np.zeros(8, dtype=np.int64)
    hist = np.zeros(max(itmax, 1))
    if E.emul_onchip_solve(n, m, ld, T, ntx, nty, sms, dx, dy, beta, mode, fp, bufs, itmax, tol, ints, dbls, hist.ctypes.data, plan) != 0:
        return None
    want = port.poisson(f, dx, dy, itmax, tol, beta, redblack=True, history=True)
    return dict(u=bufs[ints[1]][:, :m], ints=ints, dbls=dbls, plan=[int(x) for x in plan], want=want, hist=hist)


@pytest.mark.parametrize("T", [2, 4, 6, 8])
@pytest.mark.parametrize("shape", [(64, 64), (40, 72), (100, 100), (33, 47), (131, 90), (9, 200)])
def test_onchip_schedule_bitwise_vs_oracle(emul, port, shape, T):
    """Register-resident patches, published perimeters, 2T-cell halos reloaded after every pass, three rotating buffers:
    == plain red-black sweeps of the oracle, bit for bit; one tile and several tiles in x and y (odd and even tile origins),
    both arithmetic paths, converging solves (redo pass included) and itmax stops."""
    n, m = shape
    for mode, dx, dy in ((0, None, None), (1, 0.013, 0.02)):
        for ntx, nty, itmax, tol in ((0, 0, 5000, 1e-3), (1, 1, 5000, 1e-3), (2, 3, 3 * T + 1, 0.0), (3, 2, 2 * T, 0.0), (2, 2, 1, 0.0)):
            r = _onchip(emul, port, n, m, T, ntx, nty, itmax, tol, mode, dx, dy, seed=n + T)
            if r is None:
                continue                               # (no plan with that many tiles for this grid)
            w = r["want"]
            where = (shape, T, mode, ntx, nty, itmax, tol, r["plan"])
            assert int(r["ints"][0]) == (1 if w["status"] == 0 else 2), where
            assert int(r["ints"][5]) == w["k"] and int(r["ints"][2]) == w["k"] + 1, where
            assert r["u"].tobytes() == w["u"].tobytes(), where
            np.testing.assert_allclose(r["hist"][:w["k"] + 1], w["history"], rtol=1e-12, err_msg=str(where))
            assert abs(r["dbls"][1] - w["e"]) <= 1e-12 * max(w["e"], 1e-300), where


def test_onchip_every_stop_position(emul, port):
    """The converged sweep at every position inside a pass, for every T: the lagged decision either takes the pass' output
    (last sweep) or reloads the pass' input and recomputes exactly the converged number of sweeps."""
    n = 40
    f_seed = 3
    base = _onchip(emul, port, n, n, 4, 2, 2, 5000, 1e-9, seed=f_seed)
    hist = base["want"]["history"]
    ks = [k for k in range(1, len(hist)) if hist[k] < hist[:k].min()][:18]
    assert len(ks) >= 16 and len({k % 8 for k in ks}) == 8
    for T in (2, 4, 6, 8):
        for k in ks:
            tol = 0.5 * (hist[k] + hist[:k].min())
            r = _onchip(emul, port, n, n, T, 2, 2, 5000, tol, seed=f_seed)
            assert r["want"]["k"] == k
            assert int(r["ints"][0]) == 1 and int(r["ints"][5]) == k, (T, k)
            assert r["u"].tobytes() == r["want"]["u"].tobytes(), (T, k)


def test_onchip_planner_properties(emul, port):
    for n, m, sms in [(64, 64, 148), (128, 128, 148), (256, 256, 148), (1024, 1024, 148), (1024, 1024, 132), (512, 2048, 148), (33, 47, 148)]:
        r = _onchip(emul, port, n, m, 0, 0, 0, 1, 0.0, sms=sms)
        assert r is not None, (n, m)
        T, H, OW, OH, NPX, NPY, ntx, nty = r["plan"]
        assert H == 2 * T and T in (2, 4, 6, 8) and OW % 4 == 0 and OH % 8 == 0
        assert ntx * nty <= sms and NPX * NPY <= 352                 # compute threads; + one service warp per CTA
        assert 4 * NPX == OW + 2 * H and 8 * NPY >= OH + 2 * H
        assert ntx * OW >= m and nty * OH >= n
        assert (ntx == 1 or OW >= H) and (nty == 1 or OH >= H)
    assert _onchip(emul, port, 2048, 2048, 0, 0, 0, 1, 0.0) is None      # does not fit the register files: streaming kernel


@pytest.mark.parametrize("trim", [0, 8, 11])
def test_stream_schedule_slab_with_trimmed_boundary_chunks(emul, port, trim, monkeypatch):
    """A middle slab (neighbours on both sides): its first / last chunk is `trim` rows shorter than the others
    (PassGeom::trim_lo/hi).  One pass from the exact global state must reproduce the oracle's T sweeps on the owned
    rows, bit for bit, whatever the chunk heights."""
    monkeypatch.setenv("CNV_POISSON_TRIM", str(trim))
    T, gn, m = 4, 160, 72
    r0, r1, h = 32, 128, 2 * T
    rng = np.random.default_rng(trim)
    f = rng.standard_normal((gn, m))
    dx, dy, beta = 1.0 / m, 1.0 / gn, port.beta(gn, m)
    u0, _ = port.poisson_sweeps(f, dx, dy, 3, beta)              # a non-trivial global state
    want, _ = port.poisson_sweeps(f, dx, dy, T, beta, u=u0.copy())
    ld = (m + 15) // 16 * 16
    lo, hi = r0 - h, r1 + h
    a, b, fp = np.zeros((hi - lo, ld)), np.zeros((hi - lo, ld)), np.zeros((hi - lo, ld))
    a[:, :m] = u0[lo:hi]
    fp[:, :m] = f[lo:hi]
    norms = np.zeros(T)
    for chunks in (3, 4):
        b[:] = 0
        assert emul.emul_pass(T, hi - lo, m, ld, lo, gn, h, h + (r1 - r0), 0, chunks, dx, dy, beta, 0, a, fp, b, T, norms) == 0
        assert b[h:h + (r1 - r0), :m].tobytes() == want[r0:r1].tobytes(), (trim, chunks)


@pytest.mark.parametrize("own,cols,lower,upper", [(512, 4096, True, True), (512, 4096, False, True), (4096, 4096, True, True),
                                                  (4096, 4096, True, False), (2048, 16384, True, True),
                                                  (4096, 4096, False, False), (1024, 1024, False, False)])
def test_production_slab_plans_bitwise(emul, port, own, cols, lower, upper):
    """The shapes of the benchmark runs (single-GPU 4096^2 and 1024^2; slabs: 4096^2 over 8 GPUs, 4096 x 4096 per GPU,
    16384^2 over 8 GPUs) with the planner's own choice of strips / chunks / trimmed boundary chunks at T = 8: one pass from the exact global
    state reproduces 8 oracle sweeps on the owned rows, bit for bit."""
    T = 8
    h, pad = 2 * T, 6 * T
    hlo, hhi = (h if lower else 0), (h if upper else 0)
    nloc = own + hlo + hhi
    gn = nloc + (pad if lower else 0) + (pad if upper else 0)     # a "global" grid that extends beyond the halos
    g0 = pad if lower else 0
    rng = np.random.default_rng(own + cols)
    f = rng.standard_normal((gn, cols))
    dx = dy = 1.0 / cols
    beta = port.beta(cols, cols)
    u0, _ = port.poisson_sweeps(f, dx, dy, 2, beta)
    want, _ = port.poisson_sweeps(f, dx, dy, T, beta, u=u0.copy())
    ld = (cols + 15) // 16 * 16
    a, b, fp = np.zeros((nloc, ld)), np.zeros((nloc, ld)), np.zeros((nloc, ld))
    a[:, :cols] = u0[g0:g0 + nloc]
    fp[:, :cols] = f[g0:g0 + nloc]
    norms = np.zeros(T)
    assert emul.emul_pass(T, nloc, cols, ld, g0, gn, hlo, hlo + own, 0, 0, dx, dy, beta, 0, a, fp, b, T, norms) == 0
    assert b[hlo:hlo + own, :cols].tobytes() == want[g0 + hlo:g0 + hlo + own].tobytes()


@pytest.mark.parametrize("seed", [1, 2])
def test_schedules_fuzz_vs_oracle(emul, port, seed, monkeypatch):
    """Seeded random sweep over both pass kernels' schedules: grid shape, temporal depth, forced strip width / chunk
    count, trimmed boundary chunks, slabs with a lower and/or upper neighbour, both arithmetic paths and spacings,
    short passes -- the owned rows after one pass from an exact global state equal the oracle's sweeps, bit for bit."""
    rng = np.random.default_rng(seed)
    checked = 0
    for case in range(90):
        T = int(rng.choice([1, 2, 4, 6, 8]))
        own, cols = int(rng.integers(3, 150)), int(rng.integers(3, 150))
        lower, upper = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        if T == 1:
            lower = upper = False                       # slabs need T >= 2
        h, pad = 2 * T, 6 * T
        if lower or upper:
            own = max(own, 2 * h + 2)
        hlo, hhi = (h if lower else 0), (h if upper else 0)
        nloc = own + hlo + hhi
        gn = nloc + (pad if lower else 0) + (pad if upper else 0)
        g0 = pad if lower else 0
        monkeypatch.setenv("CNV_POISSON_TRIM", str(int(rng.choice([0, 3, 8, 32]))))
        ws, ch = int(rng.choice([0, 0, 64, 96, 160])), int(rng.choice([0, 0, 1, 2, 3, 5]))
        if ws and ws - 2 * max(8, 2 * T) < 8:
            ws = 0
        mode = int(rng.integers(0, 2))
        dx, dy = (1.0 / 64, 1.0 / 64) if rng.integers(0, 2) else (0.013, 0.02)
        nsw = int(rng.integers(1, T + 1))
        f = rng.standard_normal((gn, cols))
        beta = 1.0 + 0.9 * rng.random()
        u0, _ = port.poisson_sweeps(f, dx, dy, 2, beta)
        want, wn = port.poisson_sweeps(f, dx, dy, nsw, beta, u=u0.copy())
        ld = (cols + 15) // 16 * 16
        a, b, fp = np.zeros((nloc, ld)), np.zeros((nloc, ld)), np.zeros((nloc, ld))
        a[:, :cols] = u0[g0:g0 + nloc]
        fp[:, :cols] = f[g0:g0 + nloc]
        norms = np.zeros(8)
        assert emul.emul_pass(T, nloc, cols, ld, g0, gn, hlo, hlo + own, ws, ch, dx, dy, beta, mode, a, fp, b, nsw, norms) == 0
        where = dict(case=case, T=T, own=own, cols=cols, lower=lower, upper=upper, ws=ws, ch=ch, mode=mode, nsw=nsw)
        assert b[hlo:hlo + own, :cols].tobytes() == want[g0 + hlo:g0 + hlo + own].tobytes(), where
        if not (lower or upper):
            np.testing.assert_allclose(norms[:nsw], wn, rtol=1e-12, err_msg=str(where))
        checked += 1
    assert checked >= 70


def test_stream_schedule_nonuniform_spacing(emul, port):
    got, want, gn, on = _emul_sweeps(emul, port, 50, 38, 4, 2, 0, dx=0.013, dy=0.02)
    assert got.tobytes() == want.tobytes()


def test_planner_properties(emul):
    for (nrows, ncols) in [(64, 64), (128, 128), (1024, 1024), (4096, 4096), (2064, 16384), (512, 4096), (3, 3), (7, 1000)]:
        for T in (1, 2, 4, 6, 8):
            o = np.zeros(10, dtype=np.int64)
            emul.emul_plan(nrows, ncols, (ncols + 15) // 16 * 16, 0, nrows, 0, nrows, T, 0, 0, o)
            WS, HX, Wout, Hout, nstrips, nchunks, threads, smem, tlo, thi = (int(x) for x in o)
            assert WS > 0, (nrows, ncols, T)
            assert WS % 4 == 0 and HX % 4 == 0 and HX >= 2 * T and Wout == WS - 2 * HX
            assert nstrips * Wout >= ncols
            # chunks tile the rows: first = Hout - trim_lo, interior = Hout, the last one takes the remainder
            last = nrows - ((nchunks - 1) * Hout - tlo) if nchunks > 1 else nrows
            assert Hout - tlo >= 1 and 1 <= last <= Hout - thi + nchunks, (nrows, ncols, T, Hout, nchunks, tlo, thi)
            assert tlo == 0 and thi == 0                       # whole domain: no chunk next to a neighbour slab
            assert threads == T * WS // 4 and threads <= (640 if T >= 6 else 320)
            assert smem <= 227 * 1024 - 1024


def _ctl(itmax, tol):
    return np.array([0, 0, 0, 0, itmax, -1, 0], dtype=np.int32), np.array([tol, 0.0, 0.0])


def test_state_machine_matches_reference_loop(emul):
    """decide(): stop at the first sweep with e < tol, whatever its position inside a pass (redo when it
    is not the last sweep of the pass), itmax -> state 2; compared with the reference loop semantics
    (src/poisson.c:234-284) on random norm sequences."""
    rng = np.random.default_rng(0)
    for trial in range(300):
        T = int(rng.choice([1, 2, 4, 8]))
        itmax = int(rng.integers(1, 40))
        seq = rng.uniform(0.5, 2.0, size=itmax)
        if rng.random() < 0.8:
            seq[int(rng.integers(0, itmax))] = 0.1
        tol = 0.3
        want_k = next((k for k in range(itmax) if seq[k] < tol), None)
        ints, dbls = _ctl(itmax, tol)
        hist = np.zeros(itmax + 8)
        base, guard = 0, 0
        while ints[0] == 0:
            nsw = emul.emul_pass_sweeps(int(ints[2]), int(ints[3]), itmax, T)
            assert 1 <= nsw <= T
            base = int(ints[2])
            e = np.zeros(8)
            e[:nsw] = seq[base:base + nsw]
            cur_before = int(ints[1])
            emul.emul_decide(ints, dbls, e, nsw, hist.ctypes.data)
            if ints[3] > 0:                                   # redo requested: the pass input must stay current
                assert ints[1] == cur_before and ints[2] == base
            guard += 1
            assert guard < 100
        if want_k is None:
            assert ints[0] == 2 and ints[2] == itmax and ints[5] == itmax - 1
            assert dbls[2] == seq[itmax - 1]
        else:
            assert ints[0] == 1 and ints[5] == want_k and ints[2] == want_k + 1
            assert dbls[1] == seq[want_k]
            np.testing.assert_array_equal(hist[:want_k + 1], seq[:want_k + 1])


def test_constant_divisor_division_is_correctly_rounded(emul):
    """exact.h xdiv_const (Markstein correction with rd = RN(1/d)) vs the hardware division."""
    rng = np.random.default_rng(1)
    a = np.concatenate([rng.standard_normal(200000) * 10.0 ** rng.integers(-30, 30, 200000),
                        np.ldexp(rng.integers(1, 2 ** 53, 100000).astype(np.float64), -60)])
    for d in (2 * (0.01 ** 2 + 0.013 ** 2), 2 * (1 / 100.0 ** 2 + 1 / 100.0 ** 2), 3.0, 1e-7, 0.1, 7.0 / 3.0, 2 * (1.0 / 257 ** 2 * 2)):
        assert emul.emul_check_div(d, np.ascontiguousarray(a), a.size) == 0, d


# ---- the C ABI library: loads on CPU and exports every declared symbol -----------------------------
def _declared(header, pattern):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(pattern, txt))


def test_cabi_exports_every_declared_symbol():
    import fluid_dynamics1_b200 as fd
    from fluid_dynamics1_b200 import _lib
    L = fd.lib()           # binding resolves every name in CNV_API (AttributeError otherwise)
    D = fd.dropin()
    declared = _declared("cnavier_b200.h", r"\b(cnv_[a-z0-9_]+)\s*\(")
    assert declared == set(_lib.CNV_API), declared ^ set(_lib.CNV_API)
    for name in declared:
        assert hasattr(L, name)
    dropin_decl = _declared("cnavier_dropin.h", r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;")
    dropin_decl -= {"defined"}
    assert dropin_decl == set(_lib.DROPIN_API), dropin_decl ^ set(_lib.DROPIN_API)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.DROPIN_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert dropin_decl <= exported, dropin_decl - exported
    assert L.cnv_device_count() >= 0 and b"sm_100a" in L.cnv_version()


def test_cabi_scalars_and_coefficients_on_cpu(port):
    import fluid_dynamics1_b200 as fd
    assert fd.sor_beta(64, 64) == port.beta(64, 64)
    assert fd.num_steps(0.58, 0.005) == 115 and fd.num_steps(20.0, 0.005) == 4000
    for order in (2, 4, 6):
        for deriv in (1, 2):
            for n, h in ((7, 1 / 7), (16, 0.37), (40, 1 / 40)):
                assert np.array_equal(fd.diff_matrix(n, order, deriv, h), port.diff_dense(n, order, deriv, h))
    with pytest.raises(ValueError):
        fd.diff_matrix(8, 5, 1, 0.1)


def test_config_system_matches_reference(tmp_path, ref):
    """Same parse results as the reference's config.c on its shipped files and on the documented quirks."""
    import fluid_dynamics1_b200 as fd
    from fluid_dynamics1_b200._lib import Config
    cases = {
        "quirks.txt": "# comment\n; other comment\n\nRe = 400.5   # trailing comment dropped\nnx=33\n ny = 35\npoisson_tol = 5E-4\n"
                      "bogus_key = 3\n   # not a comment: first char is a space\nu4 = 2.5\norder = 4\nv1 = -0.25\ntf = 0.58\n",
        "default_like.txt": "Re = 1000.0\nLx = 1\nLy = 1\nnx = 64\nny = 64\ndt = 0.005\ntf = 20.0\nmax_co = 1.0\norder = 6\n"
                            "poisson_max_it = 10000\npoisson_tol = 1E-3\npoisson_type = 2\nopenmp_enabled = 1\noutput_interval = 10\n"
                            "u4 = 1.0\n",
        "highre_like.txt": "Re = 5000.0\nnx = 128\nny = 128\ndt = 0.001\ntf = 10.0\nmax_co = 0.5\npoisson_max_it = 15000\n"
                           "poisson_tol = 5E-4\noutput_interval = 5\n",
    }
    assert C.sizeof(Config) == ref.L.ref_config_size()
    for name, text in cases.items():
        p = tmp_path / name
        p.write_text(text)
        mine = fd.load_config_from_file(str(p))
        theirs = Config()
        ref.L.ref_load_config_from_file(str(p).encode(), C.byref(theirs))
        a, b = mine.as_dict(), theirs.as_dict()
        a.pop("openmp_enabled"); b.pop("openmp_enabled")          # default differs by build flag only
        assert a == b, name
    missing = fd.load_config_from_file(str(tmp_path / "does_not_exist.txt"))       # unreadable file -> defaults
    d = fd.load_default_config().as_dict()
    assert missing.as_dict() == d and d["nx"] == 64 and d["u4"] == 1.0 and d["poisson_tol"] == 1e-3


def test_no_oracle_or_cpu_fallback_in_product():
    """The product path never imports / links the oracle and has no CPU compute fallback."""
    pkg = os.path.join(ROOT, "fluid_dynamics1_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cc", ".h", ".cuh", "Makefile")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in txt.lower() or fn == "Makefile" and "oracle" not in txt, (dirpath, fn)
    out = subprocess.run(["ldd", os.path.join(pkg, "libcnavier_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "cnavier_ref" not in out
    import fluid_dynamics1_b200 as fd
    if fd.lib().cnv_device_count() == 0:
        with pytest.raises(RuntimeError):
            fd.poisson_sor(np.zeros((8, 8)), 0.1, 0.1, 10, 1e-3)
        with pytest.raises(RuntimeError):
            fd.Simulation(fd.load_default_config())


def test_dropin_host_linear_algebra_vs_reference(ref):
    """The host-side part of the drop-in library (no GPU involved): the reference's `mtrx` container and dense helpers that
    an unmodified main.c still calls (src/linearalg.c: initm/eye/reshape/kronecker/mtrxmul/mtrxcpy/invsig/maxel/minel/freem)
    and Diff1/Diff2 (src/finitediff.c:51, :178), called through the reference's own by-value `mtrx` ABI in BOTH libraries
    (libcnavier_dropin.so vs the compiled reference oracle/_ref/libcnavier_ref_ser.so) and compared bit for bit."""
    import fluid_dynamics1_b200 as fd
    from fluid_dynamics1_b200 import _lib
    D = fd.dropin()
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcnavier_ref_ser.so"), mode=C.RTLD_LOCAL)
    names = ("initm", "eye", "reshape", "kronecker", "mtrxmul", "mtrxcpy", "invsig", "maxel", "minel", "freem", "Diff1", "Diff2")
    for name in names:
        res, args = _lib.DROPIN_API[name]
        fn = getattr(R, name)
        fn.restype, fn.argtypes = res, args

    def fill(L, a):
        m = L.initm(a.shape[0], a.shape[1])
        for i in range(a.shape[0]):
            for j in range(a.shape[1]):
                m.M[i][j] = a[i, j]
        return m

    def read(m):
        return np.array([[m.M[i][j] for j in range(m.n)] for i in range(m.m)])

    rng = np.random.default_rng(3)
    a, b = rng.standard_normal((5, 4)), rng.standard_normal((4, 6))
    for L in (D, R):
        assert read(L.initm(3, 2)).tobytes() == np.zeros((3, 2)).tobytes()
    got, want = [], []
    for L, out in ((D, got), (R, want)):
        A, B = fill(L, a), fill(L, b)
        out.append(read(L.eye(4)))
        out.append(read(L.mtrxmul(A, B)))                  # ascending-k accumulation (src/linearalg.c:254-265)
        S1, S2 = fill(L, a[:4, :4]), fill(L, b[:4, :4])
        out.append(read(L.kronecker(S1, S2)))              # (the reference's index arithmetic only holds for equal square sizes, Q17)
        out.append(read(L.reshape(A, 4, 5)))
        out.append(read(L.reshape(A, 10, 2)))
        out.append(np.array([L.maxel(A), L.minel(A)]))
        Cc = L.initm(5, 4)
        L.mtrxcpy(Cc, A)
        L.invsig(Cc)
        out.append(read(Cc))
        for order in (2, 4, 6):
            for n in (7, 12):
                out.append(read(L.Diff1(n, order, 0.125)))
                out.append(read(L.Diff2(n, order, 0.3)))
        for m in (A, B, Cc):
            L.freem(m)                                     # same ownership convention: one malloc per row + the row table
    assert len(got) == len(want)
    for k, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape and g.tobytes() == w.tobytes(), k
    # the dense operator route an unmodified main.c takes: DX = kron(I, Diff1), DX * vec(A) -- identical in both libraries
    n = 6
    x = rng.standard_normal((n, n))
    res = []
    for L in (D, R):
        DX = L.kronecker(L.eye(n), L.Diff1(n, 4, 1.0 / n))
        v = L.reshape(fill(L, x), n * n, 1)
        res.append(read(L.reshape(L.mtrxmul(DX, v), n, n)))
    assert res[0].tobytes() == res[1].tobytes()


def test_vtk_writer_vs_reference_printvtk(tmp_path):
    """cnv_vtk_write (the driver's writer) against the reference's own printvtk (src/utils.c:38-100, compiled
    into oracle/_ref): same file names (one counter across all fields: stream-function-1-0, vorticity-1-1, ...), same
    bytes, append mode.  Host code only."""
    import fluid_dynamics1_b200 as fd
    from fluid_dynamics1_b200 import _lib
    L = fd.lib()
    so = os.path.join(ROOT, "oracle", "_ref", "libcnavier_ref_ser.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built")
    R = C.CDLL(so, mode=C.RTLD_LOCAL)
    R.initm.restype, R.initm.argtypes = _lib.Mtrx, [C.c_int, C.c_int]
    R.printvtk.restype, R.printvtk.argtypes = None, [_lib.Mtrx, C.c_char_p, C.c_char_p]
    rng = np.random.default_rng(8)
    mine, theirs = tmp_path / "mine", tmp_path / "theirs"
    fields = [("stream-function", rng.standard_normal((5, 7)) * 1e-3), ("vorticity", rng.standard_normal((5, 7)) * 40),
              ("x-velocity", np.array([[0.0, -0.0, 1.0, 0.9999995, -1e-7, 123456.789, 5e-7]] * 5)), ("y-velocity", np.zeros((5, 7)))]
    for rep in range(2):                                     # two output steps: the counter keeps counting (0..7)
        for title, a in fields:
            a = np.ascontiguousarray(a + rep)
            used = L.cnv_vtk_write(a, a.shape[0], a.shape[1], title.encode(), str(mine).encode())
            m = R.initm(a.shape[0], a.shape[1])
            for i in range(a.shape[0]):
                for j in range(a.shape[1]):
                    m.M[i][j] = a[i, j]
            R.printvtk(m, title.encode(), str(theirs).encode())
            assert used == rep * 4 + [t for t, _ in fields].index(title)
    a_files, b_files = sorted(os.listdir(mine)), sorted(os.listdir(theirs))
    assert a_files == b_files and len(a_files) == 8
    for name in a_files:
        assert (mine / name).read_bytes() == (theirs / name).read_bytes(), name


def test_config_printing_matches_reference(tmp_path, ref):
    """print_config (src/config.c:230): the drop-in library prints what the reference prints, character for character, for the
    default configuration and for a parsed file -- except the two OpenMP lines, which describe the build (both libraries run
    in a child process so that their C stdio output can be captured).  print_usage: same synopsis line; the help text below
    it is this repo's own wording."""
    import sys
    cfgfile = tmp_path / "c.txt"
    cfgfile.write_text("Re = 400.5\nnx = 33\nny = 35\npoisson_tol = 5E-4\nu4 = 2.5\norder = 4\nv1 = -0.25\ntf = 0.58\noutput_interval = 7\n")
    code = r"""
import ctypes as C, sys
sys.path.insert(0, {root!r})
from fluid_dynamics1_b200._lib import Config
which, path = sys.argv[1], sys.argv[2]
if which == "mine":
    import fluid_dynamics1_b200 as fd
    L = fd.dropin()
else:
    L = C.CDLL({ref!r}, mode=C.RTLD_LOCAL)
    L.load_default_config.restype = Config
    L.load_config_from_file.restype, L.load_config_from_file.argtypes = Config, [C.c_char_p]
    L.print_config.argtypes = [C.POINTER(Config)]
    L.print_usage.argtypes = [C.c_char_p]
for cfg in (L.load_default_config(), L.load_config_from_file(path.encode())):
    L.print_config(C.byref(cfg))
L.print_usage(b"./cnavier")
C.CDLL(None).fflush(None)
""".format(root=ROOT, ref=os.path.join(ROOT, "oracle", "_ref", "libcnavier_ref_ser.so"))
    outs = {}
    for which in ("mine", "ref"):
        r = subprocess.run([sys.executable, "-c", code, which, str(cfgfile)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[which] = r.stdout
    assert "Re" in outs["ref"] and len(outs["ref"]) > 200

    def config_part(text):
        lines = text.splitlines()
        cut = next(i for i, ln in enumerate(lines) if ln.startswith("Usage:"))
        return [ln for ln in lines[:cut] if "OpenMP" not in ln], lines[cut]
    mine, ref_ = config_part(outs["mine"]), config_part(outs["ref"])
    assert mine[0] == ref_[0] and len(mine[0]) > 30
    assert mine[1] == ref_[1] == "Usage: ./cnavier [config_file] [output_folder]"


def test_driver_courant_check_matches_reference_executable(tmp_path):
    """The driver's stability check (src/main.c:165-174: Courant numbers from u1, message + exit status 1) happens before
    any GPU work, so the executable can be compared with the reference executable on the CPU: same three lines, same
    exit status; a stable configuration passes the check in both."""
    exe = os.path.join(ROOT, "fluid_dynamics1_b200", "cnavier_b200")
    refexe = os.path.join(ROOT, "oracle", "_ref", "cnavier_ser")
    if not (os.path.exists(exe) and os.path.exists(refexe)):
        pytest.skip("executables not built")
    cfg = tmp_path / "unstable.txt"
    cfg.write_text("nx = 8\nny = 8\ndt = 0.1\ntf = 0.35\nu1 = 100.0\nmax_co = 1.0\n")
    outs = []
    for binary in (exe, refexe):
        r = subprocess.run([binary, str(cfg), "t"], capture_output=True, text=True, cwd=tmp_path, timeout=120)
        lines = r.stdout.splitlines()
        assert "Unstable Solution!" in lines, r.stdout[-1500:]
        i = lines.index("Unstable Solution!")
        outs.append((r.returncode, lines[i:i + 3]))
    assert outs[0] == outs[1] and outs[0][0] == 1
    assert outs[0][1][1].startswith("r1: 80.0") and outs[0][1][2].startswith("r2: 80.0")


def test_dropin_bad_order_exits_like_reference():
    """Diff1 / Diff2 with an order other than 2, 4, 6 (src/finitediff.c:150-151, 289-290), mtrxmul / reshape with
    mismatching shapes (src/linearalg.c:241-247, 378-384): message and exit(1) -- the drop-in library and the compiled
    reference print the same text and end with the same status (each in a child process, since both end the process)."""
    import sys
    code = r"""
import ctypes as C, sys
sys.path.insert(0, {root!r})
from fluid_dynamics1_b200 import _lib
which, fn = sys.argv[1], sys.argv[2]
if which == "mine":
    import fluid_dynamics1_b200 as fd
    L = fd.dropin()
else:
    L = C.CDLL({ref!r}, mode=C.RTLD_LOCAL)
    for name in ("Diff1", "Diff2"):
        getattr(L, name).restype, getattr(L, name).argtypes = _lib.DROPIN_API[name]
for name in ("initm", "mtrxmul", "reshape"):
    getattr(L, name).restype, getattr(L, name).argtypes = _lib.DROPIN_API[name]
if fn in ("Diff1", "Diff2"):
    getattr(L, fn)(8, 3, 0.125)
elif fn == "mtrxmul":
    L.mtrxmul(L.initm(2, 3), L.initm(2, 3))      # inner dimensions differ (src/linearalg.c:241-247)
else:
    L.reshape(L.initm(2, 3), 4, 2)               # element counts differ (src/linearalg.c:378-384)
print("returned")
""".format(root=ROOT, ref=os.path.join(ROOT, "oracle", "_ref", "libcnavier_ref_ser.so"))
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libcnavier_ref_ser.so")):
        pytest.skip("oracle/_ref not built")
    for fn in ("Diff1", "Diff2", "mtrxmul", "reshape"):
        res = []
        for which in ("mine", "ref"):
            r = subprocess.run([sys.executable, "-c", code, which, fn], capture_output=True, text=True, timeout=120)
            res.append((r.returncode, r.stdout.strip()))
        assert res[0] == res[1], res
        assert res[0][0] == 1 and "returned" not in res[0][1] and "rror" in res[0][1]


def test_host_allocator_semantics_without_a_device():
    """cnv_host_alloc / cnv_host_free / fd.host_empty are storage, not compute: without a GPU they hand out plain heap memory
    (not page-locked), any size, freed exactly once when the last array over a block dies."""
    import gc

    import fluid_dynamics1_b200 as fd

    L = fd.lib()
    a = fd.host_empty((300, 200))
    a[...] = 3.0
    big = fd.host_empty((1024, 513))
    big[...] = 1.0
    assert a.shape == (300, 200) and a.flags.c_contiguous and a.sum() == 180000.0 and big.sum() == 1024 * 513
    assert L.cnv_host_is_pinned(a.ctypes.data) == 0 and L.cnv_host_is_pinned(big.ctypes.data) == 0
    view = big[10:20]          # a view keeps the block alive
    del big
    gc.collect()
    assert view.sum() == 10 * 513
    p = L.cnv_host_alloc(0)    # zero bytes: still a valid, freeable block
    assert p
    L.cnv_host_free(p)
    L.cnv_host_free(None)
