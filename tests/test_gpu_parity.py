"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle
(oracle/cnavier_oracle.c, pinned in test_oracle_pinned.py) and the committed golden fixtures.

Bar: the whole path is fp64 with individually rounded operations in the reference's association
order, so fields are compared BITWISE wherever the summation order cannot differ; the only quantity
whose rounding may differ is the L1 convergence norm (a grid-wide sum), checked to 1e-12 relative.
The north_star tolerance (relative L2 <= 1e-8 on psi, w, u, v) is asserted as well.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, load_golden, rel_l2, sine_rhs

pytestmark = pytest.mark.gpu

fd = pytest.importorskip("fluid_dynamics1_b200")
from oracle import api  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    fd.require_gpu()


@pytest.fixture(params=["stream", "onchip"])
def small_grid_path(request, monkeypatch):
    """The two Poisson kernels a small grid can take: the streaming pass kernel (one launch per T sweeps) and the persistent
    on-chip kernel (register-resident tiles, the whole solve in one launch; the default wherever it has a plan)."""
    monkeypatch.setenv("CNV_POISSON_ONCHIP", "1" if request.param == "onchip" else "0")
    return request.param


# ---- stencils ---------------------------------------------------------------------------------
@pytest.mark.parametrize("order", [2, 4, 6])
@pytest.mark.parametrize("shape", [(7, 7), (16, 16), (33, 50), (64, 64), (200, 131)])
def test_apply_operator_bitwise(port, order, shape):
    rng = np.random.default_rng(order * 100 + shape[0])
    A = rng.standard_normal(shape) * 10.0 ** rng.integers(-3, 3, shape)
    for axis in (0, 1):
        for deriv in (1, 2):
            h = 1.0 / shape[axis]
            got = fd.apply_operator(A, axis, deriv, order, h)
            assert got.tobytes() == port.apply(A, axis, deriv, order, h).tobytes(), (axis, deriv)


def test_apply_operator_bad_order():
    with pytest.raises(ValueError):
        fd.apply_operator(np.zeros((8, 8)), 0, 1, 3, 0.1)


def test_diff_matrix_vs_oracle(port):
    for order in (2, 4, 6):
        for deriv in (1, 2):
            assert np.array_equal(fd.diff_matrix(12, order, deriv, 0.37), port.diff_dense(12, order, deriv, 0.37))


def test_pointwise_bitwise(port):
    rng = np.random.default_rng(3)
    a = [rng.standard_normal((37, 53)) for _ in range(7)]
    assert np.array_equal(fd.euler(*a, 1000.0, 0.005), port.euler(*a, 1000.0, 0.005))
    assert np.array_equal(fd.continuity(a[0], a[1]), a[0] + a[1])
    assert np.array_equal(fd.vorticity(a[0], a[1]), a[1] - a[0])           # second minus first (quirk Q18)
    e = fd.error(a[0], a[1])
    assert abs(e - np.abs(a[0] - a[1]).sum()) <= 1e-12 * e


@pytest.mark.parametrize("order", [2, 4, 6])
@pytest.mark.parametrize("shape", [(16, 16), (33, 50), (64, 64), (200, 131)])
def test_pressure_rhs_bitwise(port, order, shape):
    """f = dudx**2 + dvdy**2 + 2*dudy*dvdx (the commented recipe of src/main.c:421-427), bitwise vs the oracle."""
    rng = np.random.default_rng(order * 100 + shape[0])
    u, v = rng.standard_normal(shape), rng.standard_normal(shape)
    dx, dy = 1.0 / shape[1], 1.0 / shape[0]
    assert fd.pressure_rhs(u, v, order, dx, dy).tobytes() == port.pressure_rhs(u, v, order, dx, dy).tobytes()


def test_simulation_pressure_vs_oracle(port):
    """Pressure of the cavity flow after 20 steps of config_default: p = poisson(-f) with the configured SOR solve,
    sweep count and field bitwise vs the oracle; psi and the other fields are left untouched."""
    sim = fd.Simulation(dict(api.CONFIG_DEFAULT))
    sim.step(20)
    before = sim.fields()
    r = sim.pressure()
    after = sim.fields()
    n = 64
    f = port.pressure_rhs(before["u"], before["v"], 6, 1.0 / n, 1.0 / n)
    want = port.poisson(-f, 1.0 / n, 1.0 / n, 10000, 1e-3, port.beta(n, n), redblack=True, sor=True)
    assert r["status"] == 0 and r["k"] == want["k"]
    assert r["p"].tobytes() == want["u"].tobytes()
    assert "%E" % r["e"] == "%E" % want["e"]
    for name in before:
        assert before[name].tobytes() == after[name].tobytes(), name
    r2 = sim.pressure(itmax=5, tol=1e-30)
    assert r2["status"] == 1
    sim.close()


# ---- Poisson ------------------------------------------------------------------------------------
@pytest.mark.parametrize("T", [1, 2, 4, 8])
@pytest.mark.parametrize("n", [16, 64, 100, 257])
def test_poisson_sor_vs_oracle(port, n, T, small_grid_path):
    """Same sweep count, same field bits, same residual (to summation rounding) as the red-black
    reference build; n=64 is the pow2 fast path, 100 and 257 the general (Markstein division) path."""
    f = sine_rhs(n)
    dx = 1.0 / n
    beta = port.beta(n, n)
    got = fd.poisson_sor(f, dx, dx, 20000, 1e-3, beta, T=T, history=True)
    want = port.poisson(f, dx, dx, 20000, 1e-3, beta, redblack=True, history=True)
    assert got["k"] == want["k"]
    assert got["u"].tobytes() == want["u"].tobytes()
    assert abs(got["e"] - want["e"]) <= 1e-12 * want["e"]
    assert "%E" % got["e"] == "%E" % want["e"]
    np.testing.assert_allclose(got["history"], want["history"], rtol=1e-12)


def test_poisson_sine_golden_counts():
    import json
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "poisson_sine.json")) as fh:
        g = json.load(fh)
    for n in (64, 128, 256):
        r = fd.poisson_sor(sine_rhs(n), 1.0 / n, 1.0 / n, 20000, 1e-3, fd.sor_beta(n, n))
        assert r["k"] == g[str(n)]["rb_k"]
        assert "%.10f" % r["e"] == "%.10f" % g[str(n)]["rb_e"]


@pytest.mark.parametrize("shape,dx,dy", [((40, 72), 1 / 40, 1 / 72), ((96, 33), 0.01, 0.013), ((5, 9), 0.2, 0.1)])
def test_poisson_nonsquare_random_rhs(port, shape, dx, dy, small_grid_path):
    rng = np.random.default_rng(shape[0])
    f = rng.standard_normal(shape)
    beta = port.beta(*shape)
    for T in (1, 4):
        got = fd.poisson_sor(f, dx, dy, 50000, 1e-6, beta, T=T)
        want = port.poisson(f, dx, dy, 50000, 1e-6, beta, redblack=True)
        assert got["k"] == want["k"]
        assert got["u"].tobytes() == want["u"].tobytes()


def test_poisson_gauss_seidel_variant(port, small_grid_path):
    """poisson()/poisson_log(): beta == 1, no relaxation term (src/poisson.c:62-109, 176-222)."""
    f = sine_rhs(48)
    got = fd.poisson_sor(f, 1 / 48, 1 / 48, 50000, 1e-3, 1.0)
    want = port.poisson(f, 1 / 48, 1 / 48, 50000, 1e-3, 1.0, redblack=True, sor=False)
    assert got["k"] == want["k"]
    assert np.array_equal(got["u"], want["u"])       # == ignores the sign of zero


def test_poisson_itmax(port, small_grid_path):
    """itmax reached: status 1 (the drop-in layer turns it into message + exit(1))."""
    f = sine_rhs(32)
    for itmax in (1, 3, 4, 5, 9):
        r = fd.poisson_sor(f, 1 / 32, 1 / 32, itmax, 1e-12, fd.sor_beta(32, 32), raise_on_itmax=False)
        assert r["status"] == 1 and r["k"] == itmax - 1
        u, norms = port.poisson_sweeps(f, 1 / 32, 1 / 32, itmax, port.beta(32, 32))
        assert r["u"].tobytes() == u.tobytes()
    with pytest.raises(fd.PoissonNotConverged):
        fd.poisson_sor(f, 1 / 32, 1 / 32, 5, 1e-12, fd.sor_beta(32, 32))


def test_poisson_zero_rhs_and_tiny_grid(port, small_grid_path):
    r = fd.poisson_sor(np.zeros((8, 8)), 0.1, 0.1, 10, 1e-3, 1.5)
    assert r["k"] == 0 and r["e"] == 0.0 and not r["u"].any()
    f = np.ones((3, 3))
    got = fd.poisson_sor(f, 0.5, 0.5, 100, 1e-9, 1.2)
    want = port.poisson(f, 0.5, 0.5, 100, 1e-9, 1.2, redblack=True)
    assert got["k"] == want["k"] and np.array_equal(got["u"], want["u"])


def test_poisson_every_stop_position(port, small_grid_path):
    """The converged sweep can fall on any position inside a temporal block: sweep through
    tolerances so that sweeps-1 takes many consecutive values, for every T."""
    n = 40
    f = sine_rhs(n)
    beta = port.beta(n, n)
    full = port.poisson(f, 1 / n, 1 / n, 5000, 1e-9, beta, redblack=True, history=True)
    hist = full["history"]
    # the SOR update norm first grows, then decays: use stop positions on the decaying tail, where
    # hist[k] is the first value below the tolerance
    ks = [k for k in range(1, len(hist)) if hist[k] < hist[:k].min()][:18]
    assert len(ks) >= 16 and len({k % 8 for k in ks}) == 8
    for T in (2, 4, 8):
        for k in ks:
            tol = 0.5 * (hist[k] + hist[:k].min())
            got = fd.poisson_sor(f, 1 / n, 1 / n, 5000, tol, beta, T=T)
            want = port.poisson(f, 1 / n, 1 / n, 5000, tol, beta, redblack=True)
            assert want["k"] == k
            assert got["k"] == k, (T, k)
            assert got["u"].tobytes() == want["u"].tobytes()


@pytest.mark.parametrize("n,sweeps", [(1024, 24), (4096, 20)])
def test_poisson_full_size_fixed_sweeps(port, n, sweeps, monkeypatch):
    """BASELINE sizes: K sweeps of the 1024^2 / 4096^2 cavity grids, bitwise against the oracle
    (OpenMP red-black) and identical for every temporal block depth.  T = 8 at 4096^2 is exactly the plan bench.py
    times (streaming kernel); 20 sweeps = 3 passes, the last one partial (4 of 8 levels active)."""
    rng = np.random.default_rng(n)
    f = rng.standard_normal((n, n))
    beta = port.beta(n, n)
    want, norms = port.poisson_sweeps(f, 1 / n, 1 / n, sweeps, beta)
    monkeypatch.setenv("CNV_POISSON_ONCHIP", "0")     # the streaming pass kernel (1024^2 takes the on-chip kernel by default)
    for T in (8, 6, 4, 2, 1):
        s = fd.PoissonSolver(n, n, T)
        assert not s.plan["onchip"]
        s.set_consts(1 / n, 1 / n, beta)
        s.upload(f)
        r = s.solve(sweeps, 0.0)
        assert r["status"] == 1 and r["sweeps"] == sweeps
        got = s.download(r["buf"])
        assert got.tobytes() == want.tobytes(), T
        assert abs(r["e"] - norms[-1]) <= 1e-11 * norms[-1]
        s.close()


@pytest.mark.parametrize("T,stop", [(8, 18), (8, 15), (6, 13)])
def test_poisson_4096_stop_inside_a_pass_streaming_kernel(port, T, stop):
    """The plan bench.py times (4096^2, streaming kernel, T = 8; T = 6 as well) with a REAL stop decision: beta = 1.5
    makes the update norm decay monotonically, the tolerance sits between the norms of sweeps `stop`-1 and `stop`, so the
    solve runs two full passes, detects the hit inside the third (T = 8, stop 18: level 2 of 8) and recomputes it with
    exactly the converged number of sweeps ("redo" pass); stop = 15 is the last sweep of a pass (no redo).  Sweep count,
    logged residual string and the field bitwise vs the oracle's red-black solve."""
    n = 4096
    rng = np.random.default_rng(n)
    f = rng.standard_normal((n, n))
    beta = 1.5
    _, norms = port.poisson_sweeps(f, 1 / n, 1 / n, stop + 1, beta)
    assert norms[stop] < norms[:stop].min()
    tol = 0.5 * (norms[stop] + norms[stop - 1])
    want = port.poisson(f, 1 / n, 1 / n, 1000, tol, beta, redblack=True)
    assert want["k"] == stop
    s = fd.PoissonSolver(n, n, T)
    s.set_consts(1 / n, 1 / n, beta)
    s.upload(f)
    r = s.solve(1000, tol)
    assert r["status"] == 0 and r["k"] == stop and r["sweeps"] == stop + 1
    assert r["passes"] == (stop + 1 + T - 1) // T + (0 if (stop + 1) % T == 0 else 1)   # + the redo pass
    assert s.download(r["buf"]).tobytes() == want["u"].tobytes()
    assert "%E" % r["e"] == "%E" % want["e"]
    s.close()


@pytest.mark.parametrize("T,ntx,nty", [(0, 0, 0), (4, 12, 12), (4, 9, 16), (2, 16, 9), (2, 12, 12)])
def test_poisson_onchip_1024_fixed_sweeps_and_converged(port, monkeypatch, T, ntx, nty):
    """BASELINE config 3 grid on the persistent on-chip kernel: the planner's own tiling and pinned ones (T = 2 .. 8, different
    tile grids; deeper blocking does not fit 384 threads per tile at this size), 27 sweeps with tol = 0 (itmax stop inside a pass) bitwise against the oracle; then a converging solve with a
    real stop decision (beta = 1.5: monotone norms, stop inside the third pass -> the pass' input is reloaded and recomputed)."""
    n = 1024
    monkeypatch.setenv("CNV_POISSON_ONCHIP", "1")
    monkeypatch.setenv("CNV_ONCHIP_T", str(T)); monkeypatch.setenv("CNV_ONCHIP_NTX", str(ntx)); monkeypatch.setenv("CNV_ONCHIP_NTY", str(nty))
    rng = np.random.default_rng(n)
    f = rng.standard_normal((n, n))
    beta = port.beta(n, n)
    s = fd.PoissonSolver(n, n, 0)
    assert s.plan["onchip"] == 1 and (T == 0 or (s.plan["oc_T"], s.plan["oc_ntx"], s.plan["oc_nty"]) == (T, ntx, nty)), s.plan
    s.set_consts(1 / n, 1 / n, beta)
    s.upload(f)
    want, norms = port.poisson_sweeps(f, 1 / n, 1 / n, 27, beta)
    r = s.solve(27, 0.0)
    assert r["status"] == 1 and r["sweeps"] == 27
    assert s.download(r["buf"]).tobytes() == want.tobytes()
    assert abs(r["e"] - norms[-1]) <= 1e-11 * norms[-1]
    beta = 1.5
    s.set_consts(1 / n, 1 / n, beta)
    _, norms = port.poisson_sweeps(f, 1 / n, 1 / n, 20, beta)
    for stop in (18, 15, 9):
        assert norms[stop] < norms[:stop].min()
        tol = 0.5 * (norms[stop] + norms[stop - 1])
        w = port.poisson(f, 1 / n, 1 / n, 1000, tol, beta, redblack=True)
        s.upload(f)
        r = s.solve(1000, tol)
        assert r["status"] == 0 and r["k"] == stop == w["k"]
        assert s.download(r["buf"]).tobytes() == w["u"].tobytes()
        assert "%E" % r["e"] == "%E" % w["e"]
    s.close()


def test_poisson_linearity_property():
    """Size-independent property at 2048^2: the sweep operator is affine in f with a zero start, so
    K sweeps on 2f equal twice K sweeps on f exactly (scaling by 2 commutes with rounding)."""
    n, K = 2048, 16
    rng = np.random.default_rng(7)
    f = rng.standard_normal((n, n))
    s = fd.PoissonSolver(n, n, 4)
    s.set_consts(1 / n, 1 / n, fd.sor_beta(n, n))
    s.upload(f)
    a = s.download(s.solve(K, 0.0)["buf"])
    s.upload(2 * f)
    b = s.download(s.solve(K, 0.0)["buf"])
    assert np.array_equal(b, 2 * a)
    assert not np.any(a[0]) and not np.any(a[-1]) and not np.any(a[:, 0]) and not np.any(a[:, -1])   # ring stays 0
    s.close()


# ---- whole time steps -----------------------------------------------------------------------------
def test_default_config_vs_golden_fields_and_logs(golden_logs, small_grid_path):
    """config_default.txt: fields after steps 0, 10, 20 against raw dumps of the reference executable;
    Poisson log lines against the reference's shipped testRunOMP.txt."""
    g = load_golden("fields_default_rb.npz")
    sim = fd.Simulation(dict(api.CONFIG_DEFAULT))
    done = 0
    ks, es, cmax, cmin = [], [], [], []
    for idx, step in enumerate(g["dump_steps"]):
        r = sim.step(int(step) + 1 - done)
        done = int(step) + 1
        assert r["failed_step"] == 0
        ks += list(r["k"]); es += list(r["e"]); cmax += list(r["cont_max"]); cmin += list(r["cont_min"])
        f = sim.fields()
        for name in ("psi", "w", "u", "v"):
            assert rel_l2(f[name], g[name][idx]) <= 1e-8, (name, step)
            assert np.array_equal(f[name], g[name][idx]), (name, step, rel_l2(f[name], g[name][idx]))
    assert ks == golden_logs["testRunOMP"]["k"][:done]
    assert ["%E" % e for e in es] == golden_logs["testRunOMP"]["e"][:done]
    r = sim.step(352 - done)                        # the rest of the 352 shipped log lines
    ks += list(r["k"]); es += list(r["e"]); cmax += list(r["cont_max"]); cmin += list(r["cont_min"])
    assert ks == golden_logs["testRunOMP"]["k"]
    assert ["%E" % e for e in es] == golden_logs["testRunOMP"]["e"]
    # the continuity diagnostic (src/fluiddyn.c:126-154 + maxel / minel, src/main.c:387-408).  The values are
    # rounding noise (~1e-16), so their logged 7-digit strings match only if u, v and DX u + DY v are bit-identical:
    # all 351 'Continuity max / min' lines the reference shipped in testRunOMP.txt (the 352nd step's line is cut off)
    gl = golden_logs["testRunOMP"]
    assert len(gl["cont_max"]) == 351
    assert ["%E" % x for x in cmax[:351]] == gl["cont_max"]
    assert ["%E" % x for x in cmin[:351]] == gl["cont_min"]


def test_high_re_config_vs_golden(golden_logs, small_grid_path):
    g = load_golden("fields_highre_rb.npz")
    sim = fd.Simulation(dict(api.CONFIG_HIGH_RE))
    r = sim.step(6)
    f = sim.fields()
    for name in ("psi", "w", "u", "v"):
        assert np.array_equal(f[name], g[name][-1]), (name, rel_l2(f[name], g[name][-1]))
    r2 = sim.step(6)
    assert list(r["k"]) + list(r2["k"]) == golden_logs["testRunOMPHIGHRES"]["k"]
    assert ["%E" % e for e in list(r["e"]) + list(r2["e"])] == golden_logs["testRunOMPHIGHRES"]["e"]
    gl = golden_logs["testRunOMPHIGHRES"]        # 11 complete 'Continuity max / min' lines shipped
    assert ["%E" % x for x in (list(r["cont_max"]) + list(r2["cont_max"]))[:11]] == gl["cont_max"][:11]
    assert ["%E" % x for x in (list(r["cont_min"]) + list(r2["cont_min"]))[:11]] == gl["cont_min"][:11]


@pytest.mark.parametrize("order,ptype,n", [(2, 2, 48), (4, 2, 48), (6, 1, 40), (6, 2, 50)])
def test_steps_vs_oracle_other_orders_and_solver_types(port, order, ptype, n, small_grid_path):
    cfg = dict(api.CONFIG_DEFAULT, nx=n, ny=n, order=order, poisson_type=ptype, poisson_max_it=100000,
               u1=0.1, u2=-0.2, u3=0.3, v1=0.05, v2=-0.05, v3=0.02, v4=-0.01, ui=0.01, vi=-0.02, dt=0.002)
    sim = fd.Simulation(cfg)
    r = sim.step(4)
    want = port.run(cfg, 4, redblack=True)
    assert list(r["k"]) == list(want["k"])
    f = sim.fields()
    for name in ("psi", "w", "u", "v"):
        assert np.array_equal(f[name], want[name]), (name, rel_l2(f[name], want[name]))


SMALL = dict(api.CONFIG_DEFAULT, nx=32, ny=32, dt=0.004, u1=0.05, u2=-0.03, v3=0.02, v4=0.01, output_interval=1)


@pytest.mark.parametrize("order", [2, 4])
def test_other_orders_vs_reference_executable(order, small_grid_path):
    """Finite-difference order 2 / 4: fields after each of 4 steps and the Poisson iteration counts equal the raw dumps of
    the reference executable (tests/golden/fields_order{2,4}_rb.npz, OpenMP build), bit for bit."""
    g = load_golden(f"fields_order{order}_rb.npz")
    sim = fd.Simulation(dict(SMALL, order=order))
    ks = []
    for idx in range(4):
        r = sim.step(1)
        ks.append(int(r["k"][0]))
        f = sim.fields()
        for name in ("psi", "w", "u", "v"):
            assert np.array_equal(f[name], g[name][idx]), (name, idx, rel_l2(f[name], g[name][idx]))
    assert ks == list(g["k"])


def test_gauss_seidel_vs_lexicographic_reference(small_grid_path):
    """poisson_type = 1.  The reference's Gauss-Seidel sweeps lexicographically in both of its builds; the GPU path
    sweeps red-black.  Compared the way the north_star prescribes for differing orderings: both converged to the same
    tight tolerance (1e-11), fields within 1e-8 relative L2 of the serial reference executable's dumps."""
    g = load_golden("fields_gs_lex_tight.npz")
    sim = fd.Simulation(dict(SMALL, order=4, poisson_type=1, poisson_tol=1e-11, poisson_max_it=200000))
    r = sim.step(3)
    assert r["failed_step"] == 0
    f = sim.fields()
    for name in ("psi", "w", "u", "v"):
        assert rel_l2(f[name], g[name][-1]) <= 1e-8, (name, rel_l2(f[name], g[name][-1]))


def test_tight_tolerance_vs_lexicographic_reference():
    """Ordering-independent check: at poisson_tol = 1e-11 the red-black GPU fields agree with the
    SERIAL (lexicographic) reference executable within the north_star tolerance."""
    g = load_golden("fields_default_lex_tight.npz")
    sim = fd.Simulation(dict(api.CONFIG_DEFAULT, poisson_tol=1e-11, poisson_max_it=100000))
    sim.step(3)
    f = sim.fields()
    for name in ("psi", "w", "u", "v"):
        assert rel_l2(f[name], g[name][-1]) <= 1e-8, name


def test_1024_steps_vs_oracle(port):
    """BASELINE config 3 grid (1024^2, Re 1000): two full time steps against the matrix-free oracle."""
    cfg = dict(api.CONFIG_DEFAULT, nx=1024, ny=1024, dt=1e-4, poisson_max_it=20000)
    sim = fd.Simulation(cfg)
    r = sim.step(2)
    want = port.run(cfg, 2, redblack=True)
    assert list(r["k"]) == list(want["k"])
    f = sim.fields()
    for name in ("psi", "w", "u", "v"):
        assert rel_l2(f[name], want[name]) <= 1e-8, name
        assert np.array_equal(f[name], want[name]), name


# ---- the reference's own C signatures (drop-in library) ---------------------------------------------
def _to_mtrx(D, a):
    from fluid_dynamics1_b200._lib import Mtrx
    m = D.initm(a.shape[0], a.shape[1])
    for i in range(a.shape[0]):
        C.memmove(m.M[i], a[i].ctypes.data, a.shape[1] * 8)
    return m


def _from_mtrx(m):
    out = np.empty((m.m, m.n))
    for i in range(m.m):
        C.memmove(out[i].ctypes.data, m.M[i], m.n * 8)
    return out


def test_dropin_signatures(port, tmp_path):
    D = fd.dropin()
    n = 64
    f = sine_rhs(n)
    F = _to_mtrx(D, f)
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    logp = str(tmp_path / "log.txt").encode()
    fh = libc.fopen(logp, b"w")
    U = D.poisson_SOR_log(F, 1 / n, 1 / n, 20000, 1e-3, port.beta(n, n), fh)
    libc.fclose(fh)
    want = port.poisson(f, 1 / n, 1 / n, 20000, 1e-3, port.beta(n, n), redblack=True)
    assert np.array_equal(_from_mtrx(U), want["u"])
    line = open(logp).read()
    assert line == "Poisson equation solved with %d iterations - root-sum-of-squares error: %E\n" % (want["k"], want["e"])
    assert abs(D.error(U, F) - np.abs(want["u"] - f).sum()) < 1e-9
    d1 = D.Diff1(10, 6, 0.1)
    assert np.array_equal(_from_mtrx(d1), port.diff_dense(10, 6, 1, 0.1))
    rng = np.random.default_rng(0)
    arrs = [rng.standard_normal((12, 12)) for _ in range(7)]
    ms = [_to_mtrx(D, a) for a in arrs]
    D.euler(*ms, 1000.0, 0.005)
    assert np.array_equal(_from_mtrx(ms[0]), port.euler(*arrs, 1000.0, 0.005))
    c = D.continuity(ms[1], ms[2])
    assert np.array_equal(_from_mtrx(c), arrs[1] + arrs[2])
    for m in ms + [F, U, d1, c]:
        D.freem(m)


# ---- the whole driver (cnv_main / cnavier_b200 executable) against the reference executable -----------
def test_driver_log_and_vtk_match_reference_executable(tmp_path, golden_logs):
    """`cnavier_b200 cfg run` vs the unmodified reference binary on the same config file: identical Poisson
    log lines (also equal to the shipped testRunOMP.txt) and byte-identical VTK files (src/utils.c format,
    global file counter naming)."""
    import subprocess
    exe = os.path.join(os.path.dirname(fd._lib.LIB_PATH), "cnavier_b200")
    cfg = dict(api.CONFIG_DEFAULT, tf=(12 + 0.5) * 0.005, output_interval=5)       # 12 steps, dumps at t = 0, 5, 10
    mine = tmp_path / "mine"
    mine.mkdir()
    api.write_config(cfg, str(mine / "cfg.txt"))
    r = subprocess.run([exe, "cfg.txt", "run"], cwd=mine, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, CNV_METRICS_JSON=str(mine / "metrics.json")))
    assert r.returncode == 0, r.stdout[-2000:]
    assert "Poisson SOR parameter: 1.907826" in r.stdout and "Simulation complete!" in r.stdout
    import json
    met = json.loads((mine / "metrics.json").read_text())     # opt-in machine-readable run summary
    assert met["grid"] == [64, 64] and met["steps"] == 12 and met["status"] == 0
    assert met["poisson_k"] == golden_logs["testRunOMP"]["k"][:12]
    assert met["poisson_sweeps"] == sum(k + 1 for k in met["poisson_k"]) and met["poisson_cell_updates_per_s"] > 0
    log = (mine / "output" / "logs" / "run.txt").read_text()
    pl = api.parse_poisson_log(log)
    assert [k for k, _ in pl] == golden_logs["testRunOMP"]["k"][:12]
    assert [e for _, e in pl] == golden_logs["testRunOMP"]["e"][:12]
    assert "Iteration: 11 | Time: 0.055000 | Progress: 100.00% |" in log and "Continuity max:" in log
    vtk = sorted(os.listdir(mine / "output" / "run"))
    assert len(vtk) == 12 and "stream-function-1-0.vtk" in vtk and "y-velocity-1-11.vtk" in vtk
    if api.ref_binary() is None:
        pytest.skip("oracle/_ref not built: VTK byte comparison skipped")
    theirs = tmp_path / "theirs"
    theirs.mkdir()
    api.write_config(cfg, str(theirs / "cfg.txt"))
    subprocess.run([api.ref_binary(), "cfg.txt", "run"], cwd=theirs, check=True, capture_output=True, timeout=600,
                   env=dict(os.environ, OMP_NUM_THREADS="8"))
    for name in vtk:
        a = (mine / "output" / "run" / name).read_bytes()
        b = (theirs / "output" / "run" / name).read_bytes()
        assert a == b, name
    ref_pl = api.parse_poisson_log((theirs / "output" / "logs" / "run.txt").read_text())
    assert pl == ref_pl


def test_driver_exit_code_on_poisson_itmax(tmp_path):
    import subprocess
    exe = os.path.join(os.path.dirname(fd._lib.LIB_PATH), "cnavier_b200")
    cfg = dict(api.CONFIG_DEFAULT, tf=0.02, poisson_max_it=10)
    api.write_config(cfg, str(tmp_path / "cfg.txt"))
    r = subprocess.run([exe, "cfg.txt", "run"], cwd=tmp_path, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, CNV_NO_VTK="1"))
    assert r.returncode == 1                                                    # reference: exit(1), src/poisson.c:284
    assert "Error: maximum number of iterations achieved for Poisson equation." in (tmp_path / "output" / "logs" / "run.txt").read_text()


def test_unmodified_reference_main_drives_gpu_path(tmp_path, golden_logs):
    """INTEGRATION.md route 1: the reference's UNMODIFIED main.c + utils.c linked against
    libcnavier_dropin.so (instead of its own poisson.c / fluiddyn.c / finitediff.c / config.c / linearalg.c):
    the reference driver itself runs the GPU Poisson solve and Euler update.  Same log lines as its shipped
    testRunOMP.txt, same fields as the reference executable's raw dumps."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(fd._lib.LIB_PATH)), "oracle", "_ref", "cnavier_main_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/cnavier_main_dropin not built (needs /root/reference at build time)")
    cfg = dict(api.CONFIG_DEFAULT, tf=(11 + 0.5) * 0.005, output_interval=10)      # 11 steps, dumps at t = 0, 10
    api.write_config(cfg, str(tmp_path / "cfg.txt"))
    dump = tmp_path / "dump"
    dump.mkdir()
    r = subprocess.run([exe, "cfg.txt", "run"], cwd=tmp_path, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, CNAVIER_DUMP_DIR=str(dump), CNAVIER_DUMP_ONLY="1"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    pl = api.parse_poisson_log((tmp_path / "output" / "logs" / "run.txt").read_text())
    assert [k for k, _ in pl] == golden_logs["testRunOMP"]["k"][:11]
    assert [e for _, e in pl] == golden_logs["testRunOMP"]["e"][:11]
    g = load_golden("fields_default_rb.npz")
    for title, key in (("stream-function", "psi"), ("vorticity", "w"), ("x-velocity", "u"), ("y-velocity", "v")):
        got = np.fromfile(dump / (title + ".f64")).reshape(-1, 64, 64)
        assert np.array_equal(got[0], g[key][0]) and np.array_equal(got[1], g[key][1]), key


def test_poisson_weak_scaling_slab_shape_fixed_sweeps(port):
    """BASELINE config 5 per-GPU shape (2048 owned rows of a 16384-wide grid): 8 sweeps, bitwise vs the oracle."""
    nr, nc, sweeps = 2064, 16384, 8
    rng = np.random.default_rng(5)
    f = rng.standard_normal((nr, nc))
    beta = port.beta(nc, nc)
    want, norms = port.poisson_sweeps(f, 1 / nc, 1 / nc, sweeps, beta)
    s = fd.PoissonSolver(nr, nc, 0)
    s.set_consts(1 / nc, 1 / nc, beta)
    s.upload(f)
    r = s.solve(sweeps, 0.0)
    got = s.download(r["buf"])
    assert r["sweeps"] == sweeps and got.tobytes() == want.tobytes()
    assert abs(r["e"] - norms[-1]) <= 1e-11 * norms[-1]
    s.close()


def test_host_allocator_is_page_locked_and_recycled(port):
    """fd.host_empty / cnv_host_alloc (what allocm() of the drop-in library is built on): blocks of >= 1 MiB are page-locked
    (on the GPU's NUMA node where the platform says which one that is), a freed block is recycled for the next request of the
    same size, small blocks stay ordinary heap memory, and a field goes through upload / solve / download from such a block."""
    import gc

    L = fd.lib()
    a = fd.host_empty((1024, 1024))
    assert L.cnv_host_is_pinned(a.ctypes.data) == 1
    small = fd.host_empty((64, 64))
    assert L.cnv_host_is_pinned(small.ctypes.data) == 0
    addr = a.ctypes.data
    del a
    gc.collect()
    b = fd.host_empty((1024, 1024))
    assert b.ctypes.data == addr and L.cnv_host_is_pinned(b.ctypes.data) == 1   # recycled, still page-locked
    n = 520   # 2.2 MB: a page-locked block
    f = fd.host_empty((n, n))
    f[...] = np.random.default_rng(5).standard_normal((n, n))
    assert L.cnv_host_is_pinned(f.ctypes.data) == 1
    beta = port.beta(n, n)
    want = port.poisson(np.array(f), 1.0 / n, 1.0 / n, 20000, 1e-3, beta, redblack=True, sor=True)
    got = fd.poisson_sor(f, 1.0 / n, 1.0 / n, 20000, 1e-3, beta)
    assert got["k"] == want["k"] and np.array_equal(got["u"], want["u"])


def test_host_allocator_numa_placement_path():
    """The NUMA-local form of cnv_host_alloc (anonymous mapping + preferred-node policy + cudaHostRegister), forced onto node 0
    so that it runs on every box: the block is page-locked, survives recycling and release, and a solve from / into such
    blocks gives the bits of a solve from ordinary arrays."""
    import subprocess
    import sys

    code = """
import gc, numpy as np, fluid_dynamics1_b200 as fd
L = fd.lib()
assert L.cnv_host_numa_node() == 0
n = 600
f = fd.host_empty((n, n)); f[...] = np.random.default_rng(3).standard_normal((n, n))
assert L.cnv_host_is_pinned(f.ctypes.data) == 1
beta = fd.sor_beta(n, n)
a = fd.poisson_sor(f, 1.0 / n, 1.0 / n, 64, 0.0, beta, raise_on_itmax=False)
b = fd.poisson_sor(np.array(f), 1.0 / n, 1.0 / n, 64, 0.0, beta, raise_on_itmax=False)
assert np.array_equal(a["u"], b["u"]) and np.isfinite(a["u"]).all() and np.abs(a["u"]).max() > 0
addr = f.ctypes.data
del f; gc.collect()
g = fd.host_empty((n, n)); assert g.ctypes.data == addr and L.cnv_host_is_pinned(g.ctypes.data) == 1
big = [fd.host_empty((4096, 4096)) for _ in range(3)]   # beyond what the pool keeps idle only when freed: exercise release
del big; gc.collect()
print("NUMA PATH OK")
"""
    env = dict(os.environ, CNV_HOST_NUMA_NODE="0")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0 and "NUMA PATH OK" in r.stdout, r.stdout + r.stderr

