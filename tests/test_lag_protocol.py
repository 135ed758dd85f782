"""The stop-decision state machines of csrc/poisson_stream.h under a multi-domain protocol simulation, on the CPU.
  * plain machine (lag = 0): the peer path's own start-of-pass code (peer_needs_norms / peer_advance) -- a pass folds the
    norms of the previous pass, two iterate buffers;
  * lagged machine (lag = 1: lag_fold / lag_action / lag_final, the decision of the persistent on-chip kernel, where the
    domains are the tiles of one GPU; round 2 also ran it between GPUs, see profiles/scale_r2.md) -- a pass folds the
    norms of the pass before the previous one and runs speculatively, three iterate buffers.
The domains of a slab decomposition are simulated in ONE process, each pass executed by the schedule emulator (tests/emul,
the kernel's own per-thread code), with exactly the waits of the protocol -- norm flags from every domain, halo pushes of
pass p-1 from the neighbours -- and a RANDOM interleaving of the domains within those constraints (each runs as far ahead
as the protocol lets it).  Norms travel through mailbox slots indexed by the GLOBAL pass number (which keeps counting
across solves), pushes land in the neighbours' halo rows.  The result must be the single-domain oracle's field, iteration
count and residual, bit for bit, for every position of the converged sweep inside a pass and every position of the host's
batch boundary."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from fluid_dynamics1_b200.parallel import slab_layout
from oracle import api

NSLOTS = 8  # csrc/poisson_stream.h kNormSlots
dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def E():
    lib = C.CDLL(os.path.join(ROOT, "tests", "emul", "libstream_emul.so"))
    lib.emul_pass.argtypes = [C.c_int] * 10 + [C.c_double] * 3 + [C.c_int, dp, dp, dp, C.c_int, dp]
    lib.emul_lag_fold.argtypes = [ip, dp, dp, C.c_int, C.c_void_p]
    lib.emul_lag_action.argtypes = [ip, dp, C.c_int, C.c_int, ip]
    lib.emul_lag_final.argtypes = [ip, dp, C.c_int, dp, C.c_int, C.c_void_p]
    lib.emul_peer_needs_norms.argtypes = [ip, dp, C.c_int, C.c_int]
    lib.emul_peer_advance.argtypes = [ip, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip]
    return lib


class Rank:
    def __init__(self, E, r, world, rows, cols, T, f, itmax, tol, lag=1):
        self.E, self.r, self.world, self.T = E, r, world, T
        self.lag, self.lagd, self.nbuf = lag, (2 if lag else 1), (3 if lag else 2)
        self.rows, self.cols = rows, cols
        self.grow0, self.nloc, self.own_lo, self.own_hi, _, _ = slab_layout(rows, world, r, T)
        self.ld = (cols + 15) // 16 * 16
        self.floc = np.zeros((self.nloc, self.ld))
        self.floc[:, :cols] = f[self.grow0:self.grow0 + self.nloc]
        self.bufs = [np.zeros((self.nloc, self.ld)) for _ in range(self.nbuf)]
        # chain X_p: two slots like ctlbuf[p & 1]; slot 0 = reset state
        ints = np.array([0 if itmax > 0 else 2, 0, 0, 0, itmax, -1, 0, self.nbuf], dtype=np.int32)
        dbls = np.array([tol, 0.0, 0.0, 0.0])
        self.chain = [(ints.copy(), dbls.copy()), (ints.copy(), dbls.copy())]
        self.norms = np.zeros((NSLOTS, world, 8))  # my mailbox: [slot][rank][sweep]
        self.norm_flag = np.zeros(world, dtype=np.int64)
        self.halo_passes = [0, 0]                  # pushes landed in my low / high halo, in passes (cumulative)
        self.p = 0                                 # next pass of the current solve
        self.g0 = 0                                # global index of pass 0 of the current solve

    def new_solve(self, f, itmax, tol):
        """Re-initialise for the next solve (the kernel path: quiesce, zero the iterate, ready); the mailbox counters
        keep counting."""
        self.g0 += self.p
        self.p = 0
        self.floc[:, :self.cols] = f[self.grow0:self.grow0 + self.nloc]
        for b in self.bufs:
            b[:] = 0.0
        ints = np.array([0 if itmax > 0 else 2, 0, 0, 0, itmax, -1, 0, self.nbuf], dtype=np.int32)
        dbls = np.array([tol, 0.0, 0.0, 0.0])
        self.chain = [(ints.copy(), dbls.copy()), (ints.copy(), dbls.copy())]

    def gather(self, p):
        g = self.g0 + p
        assert all(self.norm_flag >= g + 1)
        e = np.zeros(8)
        for i in range(8):
            s = 0.0
            for r in range(self.world):
                s = s + self.norms[g % NSLOTS, r, i]
            e[i] = s
        return e

    def start_of_pass(self, p):
        """What the kernel's preamble does (csrc/poisson_stream.h peer_needs_norms / peer_advance): returns
        (state of this pass, action, norms available?)."""
        ints, dbls = (a.copy() for a in self.chain[0 if p == 0 else (p - 1) & 1])
        need = bool(self.E.emul_peer_needs_norms(ints, dbls, p, self.lag))
        if need and not all(self.norm_flag >= self.g0 + p - self.lagd + 1):
            return None
        e = self.gather(p - self.lagd) if need else np.zeros(8)
        act = np.zeros(4, dtype=np.int32)
        self.E.emul_peer_advance(ints, dbls, e, int(need), 0, p, self.lag, self.T, act)
        return ints, dbls, [int(x) for x in act]

    def runnable(self, ranks):
        p = self.p
        st = self.start_of_pass(p)
        if st is None:
            return False
        kind = st[2][0]
        # a working pass of the plain machine, or a speculative run of the lagged one, streams halo rows: the
        # neighbours' pushes of pass p-1 must have landed (the lagged redo pass re-reads an older buffer)
        if kind == 1:
            if self.r > 0 and self.halo_passes[0] < self.g0 + p:
                return False
            if self.r < self.world - 1 and self.halo_passes[1] < self.g0 + p:
                return False
        return True

    def run_pass(self, ranks, dx, dy, beta):
        p, T, H = self.p, self.T, 2 * self.T
        ints, dbls, (kind, bi, bo, nsw) = self.start_of_pass(p)
        if p > 0:
            self.chain[p & 1] = (ints.copy(), dbls.copy())
        norms = np.zeros(8)
        if kind != 0:
            # write-after-read check of the protocol: nobody may still have to READ the buffer this pass pushes into
            # (lagged: the neighbours read it in pass p-2; plain: in pass p-1)
            for nb in (self.r - 1, self.r + 1):
                if 0 <= nb < self.world and kind == 1:
                    assert ranks[nb].p >= p - self.lagd + 1, "push into a buffer the neighbour has not finished reading"
            geo = (self.nloc, self.cols, self.ld, self.grow0, self.rows, self.own_lo, self.own_hi)
            rc = self.E.emul_pass(T, *geo, 0, int(os.environ.get("CNV_TEST_CHUNKS", "0")), dx, dy, beta, 0, self.bufs[bi],
                                  self.floc, self.bufs[bo], nsw, norms)
            assert rc == 0
            if self.r > 0:                     # my first owned rows -> the lower neighbour's high halo
                nb = ranks[self.r - 1]
                nb.bufs[bo][nb.own_hi:nb.own_hi + H] = self.bufs[bo][self.own_lo:self.own_lo + H]
            if self.r < self.world - 1:        # my last owned rows -> the upper neighbour's low halo
                nb = ranks[self.r + 1]
                nb.bufs[bo][nb.own_lo - H:nb.own_lo] = self.bufs[bo][self.own_hi - H:self.own_hi]
        if self.r > 0:
            ranks[self.r - 1].halo_passes[1] += 1
        if self.r < self.world - 1:
            ranks[self.r + 1].halo_passes[0] += 1
        for other in ranks:                    # publish; a no-op pass only raises the flag (like the kernel): a finished
            if kind != 0:                      # rank runs ahead through its no-ops and must not touch the norm slots
                other.norms[(self.g0 + p) % NSLOTS, self.r, :] = norms
            other.norm_flag[self.r] = self.g0 + p + 1
        self.p += 1

    def finalize(self, P):
        """k_peer_finalize: the state the host reads after P launched passes."""
        if self.lag:
            ints, dbls = (a.copy() for a in self.chain[0 if P == 0 else (P - 1) & 1])
            if P >= 2:
                need = ints[0] == 0 and ints[3] == 0
                self.E.emul_lag_fold(ints, dbls, self.gather(P - 2) if need else np.zeros(8), self.T, None)
            e_last = self.gather(P - 1) if P >= 1 else np.zeros(8)
            self.E.emul_lag_final(ints, dbls, P, e_last, self.T, None)
            return ints, dbls
        st = self.start_of_pass(P)             # plain machine: S_P = what pass P itself would derive
        assert st is not None
        return st[0], st[1]


def lagged_solve(E, rows, cols, T, world, itmax, tol, batches, seed, ranks=None, fseed=11, lag=1):
    rng = np.random.default_rng(seed)
    f = np.random.default_rng(fseed).standard_normal((rows, cols))
    dx, dy = 1.0 / cols, 1.0 / rows
    port = api.port()
    beta = port.beta(rows, cols)
    if ranks is None:
        ranks = [Rank(E, r, world, rows, cols, T, f, itmax, tol, lag) for r in range(world)]
    else:
        for k in ranks:
            k.new_solve(f, itmax, tol)
    P, nb = 0, 0
    while True:
        P += batches[min(nb, len(batches) - 1)]
        nb += 1
        while any(k.p < P for k in ranks):
            ready = [k for k in ranks if k.p < P and k.runnable(ranks)]
            assert ready, "protocol deadlock"
            ready[rng.integers(len(ready))].run_pass(ranks, dx, dy, beta)
        finals = [k.finalize(P) for k in ranks]
        assert all(np.array_equal(finals[0][0], x[0]) and np.array_equal(finals[0][1], x[1]) for x in finals)
        ints, dbls = finals[0]
        if ints[0] != 0:
            break
        assert P < itmax + 8
    full = np.concatenate([k.bufs[int(ints[1])][k.own_lo:k.own_hi, :cols] for k in ranks])
    want = port.poisson(f, dx, dy, itmax, tol, beta, redblack=True)
    lagged_solve.last_ranks = ranks
    return ints, dbls, full, want, P


@pytest.mark.parametrize("world,T", [(2, 2), (3, 2), (2, 4)])
def test_lagged_machine_every_stop_position(E, world, T):
    """tol chosen so that the converged sweep falls on every position inside a pass; batch sizes 1..4 move the host's
    read-back (finalize) over every phase of the speculative pass / redo pass sequence."""
    rows, cols = 24 * world, 40
    f = np.random.default_rng(11).standard_normal((rows, cols))
    port = api.port()
    # residual history of a run that never converges (tol = 0 -> itmax sweeps): hist[k] = norm of sweep k
    hist = np.zeros(8 * T)
    for ks in range(3 * T, 5 * T + 1):
        hist[ks] = port.poisson(f, 1.0 / cols, 1.0 / rows, ks + 1, 0.0, port.beta(rows, cols), redblack=True)["e"]
    seen = set()
    for ksweep in range(3 * T, 5 * T + 1):
        # a tolerance just above the norm of sweep `ksweep`: first sweep with e < tol
        tol = hist[ksweep] * (1 + 1e-9)
        for batch in ([1], [2], [3], [4], [ksweep // T + 1, 1], [ksweep // T + 2, 2]):
            ints, dbls, full, want, P = lagged_solve(E, rows, cols, T, world, 5000, tol, batch, seed=ksweep * 7 + batch[0])
            assert ints[0] == 1 and want["status"] == 0
            assert int(ints[5]) == want["k"], (ksweep, batch, int(ints[5]), want["k"])
            assert full.tobytes() == want["u"].tobytes(), (ksweep, batch)
            assert abs(dbls[1] - want["e"]) <= 1e-12 * want["e"]
            seen.add(want["k"] % T)
    assert seen == set(range(T))


@pytest.mark.parametrize("itmax", [1, 2, 3, 4, 5, 7, 8, 9])
def test_lagged_machine_itmax(E, itmax):
    """No convergence (tol = 0): exactly itmax sweeps, state 2, whatever the batch size; the speculative pass past
    itmax is a no-op."""
    for batch in ([1], [2], [5]):
        ints, dbls, full, want, P = lagged_solve(E, 48, 40, 2, 2, itmax, 0.0, batch, seed=itmax)
        assert ints[0] == 2 and want["status"] == 1 and int(ints[2]) == itmax
        assert full.tobytes() == want["u"].tobytes()


def test_lagged_machine_first_sweep_converges(E):
    """Huge tolerance: the very first sweep is below it (hit inside pass 0, redo pass with 1 sweep)."""
    ints, dbls, full, want, P = lagged_solve(E, 48, 40, 4, 2, 100, 1e30, [1], seed=3)
    assert ints[0] == 1 and int(ints[5]) == want["k"] == 0
    assert full.tobytes() == want["u"].tobytes()


def test_lagged_machine_consecutive_solves(E):
    """Several solves on the same ranks: the global pass index (norm slots, flags, halo counters) keeps counting while
    the per-solve pass number restarts, as in a time-stepping run."""
    rows, cols, T, world = 72, 40, 2, 3
    ranks = None
    for i, (tol, batch) in enumerate([(0.05, [3]), (0.2, [1]), (0.01, [7, 2]), (1e30, [2]), (0.1, [4])]):
        ints, dbls, full, want, P = lagged_solve(E, rows, cols, T, world, 5000, tol, batch, seed=i, ranks=ranks, fseed=20 + i)
        assert ints[0] == 1 and int(ints[5]) == want["k"]
        assert full.tobytes() == want["u"].tobytes()
        ranks = lagged_solve.last_ranks


@pytest.mark.parametrize("seed", range(6))
def test_lagged_machine_random_schedules_world4(E, seed):
    ints, dbls, full, want, P = lagged_solve(E, 96, 40, 4, 4, 5000, 0.02 * (1 + seed), [2 + seed % 3], seed=100 + seed)
    assert ints[0] == 1 and int(ints[5]) == want["k"]
    assert full.tobytes() == want["u"].tobytes()


def test_lagged_machine_abstract_model_fuzz(E):
    """State machine alone, no arithmetic: a buffer is modelled by the number of sweeps applied to it, a pass by
    content[out] = content[in] + nsw and the norms h[content[in] : content[in] + nsw] of a synthetic residual history.
    For thousands of random (T, itmax, history, batch pattern) the lagged machine must end with the buffer that holds
    exactly k+1 sweeps, k = first index with h[k] < tol (or itmax sweeps, state 2) -- the plain machine's answer."""
    rng = np.random.default_rng(2024)
    for case in range(3000):
        T = int(rng.choice([1, 2, 4, 6, 8]))
        itmax = int(rng.integers(1, 60))
        n = itmax + 4 * T
        h = np.sort(rng.random(n))[::-1].copy() + 0.5          # decreasing, > 0.5
        if rng.random() < 0.8:
            kstop = int(rng.integers(0, itmax + 6))             # first sweep below tol (possibly beyond itmax)
            tol = 0.25
            h[kstop:] = rng.random(n - kstop) * 0.2 if kstop < n else h[kstop:]
        else:
            kstop, tol = n + 1, 0.0
        want_sweeps = min(kstop + 1, itmax) if kstop < itmax else itmax
        want_state = 1 if kstop < itmax else 2
        chain = [None, None]
        ints0 = np.array([0, 0, 0, 0, itmax, -1, 0, 3], dtype=np.int32)
        dbls0 = np.array([tol, 0.0, 0.0, 0.0])
        chain[0] = (ints0.copy(), dbls0.copy())
        content = [0, -1, -1]                                   # sweeps applied to each buffer (-1: garbage)
        norms = {}                                              # pass -> e[8]
        P = 0
        while True:
            P_next = P + int(rng.integers(1, 6))
            for p in range(P, P_next):
                ints, dbls = (a.copy() for a in chain[0 if p == 0 else (p - 1) & 1])
                if p >= 2:
                    need = ints[0] == 0 and ints[3] == 0
                    E.emul_lag_fold(ints, dbls, norms[p - 2] if need else np.zeros(8), T, None)
                if p > 0:
                    chain[p & 1] = (ints.copy(), dbls.copy())
                act = np.zeros(4, dtype=np.int32)
                E.emul_lag_action(ints, dbls, p, T, act)
                kind, bi, bo, nsw = (int(x) for x in act)
                e = np.zeros(8)
                if kind:
                    assert content[bi] >= 0 and bi != bo and 1 <= nsw <= T
                    if kind == 1:
                        assert bi == p % 3 and bo == (p + 1) % 3   # the rotation the write-after-read argument relies on
                    else:
                        assert bi == (p - 2) % 3 and bo == p % 3   # redo: input of pass p-2, into the speculative pass' output
                    e[:nsw] = h[content[bi]:content[bi] + nsw]
                    content[bo] = content[bi] + nsw
                norms[p] = e
            P = P_next
            ints, dbls = (a.copy() for a in chain[0 if P == 0 else (P - 1) & 1])
            if P >= 2:
                need = ints[0] == 0 and ints[3] == 0
                E.emul_lag_fold(ints, dbls, norms[P - 2] if need else np.zeros(8), T, None)
            E.emul_lag_final(ints, dbls, P, norms[P - 1], T, None)
            if ints[0] != 0:
                break
            assert P <= itmax + 12
        assert int(ints[0]) == want_state, (case, T, itmax, kstop, ints)
        assert int(ints[2]) == want_sweeps and content[int(ints[1])] == want_sweeps, (case, T, itmax, kstop, ints, content)
        if want_state == 1:
            assert int(ints[5]) == kstop and dbls[1] == h[kstop]


@pytest.mark.parametrize("world,T", [(2, 2), (3, 4)])
def test_plain_peer_machine_same_simulation(E, world, T):
    """The same simulation with the PLAIN peer machine (lag = 0: two buffers, a pass folds the norms of the previous
    pass): the kernel's start-of-pass code is shared with this test (peer_needs_norms / peer_advance), so this covers
    the default multi-GPU path's decision logic, buffer alternation, redo pass and no-op passes."""
    rows, cols = 24 * world, 40
    f = np.random.default_rng(11).standard_normal((rows, cols))
    port = api.port()
    for ksweep in range(2 * T, 3 * T + 1):
        tol = port.poisson(f, 1.0 / cols, 1.0 / rows, ksweep + 1, 0.0, port.beta(rows, cols), redblack=True)["e"] * (1 + 1e-9)
        for batch in ([1], [3], [ksweep // T + 1, 2]):
            ints, dbls, full, want, P = lagged_solve(E, rows, cols, T, world, 5000, tol, batch, seed=ksweep + batch[0], lag=0)
            assert ints[0] == 1 and int(ints[5]) == want["k"]
            assert full.tobytes() == want["u"].tobytes()
    ints, dbls, full, want, P = lagged_solve(E, rows, cols, T, world, 7, 0.0, [2], seed=1, lag=0)
    assert ints[0] == 2 and int(ints[2]) == 7 and full.tobytes() == want["u"].tobytes()


def test_peer_protocol_random_configurations(E):
    """Seeded sweep over the whole protocol simulation: 2-4 ranks, T = 2/4/8, plain and lagged machine,
    random tolerance (converging anywhere or hitting itmax) and random host batch sizes."""
    rng = np.random.default_rng(77)
    for case in range(24):
        world = int(rng.integers(2, 5))
        T = int(rng.choice([2, 4, 8]))
        lag = int(rng.integers(0, 2))
        rng.integers(0, 3)                       # (keeps the seeded sequence of the cases below unchanged)
        rows = int(rng.integers(max(4 * T + 2, 24), 70)) * world
        cols = int(rng.integers(40, 90))
        tol = float(10 ** rng.uniform(-3, 0.5))
        itmax = int(rng.choice([5000, 5000, 17, 40]))
        batches = [int(x) for x in rng.integers(1, 7, size=3)]
        ints, dbls, full, want, P = lagged_solve(E, rows, cols, T, world, itmax, tol, batches, seed=case, lag=lag, fseed=case)
        where = dict(case=case, world=world, T=T, lag=lag, rows=rows, cols=cols, tol=tol, itmax=itmax, batches=batches)
        assert int(ints[0]) == (1 if want["status"] == 0 else 2), where
        assert full.tobytes() == want["u"].tobytes(), where
        if want["status"] == 0:
            assert int(ints[5]) == want["k"], where
