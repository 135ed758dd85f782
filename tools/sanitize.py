"""Small workload touching every kernel, for compute-sanitizer (racecheck / memcheck / synccheck / initcheck):
compute-sanitizer --tool racecheck python tools/sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluid_dynamics1_b200 as fd

rng = np.random.default_rng(0)
only = sys.argv[1] if len(sys.argv) > 1 else None
for path, env in (("onchip", dict(CNV_POISSON_ONCHIP="1")),
                  ("stream", dict(CNV_POISSON_ONCHIP="0"))):
    if only and path != only:
        continue
    os.environ.update(env)
    # whole time steps (stencil kernels + Poisson solve with the stopping logic) on the reference's default grid
    sim = fd.Simulation(dict(nx=64, ny=64, Re=1000.0, dt=0.005, poisson_max_it=10000, poisson_tol=1e-3))
    r = sim.step(2)
    p = sim.pressure()
    print(path, "steps k =", list(r["k"]), "pressure k =", p["k"], flush=True)
    sim.close()
    # a non-square general-arithmetic grid with several strips / chunks, fixed sweeps
    for T in (4, 8):
        s = fd.PoissonSolver(150, 330, T)
        s.set_consts(0.01, 0.013, 1.7)
        s.upload(rng.standard_normal((150, 330)))
        res = s.solve(2 * T + 3, 0.0)
        print(path, "T", T, "plan", {k: s.plan[k] for k in ("onchip", "WS", "nstrips", "nchunks", "oc_T", "oc_ntx", "oc_nty")}, "sweeps", res["sweeps"], flush=True)
        s.close()
print("SANITIZE WORKLOAD DONE")
