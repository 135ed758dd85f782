"""Small-grid timing: full solves (reference stopping rule) and whole time steps with the persistent on-chip kernel vs the
streaming kernel.  python tools/probe_small.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluid_dynamics1_b200 as fd

def sine(n):
    x = np.arange(n) / n
    return -2 * np.pi ** 2 * np.outer(np.sin(np.pi * x), np.sin(np.pi * x))

for n in (64, 128, 256, 352):
    f = sine(n); beta = fd.sor_beta(n, n)
    for path in ("1", "0"):
        os.environ["CNV_POISSON_ONCHIP"] = path
        s = fd.PoissonSolver(n, n, 0); s.set_consts(1 / n, 1 / n, beta)
        s.upload(f); r = s.solve(100000, 1e-3)
        t0 = time.perf_counter()
        for _ in range(5):
            s.upload(f); r = s.solve(100000, 1e-3)
        dt = (time.perf_counter() - t0) / 5
        print(f"n={n} path={'onchip' if path=='1' else 'stream'} sweeps={r['sweeps']} solve={dt*1e3:.3f} ms  {dt/r['sweeps']*1e6:.2f} us/sweep", flush=True)
        s.close()
for name, cfg, steps in (("config_default 64^2", dict(), 200), ("config_high_re 128^2", dict(Re=5000.0, nx=128, ny=128, dt=0.001, poisson_max_it=15000, poisson_tol=5e-4), 100)):
    for path in ("1", "0"):
        os.environ["CNV_POISSON_ONCHIP"] = path
        sim = fd.Simulation(cfg); sim.step(5)
        t0 = time.perf_counter(); r = sim.step(steps, diagnostics=True); dt = time.perf_counter() - t0
        print(f"{name} path={'onchip' if path=='1' else 'stream'}: {dt/steps*1e3:.3f} ms/step (k last {r['k'][-1]})", flush=True)
        sim.close()
