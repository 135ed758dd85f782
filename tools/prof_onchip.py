"""One fixed-sweep solve on the persistent on-chip kernel for ncu: python tools/prof_onchip.py N SWEEPS"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
os.environ.setdefault("CNV_POISSON_ONCHIP", "1")
n, sweeps = int(sys.argv[1]), int(sys.argv[2])
import fluid_dynamics1_b200 as fd
s = fd.PoissonSolver(n, n, 0)
s.set_consts(1.0 / n, 1.0 / n, fd.sor_beta(n, n))
s.upload(np.random.default_rng(0).standard_normal((n, n)))
print(s.plan, s.solve(sweeps, 0.0))
