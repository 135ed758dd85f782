"""One fixed-sweep run of the streaming kernel for ncu: python tools/prof_one.py N T [WS] [CHUNKS] [SWEEPS]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
n, T = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 3: os.environ["CNV_POISSON_WS"] = sys.argv[3]
if len(sys.argv) > 4: os.environ["CNV_POISSON_CHUNKS"] = sys.argv[4]
sweeps = int(sys.argv[5]) if len(sys.argv) > 5 else 8 * T
import fluid_dynamics1_b200 as fd
s = fd.PoissonSolver(n, n, T)
s.set_consts(1.0 / n, 1.0 / n, fd.sor_beta(n, n))
s.upload(np.random.default_rng(0).standard_normal((n, n)))
s.reset(sweeps, 0.0); s.enqueue((sweeps + T - 1) // T); fd.lib().cnv_device_synchronize()
print(s.plan, s.state())
