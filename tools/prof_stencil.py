"""The explicit-stencil phase of a time step alone (for ncu): python tools/prof_stencil.py [N] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluid_dynamics1_b200 as fd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sim = fd.Simulation(fd.config_from_dict(dict(nx=n, ny=n, Re=1000.0, dt=5e-6, poisson_max_it=100000)))
L = fd.lib()
L.cnv_sim_stencil_phase(sim.h, reps, None)
L.cnv_device_synchronize()
sim.close()
print("done")
