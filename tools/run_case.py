"""Whole time-stepping runs of the BASELINE configurations on the device-resident stepper (timed on the host around
cnv_sim_step, fields stay in HBM, no VTK):  python tools/run_case.py default|high_re|c3|c4 [steps]
  default  config_default.txt   64^2   Re 1000  dt .005  4000 steps  tol 1e-3
  high_re  config_high_re.txt   128^2  Re 5000  dt .001  10000 steps tol 5e-4
  c3       1024^2 Re 1000 dt 1e-4 tol 1e-3 (BASELINE config 3; default 100 steps here, the config asks for 10000)
  c4       4096^2 Re 1000 dt 5e-6 tol 1e-3 (BASELINE config 4 as a whole time step; ~13k sweeps per step)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluid_dynamics1_b200 as fd

CASES = {
    "default": (dict(), 4000),
    "high_re": (dict(Re=5000.0, nx=128, ny=128, dt=0.001, tf=10.0, poisson_max_it=15000, poisson_tol=5e-4), 10000),
    "c3": (dict(nx=1024, ny=1024, Re=1000.0, dt=1e-4, poisson_max_it=20000, poisson_tol=1e-3), 100),
    "c4": (dict(nx=4096, ny=4096, Re=1000.0, dt=5e-6, poisson_max_it=100000, poisson_tol=1e-3), 3),
}
name = sys.argv[1]
cfg, steps = CASES[name]
if len(sys.argv) > 2:
    steps = int(sys.argv[2])
sim = fd.Simulation(cfg)
sim.step(2)
c0 = sim.counters()
t0 = time.perf_counter()
r = sim.step(steps, diagnostics=True)
dt = time.perf_counter() - t0
c1 = sim.counters()
n = sim.shape[0] * sim.shape[1]
sw = c1["sweeps"] - c0["sweeps"]
f = sim.fields()
print(f"{name}: grid {sim.shape[0]}x{sim.shape[1]} steps={steps} failed_step={r['failed_step']} wall={dt:.3f} s  {dt / steps * 1e3:.3f} ms/step  "
      f"sweeps={sw} ({sw / steps:.1f}/step, {dt / max(sw, 1) * 1e6:.2f} us/sweep incl. stencils)  timestep cell-updates/s={n * steps / dt:.3e}  "
      f"poisson cell-updates/s={(sim.shape[0] - 2) * (sim.shape[1] - 2) * sw / dt:.3e}  last k={int(r['k'][-1])} e={r['e'][-1]:.6E} "
      f"cont_max={r['cont_max'][-1]:.3E}  |psi|max={np.abs(f['psi']).max():.6e}", flush=True)
sim.close()
