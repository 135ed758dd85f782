"""Stress / diagnosis of the multi-GPU Poisson paths: repeat a converging solve, compare the per-sweep residual history
and the final field with the single-GPU solve, report the first deviating sweep.
torchrun ... tests/dist/slab_stress.py ROWS COLS T REPEATS"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import fluid_dynamics1_b200 as fd
from fluid_dynamics1_b200 import parallel

rows, cols, T, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
tol = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-4
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); fd.lib().cnv_set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
L = fd.lib()
rng = np.random.default_rng(5)
f = rng.standard_normal((rows, cols)); dx, dy = 1.0 / cols, 1.0 / rows; beta = fd.sor_beta(rows, cols)
itmax = 100000
if rank == 0:
    single = fd.poisson_sor(f, dx, dy, itmax, tol, beta, T=T, history=True)
bad = 0
slab = parallel.SlabPoisson(rows, cols, T, rank, world, stream=sp)
slab.set_consts(dx, dy, beta)
if rank == 0: print("backend peer" if slab.peer else ("nccl" if slab.comm else "torch"), flush=True)
for it in range(reps):
    slab.upload_owned(f[slab.row0:slab.row0 + slab.own_rows], 1.0)
    L.cnv_poisson_enable_history(slab.h, 4096 * 4)
    res = slab.solve(itmax, tol)
    full = slab.gather_result(res["buf"])
    hist = np.zeros(res["sweeps"]); L.cnv_poisson_read_history(slab.h, hist, res["sweeps"])
    if rank == 0:
        n = min(len(hist), len(single["history"]))
        rel = np.abs(hist[:n] - single["history"][:n]) / single["history"][:n]
        first = int(np.argmax(rel > 1e-10)) if (rel > 1e-10).any() else -1
        same = full.tobytes() == single["u"].tobytes() and res["k"] == single["k"]
        bad += 0 if same else 1
        print(f"rep {it}: k={res['k']} (single {single['k']}) passes={res['passes']} same={same} first_dev_sweep={first}"
              + (f" rel={rel[first]:.2e} pass={first // slab.T}" if first >= 0 else ""), flush=True)
slab.close()
if rank == 0: print("STRESS", "PASSED" if bad == 0 else f"FAILED ({bad}/{reps})", flush=True)
dist.destroy_process_group()
