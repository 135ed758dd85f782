"""Multi-GPU check of the slab-decomposed TIME STEPPING.  Launch with torch.distributed.run --nproc-per-node N.
Rank 0 compares psi, w, u, v and the per-step Poisson log values bitwise with the single-GPU simulation (and
the oracle when available)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import fluid_dynamics1_b200 as fd
from fluid_dynamics1_b200 import parallel


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    fd.lib().cnv_set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    cfg = dict(nx=n, ny=n, Re=1000.0, dt=0.005 * 64 / n, poisson_max_it=100000, u1=0.05, u2=-0.03, v3=0.02)
    sim = parallel.SlabSimulation(cfg, rank, world, T=T, stream=sp)
    res = sim.step(steps)
    fields = sim.gather_fields()
    ok = True
    if rank == 0:
        one = fd.Simulation(cfg, T=T)
        ref = one.step(steps)
        rf = one.fields()
        ok = list(res["k"]) == list(ref["k"]) and res["failed_step"] == 0
        for name in ("psi", "w", "u", "v"):
            same = fields[name].tobytes() == rf[name].tobytes()
            ok = ok and same
            print(f"  {name}: bitwise_equal={same}", flush=True)
        ok = ok and np.allclose(res["cont_max"], ref["cont_max"], rtol=0, atol=0) and np.allclose(res["cont_min"], ref["cont_min"], rtol=0, atol=0)
        print(f"world={world} {n}x{n} steps={steps}: k={list(res['k'])} single={list(ref['k'])} cont_max={res['cont_max'][-1]:.3E}", flush=True)
        try:
            from oracle import api
            o = api.port().run(dict(api.CONFIG_DEFAULT, **cfg), steps, redblack=True)
            osame = all(o[nm].tobytes() == fields[nm].tobytes() for nm in ("psi", "w", "u", "v")) and list(o["k"]) == list(res["k"])
            print("  oracle: bitwise_equal =", osame, flush=True)
            ok = ok and osame
        except Exception as exc:
            print("  oracle unavailable:", exc)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    sim.close()
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB SIM CHECK", "PASSED" if flag.item() == 1 else "FAILED", flush=True)
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
