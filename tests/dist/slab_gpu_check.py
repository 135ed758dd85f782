"""Multi-GPU check of the slab-decomposed Poisson solve.  Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist/slab_gpu_check.py [rows] [cols] [T]
Every rank runs the distributed solve; rank 0 compares the gathered field bitwise with the single-GPU
solve and with the oracle (same sweep count, same bits: red-black colouring uses the global parity)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

import fluid_dynamics1_b200 as fd
from fluid_dynamics1_b200 import parallel


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    fd.lib().cnv_set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(5)
    f = rng.standard_normal((rows, cols))
    dx, dy = 1.0 / cols, 1.0 / rows
    beta = fd.sor_beta(rows, cols)
    ok = True
    for tol, itmax in ((1e-4, 100000), (0.0, 37)):
        slab = parallel.SlabPoisson(rows, cols, T, rank, world, stream=sp)
        slab.set_consts(dx, dy, beta)
        slab.upload_owned(f[slab.row0:slab.row0 + slab.own_rows], 1.0)
        res = slab.solve(itmax, tol)
        full = slab.gather_result(res["buf"])
        slab.close()   # quiesce + barrier + unmap the neighbours' buffers before anybody frees its own
        if rank == 0:
            single = fd.poisson_sor(f, dx, dy, itmax, tol, beta, T=T, raise_on_itmax=False)
            same_k = res["k"] == single["k"] and res["status"] == single["status"]
            same_bits = full.tobytes() == single["u"].tobytes()
            print(f"world={world} {rows}x{cols} T={T} tol={tol}: k={res['k']} (single {single['k']}) e={res['e']:.6E} "
                  f"(single {single['e']:.6E}) passes={res['passes']} bitwise_equal={same_bits}", flush=True)
            ok = ok and same_k and same_bits and abs(res["e"] - single["e"]) <= 1e-12 * max(single["e"], 1e-300)
            try:
                from oracle import api
                o = api.port().poisson(f, dx, dy, itmax, tol, beta, redblack=True)
                ok = ok and o["k"] == res["k"] and o["u"].tobytes() == full.tobytes()
                print("  oracle: k =", o["k"], "bitwise_equal =", o["u"].tobytes() == full.tobytes(), flush=True)
            except Exception as exc:  # oracle not built on this box
                print("  oracle unavailable:", exc)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB CHECK", "PASSED" if ok else "FAILED", flush=True)
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
