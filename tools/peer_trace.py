"""Where does a multi-GPU pass spend its time?  Per-CTA timestamps of the peer-memory path (cnv_poisson_peer_trace):
torchrun --nproc-per-node N tools/peer_trace.py [ROWS_PER_GPU COLS PASSES]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import fluid_dynamics1_b200 as fd
from fluid_dynamics1_b200 import parallel

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
npass = int(sys.argv[3]) if len(sys.argv) > 3 else 24
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); fd.lib().cnv_set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
L = fd.lib()
slab = parallel.SlabPoisson(rows * world, cols, 8, rank, world, stream=sp)
assert slab.peer, "peer backend not active"
slab.set_consts(1.0 / cols, 1.0 / cols, fd.sor_beta(cols, cols))
slab.upload_owned(np.random.default_rng(rank).standard_normal((slab.own_rows, cols)), 1.0)
T = slab.T
for rep in range(2):                       # the second repetition is the one reported (warm)
    ctas = L.cnv_poisson_peer_trace(slab.h, npass)
    slab.zero_iterate(); slab.reset(npass * T, 0.0)
    slab.enqueue(npass)
    buf = np.zeros(npass * ctas * 6, dtype=np.uint64)
    L.cnv_poisson_peer_trace_read(slab.h, buf.ctypes.data, buf.size)
t = buf.reshape(npass, ctas, 6).astype(np.int64)
us = lambda x: x / 1e3
lines = []
for p in range(4, npass):
    s0 = t[p, :, 0].min()
    waits = t[p, :, 2]; halo = np.where(waits > 0, waits - t[p, :, 1], 0)
    push = np.where(t[p, :, 4] > 0, t[p, :, 4] - t[p, :, 3], 0)
    lines.append((us(t[p, :, 5].max() - s0), us(t[p, :, 0].max() - s0), us((t[p, :, 1] - t[p, :, 0]).max()), us(halo.max()),
                  us((t[p, :, 3] - np.maximum(t[p, :, 1], t[p, :, 2])).mean()), us((t[p, :, 3] - np.maximum(t[p, :, 1], t[p, :, 2])).max()),
                  us(push.max()), us(t[p, :, 5].max() - t[p, :, 3].max()), us(t[p, :, 0].min() - t[p - 1, :, 5].max()),
                  us(t[p, :, 0].min() - t[p - 1, :, 0].min())))
a = np.array(lines)
names = ["pass span", "start spread", "max wait for norms", "max wait for halos", "mean stream time", "max stream time", "max push time",
         "tail after last stream end", "gap to previous pass", "period"]
out = f"rank {rank}/{world} {rows}x{cols} ctas={ctas}: " + "; ".join(f"{n} {a[:, i].mean():.1f}" for i, n in enumerate(names)) + " (us, mean over passes)"
if rank == 0:
    p = npass - 2
    st = (t[p, :, 3] - np.maximum(t[p, :, 1], t[p, :, 2])).reshape(-1, slab.solver.plan["nstrips"]) / 1e3   # [chunk][strip]
    out += "\n  stream time per chunk (rows) x strip (cols), us, pass %d:\n" % p + "\n".join("   " + " ".join(f"{x:5.0f}" for x in row) for row in st)
    ex = (t[p, :, 5] - t[p, :, 3]).reshape(-1, slab.solver.plan["nstrips"]) / 1e3
    out += "\n  exit - stream end per CTA, us:\n" + "\n".join("   " + " ".join(f"{x:5.1f}" for x in row) for row in ex)
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    print("\n".join(gathered), flush=True)
dist.destroy_process_group()
