"""Per-CTA timing of the streaming pass kernel on one GPU (cnv_poisson_peer_trace works without peers too):
python tools/trace_single.py [ROWS COLS PASSES]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluid_dynamics1_b200 as fd
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
npass = int(sys.argv[3]) if len(sys.argv) > 3 else 16
L = fd.lib()
s = fd.PoissonSolver(rows, cols, 8)
s.set_consts(1.0 / cols, 1.0 / cols, fd.sor_beta(cols, cols))
s.upload(np.random.default_rng(0).standard_normal((rows, cols)))
for rep in range(2):
    ctas = L.cnv_poisson_peer_trace(s.h, npass)
    s.reset(npass * s.T, 0.0); s.enqueue(npass)
    buf = np.zeros(npass * ctas * 6, dtype=np.uint64)
    L.cnv_poisson_peer_trace_read(s.h, buf.ctypes.data, buf.size)
t = buf.reshape(npass, ctas, 6).astype(np.int64)
p = npass - 2
ns = s.plan["nstrips"]
st = (t[p, :, 3] - t[p, :, 1]).reshape(-1, ns) / 1e3
print("plan", {k: s.plan[k] for k in ("WS", "Hout", "nstrips", "nchunks", "threads")})
print("period us:", (t[p, :, 0].min() - t[p - 1, :, 0].min()) / 1e3, " span:", (t[p, :, 5].max() - t[p, :, 0].min()) / 1e3,
      " gap:", (t[p, :, 0].min() - t[p - 1, :, 5].max()) / 1e3, " start spread:", (t[p, :, 0].max() - t[p, :, 0].min()) / 1e3)
for q in (2, npass // 4, npass // 2, npass - 2):
    sq = (t[q, :, 3] - t[q, :, 1]) / 1e3
    print(f"  pass {q}: period {(t[q, :, 0].min() - t[q - 1, :, 0].min()) / 1e3:.1f} us, stream mean {sq.mean():.1f} max {sq.max():.1f}, "
          f"gap {(t[q, :, 0].min() - t[q - 1, :, 5].max()) / 1e3:.1f}")
print("stream time [chunk][strip] us:")
for row in st: print("  " + " ".join(f"{x:5.0f}" for x in row))
ex = (t[p, :, 5] - t[p, :, 3]).reshape(-1, ns) / 1e3
print("exit - stream end: mean %.1f max %.1f" % (ex.mean(), ex.max()))
