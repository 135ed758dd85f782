"""Device-time probe of the Poisson pass kernels: fixed sweeps on an nrows x ncols grid for several temporal block
depths and plans.  Prints cell-updates/s and the fraction of the 24 B/cell HBM roofline.
python tools/probe_poisson.py NROWS NCOLS "SPEC,SPEC,..." [sweeps]
SPEC = T:WS:CHUNKS (streaming kernel; 0 = planner's choice)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluid_dynamics1_b200 as fd

L = fd.lib()
peak = 6556.2e9
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
except Exception:
    pass


def run(nr, nc, T, ws=0, chunks=0, sweeps=256, reps=3):
    os.environ["CNV_POISSON_WS"] = str(ws); os.environ["CNV_POISSON_CHUNKS"] = str(chunks)
    s = fd.PoissonSolver(nr, nc, T)
    s.set_consts(1.0 / nc, 1.0 / nc, fd.sor_beta(nc, nc))
    rng = np.random.default_rng(0)
    s.upload(rng.standard_normal((nr, nc)))
    best = 1e9
    for _ in range(reps):
        s.reset(sweeps, 0.0)
        L.cnv_device_synchronize()
        t0 = time.perf_counter()
        s.enqueue((sweeps + T - 1) // T)
        L.cnv_device_synchronize()
        best = min(best, time.perf_counter() - t0)
    st = s.state()
    assert st["sweeps"] == sweeps, st
    cu = (nr - 2) * (nc - 2) * sweeps / best
    p = s.plan
    print(f"{nr}x{nc} T={T} WS={p['WS']} Hout={p['Hout']} ctas={p['nstrips']}x{p['nchunks']} thr={p['threads']} smem={p['smem']//1024}K "
          f"{best/sweeps*1e6:8.2f} us/sweep {cu:.3e} cu/s frac={cu*24/peak:.3f}", flush=True)
    s.close()


if __name__ == "__main__":
    nr, nc = int(sys.argv[1]), int(sys.argv[2])
    specs = sys.argv[3] if len(sys.argv) > 3 else "1:0:0,2:0:0,4:0:0,8:0:0"
    sweeps = int(sys.argv[4]) if len(sys.argv) > 4 else 256
    for spec in specs.split(","):
        try:
            T, ws, ch = (int(x) for x in spec.split(":"))
            run(nr, nc, T, ws, ch, sweeps)
        except Exception as e:
            print("skip", spec, e)
