"""Quick device-time probe of the streaming Poisson kernel: fixed sweeps at a given size for several
temporal block depths / strip widths.  Prints cell-updates/s and the fraction of the 24 B/cell HBM roofline."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluid_dynamics1_b200 as fd

L = fd.lib()
peak = 6556.2e9
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
except Exception:
    pass

def run(n, T, ws=0, chunks=0, sweeps=256, reps=3):
    os.environ["CNV_POISSON_WS"] = str(ws); os.environ["CNV_POISSON_CHUNKS"] = str(chunks)
    s = fd.PoissonSolver(n, n, T)
    s.set_consts(1.0 / n, 1.0 / n, fd.sor_beta(n, n))
    rng = np.random.default_rng(0)
    s.upload(rng.standard_normal((n, n)))
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
    cu = (n - 2) ** 2 * sweeps / best
    print(f"n={n} T={T} plan={ {k: s.plan[k] for k in ('WS','Hout','nstrips','nchunks','threads','smem','pow2')} } "
          f"{best/sweeps*1e6:8.2f} us/sweep  {cu:.3e} cell-updates/s  roofline_frac={cu*24/peak:.3f}", flush=True)
    s.close()

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    for T in (1, 2, 4, 8):
        run(n, T)
    for T, ws in ((4, 128), (4, 256), (4, 384), (4, 512), (2, 256), (2, 512), (2, 1024), (8, 256), (8, 384)):
        for ch in (0,):
            try:
                run(n, T, ws, ch)
            except SystemExit:
                raise
            except Exception as e:
                print("skip", T, ws, e)
