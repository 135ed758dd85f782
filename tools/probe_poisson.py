"""Device-time probe of the Poisson pass kernels: fixed sweeps on an nrows x ncols grid for several temporal block
depths and plans.  Prints cell-updates/s and the fraction of the 24 B/cell HBM roofline.
python tools/probe_poisson.py NROWS NCOLS "SPEC,SPEC,..." [sweeps]
SPEC = T:WS:CHUNKS (streaming kernel; 0 = planner's choice) or oT:NTX:NTY (persistent on-chip kernel; 0 = planner's choice)"""
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


def run_onchip(nr, nc, T, ntx, nty, sweeps=256, reps=3):
    os.environ.update(CNV_POISSON_ONCHIP="1", CNV_ONCHIP_T=str(T), CNV_ONCHIP_NTX=str(ntx), CNV_ONCHIP_NTY=str(nty))
    prof = os.environ.get("CNV_ONCHIP_PROF", "0") == "1"
    s = fd.PoissonSolver(nr, nc, 0)
    os.environ["CNV_POISSON_ONCHIP"] = "0"
    if not s.plan["onchip"]:
        print(f"{nr}x{nc} onchip T={T} tiles={ntx}x{nty}: no plan", flush=True)
        s.close()
        return
    p = s.plan
    T, ntx, nty = p["oc_T"], p["oc_ntx"], p["oc_nty"]
    s.set_consts(1.0 / nc, 1.0 / nc, fd.sor_beta(nc, nc))
    s.upload(np.random.default_rng(0).standard_normal((nr, nc)))
    best = 1e9
    for _ in range(reps + 1):
        L.cnv_device_synchronize()
        t0 = time.perf_counter()
        r = s.solve(sweeps, 0.0)
        best = min(best, time.perf_counter() - t0)
    assert r["sweeps"] == sweeps, r
    cu = (nr - 2) * (nc - 2) * sweeps / best
    print(f"{nr}x{nc} onchip T={T} tiles={ntx}x{nty} thr={p['oc_NPX']}x{p['oc_NPY']} out={p['oc_OH']}x{p['oc_OW']} {best/sweeps*1e6:8.2f} us/sweep {cu:.3e} cu/s frac={cu*24/peak:.3f} (whole solve incl. launch + read-back)", flush=True)
    if prof:
        import ctypes as C
        n = ntx * nty
        buf = (C.c_ulonglong * (8 * n))()
        got = L.cnv_poisson_onchip_profile(s.h, buf, n)
        a = np.array(list(buf), dtype=np.float64).reshape(n, 8)[:got]
        npass = (sweeps + T - 1) // T
        names = ("C:sweeps+band", "C:wait-A", "C:reduce+interior", "C:wait-D+halo", "S:publish", "S:wait-counter", "S:copy+sum+fold", "S:waitA+release+poll+waitD")
        print("   ticks per pass (mean over CTAs / max CTA): " + "  ".join(f"{nm} {a[:, i].mean()/npass:.0f}/{a[:, i].max()/npass:.0f}" for i, nm in enumerate(names)), flush=True)
        crit = int(np.argmin(a[:, 3]))   # the CTA that waits least for its neighbours sets the pace
        print(f"   critical CTA {crit} (tile {crit % ntx},{crit // ntx}): " + "  ".join(f"{nm} {a[crit, i]/npass:.0f}" for i, nm in enumerate(names)), flush=True)
        if os.environ.get("CNV_ONCHIP_PROF_ALL"):
            for c in range(got):
                print(f"     cta {c:3d} ({c % ntx:2d},{c // ntx:2d}) " + " ".join(f"{a[c, i]/npass:7.0f}" for i in range(8)), flush=True)
    s.close()


if __name__ == "__main__":
    nr, nc = int(sys.argv[1]), int(sys.argv[2])
    specs = sys.argv[3] if len(sys.argv) > 3 else "1:0:0,2:0:0,4:0:0,8:0:0"
    sweeps = int(sys.argv[4]) if len(sys.argv) > 4 else 256
    for spec in specs.split(","):
        try:
            if spec.startswith("o"):
                T, ntx, nty = (int(x) for x in spec[1:].split(":"))
                run_onchip(nr, nc, T, ntx, nty, sweeps)
                continue
            T, ws, ch = (int(x) for x in spec.split(":"))
            run(nr, nc, T, ws, ch, sweeps)
        except Exception as e:
            print("skip", spec, e)
