"""Generate the committed golden fixtures from the reference itself.  Run once, in the build
container (needs /root/reference and oracle/_ref built by `make -C oracle`):

    python tests/golden/make_golden.py

Outputs (all small, committed):
  ref_logs.json            per-step Poisson sweep counts (logged k = sweeps-1), 7-digit residual strings and the
                           7-digit "Continuity max / min" strings parsed from the reference's shipped run logs
                           (/root/reference/output/logs/testRun{NOOMP,OMP,NOOMPHIGHRES,OMPHIGHRES}.txt)
  fields_default_rb.npz    psi,w,u,v (raw fp64) after steps 0, 10, 20 of config_default.txt run by the
                           unmodified reference built with -DOPENMP_ENABLED (red-black SOR), + its k/e log
  fields_highre_rb.npz     psi,w,u,v after step 5 of config_high_re.txt (128^2), same build, + k/e log
  fields_default_lex_tight.npz  psi,w,u,v after step 2 of config_default.txt with poisson_tol=1e-11 from the
                           serial (lexicographic) reference build: ordering-independent comparison point
  fields_order2_rb.npz, fields_order4_rb.npz   psi,w,u,v after every one of 4 steps of a 32^2 cavity with moving side
                           walls, finite-difference order 2 / 4, red-black SOR (OpenMP build) -- the `order` config key
  fields_gs_lex_tight.npz  the same cavity, order 4, poisson_type = 1 (Gauss-Seidel), poisson_tol = 1e-11, 3 steps, serial
                           build.  The reference's Gauss-Seidel variants sweep LEXICOGRAPHICALLY in both builds
                           (src/poisson.c:80-86, 193-200; the OpenMP pragma merely runs that loop in parallel above 64^2),
                           so the red-black GPU path is compared with it at a tolerance where the ordering is immaterial
  poisson_sine.json        sweep counts of the reference's poisson_SOR_log on f=-2pi^2 sin(pi x) sin(pi y)
"""
import json, os, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import api  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REFLOGS = "/root/reference/output/logs"


def logs():
    res = {}
    for name in ("testRunNOOMP", "testRunOMP", "testRunNOOMPHIGHRES", "testRunOMPHIGHRES"):
        with open(os.path.join(REFLOGS, name + ".txt")) as f:
            pl = api.parse_poisson_log(f.read())
        with open(os.path.join(REFLOGS, name + ".txt")) as f:
            cl = api.parse_continuity_log(f.read())
        res[name] = {"k": [k for k, _ in pl], "e": [e for _, e in pl],
                     "cont_max": [a for a, _ in cl], "cont_min": [b for _, b in cl]}
        print(name, len(pl), "poisson lines,", len(cl), "continuity lines")
    with open(os.path.join(OUT, "ref_logs.json"), "w") as f:
        json.dump(res, f)


def fields(cfg, steps, fname, serial=False):
    # steps run = (int)(tf/dt - 1) + 1 ; pick tf in the middle of the bracket so rounding cannot bite
    cfg = dict(cfg, tf=(steps + 0.5) * cfg["dt"])
    assert api.port().num_steps(cfg["tf"], cfg["dt"]) == steps
    with tempfile.TemporaryDirectory() as td:
        log, d = api.run_reference_binary(cfg, td, serial=serial, threads=8)
    pl = api.parse_poisson_log(log)
    assert len(pl) == steps
    np.savez_compressed(os.path.join(OUT, fname), k=np.array([k for k, _ in pl], dtype=np.int32),
                        e=np.array([float(e) for _, e in pl]), steps=steps,
                        dump_steps=np.arange(0, steps, cfg["output_interval"]), **d)
    print(fname, {k: v.shape for k, v in d.items()}, [k for k, _ in pl][:8])


def sine():
    res = {}
    for n in (64, 128, 256):
        x = np.arange(n) / n
        f = -2 * np.pi ** 2 * np.outer(np.sin(np.pi * x), np.sin(np.pi * x))
        beta = api.port().beta(n, n)
        rb = api.ref().poisson(f, 1.0 / n, 1.0 / n, 20000, 1e-3, beta)
        lex = api.ref(serial=True).poisson(f, 1.0 / n, 1.0 / n, 20000, 1e-3, beta)
        res[str(n)] = {"rb_k": rb["k"], "rb_e": rb["e"], "lex_k": lex["k"], "lex_e": lex["e"],
                       "rb_u_sum": float(rb["u"].sum()), "rb_u_l2": float(np.sqrt((rb["u"] ** 2).sum()))}
    with open(os.path.join(OUT, "poisson_sine.json"), "w") as f:
        json.dump(res, f)
    print(res)


if __name__ == "__main__":
    logs()
    if sys.argv[1:] == ["logs"]:
        sys.exit(0)
    sine()
    fields(dict(api.CONFIG_DEFAULT), 21, "fields_default_rb.npz")
    fields(dict(api.CONFIG_DEFAULT, poisson_tol=1e-11, poisson_max_it=100000, output_interval=2), 3,
           "fields_default_lex_tight.npz", serial=True)
    fields(dict(api.CONFIG_HIGH_RE), 6, "fields_highre_rb.npz")
    small = dict(api.CONFIG_DEFAULT, nx=32, ny=32, dt=0.004, u1=0.05, u2=-0.03, v3=0.02, v4=0.01, output_interval=1)
    fields(dict(small, order=2), 4, "fields_order2_rb.npz")
    fields(dict(small, order=4), 4, "fields_order4_rb.npz")
    fields(dict(small, order=4, poisson_type=1, poisson_tol=1e-11, poisson_max_it=200000), 3, "fields_gs_lex_tight.npz", serial=True)
