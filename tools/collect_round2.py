"""Turn the outputs of tools/round2_gpu.sh (gpurun_out/r2_*) into one markdown report:
    python tools/collect_round2.py [gpurun_out] > profiles/ab_r2.md
Every section is optional: whatever files exist are reported, missing ones are listed as not run."""
import glob
import json
import os
import re
import sys

D = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"


def last_json(path):
    """The bench line = the last line of the file that parses as a JSON object."""
    if not os.path.exists(path):
        return None
    for ln in reversed(open(path, errors="replace").read().splitlines()):
        ln = ln.strip()
        if ln.startswith("{") and ln.endswith("}"):
            try:
                return json.loads(ln)
            except json.JSONDecodeError:
                continue
    return None


def tail(path, n=3):
    if not os.path.exists(path):
        return ["(not run)"]
    return open(path, errors="replace").read().splitlines()[-n:]


def bench_row(label, j, base=None):
    if j is None:
        return f"| {label} | not run | | | | |"
    r = j.get("roofline", {})
    rel = f"{j['value'] / base['value']:.3f}" if base and base.get("value") else ""
    clocks = j.get("clocks") or {}
    return (f"| {label} | {j['value']:.4e} | {j['ms_per_step']:.2f} | {r.get('launch_us', float('nan')):.1f} | {r.get('frac', float('nan')):.3f} | "
            f"{rel} | {clocks.get('sm_mhz')} {','.join(clocks.get('reasons') or [])} |")


def section_single():
    out = ["## Single GPU, 4096², 1024 sweeps per step (`round2_gpu.sh single`)", "",
           "| variant | cell-updates/s | ms/step | µs per pass | roofline frac | vs default | clocks |", "|---|---|---|---|---|---|---|"]
    base = last_json(os.path.join(D, "r2_bench_n1.json"))
    again = last_json(os.path.join(D, "r2_bench_n1_again.json"))
    out.append(bench_row("default (first run, with cpu_baseline)", base))
    out.append(bench_row("default (again, right after the lean run)", again, base))
    out.append(bench_row("CNV_POISSON_EDGE=0 (equal chunks, the plan measured in round 1)", last_json(os.path.join(D, "r2_bench_n1_edge0.json")), again or base))
    out.append(bench_row("CNV_LIB=lean (leaner step body)", last_json(os.path.join(D, "r2_bench_n1_lean.json")), again or base))
    if base:
        for key in ("e2e", "stencil_phase", "timestep_1024", "timestep_4096", "cpu_baseline"):
            if key in base:
                out.append("")
                out.append(f"`{key}`: `{json.dumps(base[key])}`")
    out += ["", "GPU parity suite: `" + " / ".join(tail(os.path.join(D, "r2_pytest_gpu.log"), 2)) + "`", "",
            "Parity subset on the lean library: `" + " / ".join(tail(os.path.join(D, "r2_pytest_lean.log"), 2)) + "`", ""]
    return out


PROBE = re.compile(r"^(\d+)x(\d+) T=(\d+) (tile|WS)(.*?)\s+([\d.]+) us/sweep\s+([\d.e+]+) cu/s frac=([\d.]+)")


def section_probes():
    out = ["## Stationary-tile kernel vs streaming kernel (`tools/probe_poisson.py`, µs per sweep)", ""]
    rows = {}
    for fn, tag in (("r2_probe_tile.log", ""), ("r2_probe_tile_pdl.log", " +PDL")):
        path = os.path.join(D, fn)
        if not os.path.exists(path):
            continue
        for ln in open(path, errors="replace"):
            m = PROBE.match(ln.strip())
            if not m:
                continue
            nr, nc, T, kind, rest, us, cu, frac = m.groups()
            name = ("tile" if kind == "tile" else "stream") + tag
            rows.setdefault((int(nr), int(nc)), []).append((name, int(T), float(us), float(cu), float(frac), (rest if kind == "tile" else "WS" + rest).strip()))
    if not rows:
        return out + ["(not run)", ""]
    out += ["| grid | kernel | T | µs/sweep | cell-updates/s | frac of 24 B/cell roofline | plan |", "|---|---|---|---|---|---|---|"]
    for shape in sorted(rows):
        best = min(r[2] for r in rows[shape])
        for name, T, us, cu, frac, rest in sorted(rows[shape], key=lambda r: r[2]):
            mark = " **best**" if us == best else ""
            out.append(f"| {shape[0]}×{shape[1]} | {name}{mark} | {T} | {us:.2f} | {cu:.3e} | {frac:.3f} | {rest[:70]} |")
    return out + [""]


def section_multi():
    out = ["## Multi-GPU (`round2_gpu.sh lagtest|scale|extra N`)", ""]
    ns = sorted({int(m.group(1)) for f in glob.glob(os.path.join(D, "r2_scale_peer_*.json")) for m in [re.search(r"_(\d+)\.json$", f)] if m})
    if not ns:
        return out + ["(not run)", ""]
    single = last_json(os.path.join(D, "r2_bench_n1_again.json")) or last_json(os.path.join(D, "r2_bench_n1.json"))
    out += ["| GPUs | run | cell-updates/s | ms/step | µs per pass | weak-scaling efficiency vs 1 GPU | vs plain peer |", "|---|---|---|---|---|---|---|"]
    for n in ns:
        plain = last_json(os.path.join(D, f"r2_scale_peer_{n}.json"))
        for label, fn in (("weak, peer (plain machine)", f"r2_scale_peer_{n}.json"), ("weak, peer + lagged decision", f"r2_scale_lag_{n}.json"),
                          ("strong 4096², peer", f"r2_strong_peer_{n}.json"), ("strong 4096², peer + lagged", f"r2_strong_lag_{n}.json"),
                          ("strong 4096², nccl, streaming kernel", f"r2_strong_nccl_{n}.json"),
                          ("strong 4096², nccl, tile kernel T=4", f"r2_strong_tile_T4_{n}.json"),
                          ("strong 4096², nccl, tile kernel T=8", f"r2_strong_tile_T8_{n}.json"),
                          ("strong 4096², peer, tile kernel T=4", f"r2_strong_tilepeer_T4_{n}.json"),
                          ("strong 4096², peer, tile kernel T=8", f"r2_strong_tilepeer_T8_{n}.json")):
            j = last_json(os.path.join(D, fn))
            if j is None:
                out.append(f"| {n} | {label} | not run | | | | |")
                continue
            eff = f"{j['value'] / (n * single['value']):.3f}" if single and label.startswith("weak") else ""
            ref = plain if label.startswith("weak") else last_json(os.path.join(D, f"r2_strong_peer_{n}.json"))
            rel = f"{j['value'] / ref['value']:.3f}" if ref else ""
            out.append(f"| {n} | {label} | {j['value']:.4e} | {j['ms_per_step']:.2f} | {j.get('roofline', {}).get('launch_us', float('nan')):.1f} | {eff} | {rel} |")
        out.append(f"| {n} | lagged GPU tests | `{' / '.join(tail(os.path.join(D, f'r2_lag_tests_{n}.log'), 2))}` | | | | |")
        out.append(f"| {n} | tile-kernel peer tests | `{' / '.join(tail(os.path.join(D, f'r2_tile_peer_tests_{n}.log'), 2))}` | | | | |")
        out.append(f"| {n} | plain multi-GPU suite | `{' / '.join(tail(os.path.join(D, f'r2_multi_tests_{n}.log'), 2))}` | | | | |")
        out.append(f"| {n} | tile-kernel slab parity | `{' / '.join(tail(os.path.join(D, f'r2_tile_slab_check_{n}.log'), 1))}` | | | | |")
    for n in ns:
        for tag in ("peer", "lag"):
            p = os.path.join(D, f"r2_trace_{tag}_{n}.log")
            if os.path.exists(p):
                lines = [ln for ln in open(p, errors="replace").read().splitlines() if ln.startswith("rank 0/")]
                if lines:
                    out += ["", f"Trace, {n} GPUs, {tag}: `{lines[0][:600]}`"]
    return out + [""]


def main():
    print("# Round 2 — A/B results of the variants prepared at the end of round 1\n")
    print("Generated by `tools/collect_round2.py` from the outputs of `tools/round2_gpu.sh`.\n")
    for sec in (section_single, section_probes, section_multi):
        print("\n".join(sec()))


if __name__ == "__main__":
    main()
