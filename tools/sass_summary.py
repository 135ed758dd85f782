"""Instruction histogram of selected kernels from the built library (no GPU needed):
python tools/sass_summary.py > profiles/sass_r2.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "fluid_dynamics1_b200", "libcnavier_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = {}
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        funcs[cur].append(m.group(1))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
want = sys.argv[1:] or ["k_poisson_pass<8, true, false>", "k_poisson_pass<8, true, true>", "k_poisson_onchip<true>", "k_euler_fused<3>",
                        "k_velocity<3>", "k_continuity<3>"]
print("# SASS of the hot kernels (cuobjdump -sass of libcnavier_b200.so, sm_100a; `tools/sass_summary.py`)\n")
print("No tensor-core instruction anywhere (`UTC*MMA`, `HMMA`: 0) -- nothing on this path is a dense contraction.  fp64 arithmetic is "
      "`DADD` / `DMUL` only (separately rounded, like the reference's x86 build); the few `DFMA` belong to the exact constant-divisor "
      "division of the general (non power-of-two spacing) path.\n")
for name, ins in funcs.items():
    d = demangle(name).replace("void cnv::", "").replace("cnv::", "")
    if not any(w in d for w in want):
        continue
    h = collections.Counter(ins)
    grp = collections.Counter()
    for k, v in h.items():
        grp[k.split(".")[0]] += v
    tot = len(ins)
    keys = ["DADD", "DMUL", "DFMA", "LDS", "STS", "LDGSTS", "LDG", "STG", "BAR", "SHFL", "MEMBAR", "ATOM", "ATOMG", "RED", "CCTL", "IMAD", "MOV", "FSEL", "SEL", "ISETP", "BRA", "UTCHMMA", "HMMA", "UTMALDG"]
    print(f"## `{d}` -- {tot} instructions\n")
    print("| " + " | ".join(k for k in keys if grp[k] or k in ("DFMA", "HMMA", "UTMALDG")) + " |")
    print("|" + "---|" * len([k for k in keys if grp[k] or k in ("DFMA", "HMMA", "UTMALDG")]))
    print("| " + " | ".join(str(grp[k]) for k in keys if grp[k] or k in ("DFMA", "HMMA", "UTMALDG")) + " |")
    wide = {k: v for k, v in h.items() if re.search(r"\.(64|128|256)|\.SYS|\.GPU|ENL2", k) and k.split(".")[0] in ("LDS", "STS", "LDG", "STG", "LDGSTS", "MEMBAR", "ATOM", "ATOMG", "RED", "ST", "LD")}
    print("\nmemory / synchronisation forms: " + ", ".join(f"`{k}` x{v}" for k, v in sorted(wide.items(), key=lambda kv: -kv[1])[:18]) + "\n")
