"""profiles/rNN_sass_summary.md: per kernel of libcetkmc.so the SASS size and the mnemonics that show how
it touches memory (vector / byte loads, reductions, TMA, mbarrier, shared memory) and where its
arithmetic goes (fp64, integer), from `cuobjdump -sass`.  No GPU needed."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cet-driven-simulation-for-3d-printing-am-kmc-approach_b200", "libcetkmc.so")
GROUPS = [
    ("LDG.128", r"^LDG\.E(\.\w+)*\.128"), ("LDG.64", r"^LDG\.E(\.\w+)*\.64"), ("LDG.U8/S8", r"^LDG\.E(\.\w+)*\.[US]8"),
    ("LDGSTS (cp.async)", r"^LDGSTS"), ("LDG other", r"^LDG"), ("STG", r"^STG"), ("RED/ATOMG", r"^(RED|ATOMG|ATOM)\b"), ("ATOMS", r"^ATOMS"),
    ("LDS", r"^LDS"), ("STS", r"^STS"), ("UTMALDG (TMA)", r"^UTMALDG"), ("SYNCS (mbarrier)", r"^SYNCS"),
    ("BAR", r"^BAR"), ("SHFL", r"^SHFL"), ("REDUX", r"^REDUX"), ("DFMA", r"^DFMA"), ("DMUL", r"^DMUL"), ("DADD", r"^DADD"),
    ("DSETP", r"^DSETP"), ("MUFU", r"^MUFU"), ("IMAD*", r"^IMAD"), ("LOP3", r"^LOP3"), ("POPC", r"^POPC"),
]


def main(out):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    kernels = collections.OrderedDict()
    name = None
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            kernels[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m and name:
            ins = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip())
            kernels[name].append(ins)
    with open(out, "w") as f:
        f.write(f"# SASS summary of libcetkmc.so ({', '.join(arch)}; `cuobjdump -sass`, scripts/sass_summary.py)\n\n"
                "Static instruction counts per kernel (not execution counts).  The TMA path shows as `UTMALDG` + `SYNCS`\n"
                "(mbarrier) in `rates_dense_kernel` (the default dense rate kernel) and `rates_tile3d_kernel<0, *>`; the\n"
                "streaming kernels load with `LDG.E.128`; the refresh kernel gathers class codes with `LDG.E.U8` and fetches\n"
                "its pair operands with `LDGSTS` (cp.async, 8 bytes each); the pick kernel gathers with `LDG.E.U8` / `LDG.E.64`.\n\n")
        f.write("| kernel | SASS instr | " + " | ".join(g for g, _ in GROUPS) + " |\n|---|---|" + "---|" * len(GROUPS) + "\n")
        for k, ins in kernels.items():
            if not ins:
                continue
            counts = []
            used = [False] * len(ins)
            for g, pat in GROUPS:
                rx = re.compile(pat)
                n = 0
                for q, i in enumerate(ins):
                    if not used[q] and rx.match(i):
                        n += 1
                        used[q] = True
                counts.append(n)
            f.write(f"| `{k[:70]}` | {len(ins)} | " + " | ".join(str(c) if c else "" for c in counts) + " |\n")
        f.write("\n## Excerpt: the TMA issue of `rates_dense_kernel<true>` (two 3-D boxes on one mbarrier)\n\n```\n")
        tma = [k for k in kernels if k.startswith("void cet::rates_dense_kernel<true>")]
        if tma:
            ins = kernels[tma[0]]
            for q, i in enumerate(ins):
                if i.startswith(("UTMALDG", "SYNCS")):
                    f.write(f"{q:5d}  {i}\n")
        f.write("```\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.md"))
