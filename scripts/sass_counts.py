"""Opcode counts per kernel from the built objects (cuobjdump -sass), written to profiles/ as evidence of which hardware
paths the kernels use: UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor loads), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit),
UBLKCP (cp.async.bulk), DFMA/DMUL/DADD (FP64 pipe), MUFU.RCP64H, SHFL, ATOM...   Run in the build container (no GPU):
    python scripts/sass_counts.py profiles/r2_sass_opcode_counts.md"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU.RCP64H", "MUFU.EX2", "FFMA", "SHFL",
        "LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOM", "RED", "BAR"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, stdout=subprocess.PIPE, text=True).stdout.splitlines()
    return [re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", o).split("(")[0].replace("void ", "") for o in out]


def main(out):
    rows = []
    for obj in sorted(glob.glob(os.path.join(ROOT, "bitswap_b200", "build", "*.o"))):
        txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
        cur, counts = None, None
        for line in txt.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                if cur:
                    rows.append((os.path.basename(obj), cur, counts))
                cur, counts = m.group(1), collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and cur:
                op = m.group(1)
                counts["total"] += 1
                for k in KEYS:
                    if op == k or op.startswith(k + "."):
                        counts[k] += 1
        if cur:
            rows.append((os.path.basename(obj), cur, counts))
    names = demangle([r[1] for r in rows])
    with open(out, "w") as f:
        f.write("# SASS opcode counts per kernel (cuobjdump -sass of bitswap_b200/build/*.o, sm_100a)\n\n"
                "Static instruction counts (one per SASS line, not execution counts).  UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, "
                "LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, DFMA/DMUL/DADD = FP64 pipe.\n\n"
                "| object | kernel | total | " + " | ".join(KEYS) + " |\n|---|---|---|" + "---|" * len(KEYS) + "\n")
        for (obj, _, c), n in zip(rows, names):
            if c["total"] < 24:
                continue
            f.write(f"| {obj} | `{n}` | {c['total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |\n")
    print("wrote", out, len(rows), "kernels")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_opcode_counts.md"))
