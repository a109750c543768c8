"""SASS opcode histogram of the shipped library, per kernel family: which tensor / TMA / async-copy instructions each
kernel class actually contains.  `python profiles/sass_opcodes.py > profiles/sass_opcodes_r02.txt` (needs cuobjdump)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bayesian_quadrature_b200", "libbq_b200.so")
WATCH = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "UTCMMA", "UTCHMMA", "LDTM", "HMMA", "IMMA",
         "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "REDUX", "ATOM", "RED"]


def main():
    p = subprocess.Popen(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True)
    fam = None
    hist = collections.defaultdict(collections.Counter)
    nk = collections.Counter()
    for line in p.stdout:
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            name = re.sub(r"^void ", "", name)
            fam = re.sub(r"\(.*", "", name)
            fam = re.sub(r"<.*", "<...>", fam) if fam.count(",") > 3 else fam
            nk[fam] += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and fam:
            op = m.group(1)
            hist[fam][op] += 1
            hist[fam]["_total"] += 1
    p.wait()
    print("# SASS opcode histogram of libbq_b200.so (sm_100a), per kernel family: instances | total instructions | watched opcodes")
    print("# tcgen05 (UTC*MMA / LDTM) has no f64 kind: DMMA.8x8x4 is Blackwell's FP64 tensor path; UBLKCP = cp.async.bulk (TMA bulk")
    print("# copy of the operand fragments), SYNCS = mbarrier, LDGSTS = cp.async")
    for fam in sorted(hist, key=lambda f: -hist[f]["_total"]):
        h = hist[fam]
        parts = ["%s %d" % (w, sum(v for k, v in h.items() if k.split(".")[0].split("_")[0] == w or k.startswith(w + "."))) for w in WATCH]
        parts = [q for q in parts if not q.endswith(" 0")]
        print("%-60s kernels %3d  instr %8d  | %s" % (fam[:60], nk[fam], h["_total"], ", ".join(parts)))


if __name__ == "__main__":
    sys.exit(main())
