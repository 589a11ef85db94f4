"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md):
tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, legacy mma.sync -> HMMA,
plus MUFU.EX2 (softmax) and the register / spill figures ptxas reported.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt      (no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "voice-activity-detection_b200", "lib", "libvadb200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2",
       "F2FP", "STL", "LDL"]


def short_name(mangled):
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    name = name.replace("(anonymous namespace)::", "").replace("vadb::", "").replace("void ", "")
    return re.sub(r"\(.*", "", name)


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = short_name(m.group(1))
            order.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        ins = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not ins:
            continue
        op = ins.group(1)
        counts[cur]["_total"] += 1
        for p in PAT:
            if op.startswith(p):
                counts[cur][p] += 1
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); regenerate: python tools/sass_summary.py")
    print(f"# {'kernel':58s} {'instrs':>7s} " + " ".join(f"{p:>8s}" for p in PAT))
    for k in order:
        c = counts[k]
        print(f"{k[:60]:60s} {c['_total']:7d} " + " ".join(f"{c[p]:8d}" for p in PAT))
    tot = collections.Counter()
    for k in order:
        tot.update(counts[k])
    print(f"{'TOTAL':60s} {tot['_total']:7d} " + " ".join(f"{tot[p]:8d}" for p in PAT))
    # registers / spills as reported by ptxas at build time
    logdir = os.path.join(ROOT, "voice-activity-detection_b200", "csrc", "build")
    print("\n# ptxas -v (registers, spills) per entry point")
    for f in sorted(os.listdir(logdir)):
        if not f.endswith(".ptxas.log"):
            continue
        txt = open(os.path.join(logdir, f)).read()
        for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
            name = short_name(m.group(1))
            print(f"{f[:-10]:14s} {name[:70]:70s} regs {m.group(5):>3s}  spill st/ld {m.group(3)}/{m.group(4)} B")


if __name__ == "__main__":
    sys.exit(main())
