"""Static SASS opcode histogram of every kernel in libcfp_b200.so (cuobjdump -sass), as a markdown table: which
instruction families the compiled code consists of -- in particular DMMA (FP64 tensor core), DFMA/DMUL/DADD (FP64
vector), LDS/STS, LDG/STG, and that no UTMALDG/UBLKCP/tcgen05 (TMA / 5th-gen tensor core) appear: FP64 has no tcgen05
kind, the lattice kernels keep their working set in registers and shared memory.
Usage: python tools/sass_histogram.py [library.so] > profiles/rN_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHOW = ["DMMA", "DFMA", "DMUL", "DADD", "DSETP", "MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL", "LDC", "SHFL", "BAR",
        "BRA", "IMAD", "ISETP", "FSEL", "LOP3", "UTMALDG", "UBLKCP", "UTCMMA"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "cosmicfishpie_b200", "libcfp_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    hist, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            hist[cur][m.group(1)] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode histogram per kernel (static instruction counts, `cuobjdump -sass`, sm_100a)\n")
    print("| kernel | total | " + " | ".join(SHOW) + " |")
    print("|---|---|" + "---|" * len(SHOW))
    for (name, c), dn in zip(hist.items(), demangled):
        short = dn.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        short = re.sub(r"\((?!int\)|bool\)).*", "", short)  # drop the argument list, keep template arguments
        print(f"| `{short}` | {sum(c.values())} | " + " | ".join(str(c.get(k, 0)) for k in SHOW) + " |")
    tot = collections.Counter()
    for c in hist.values():
        tot.update(c)
    print(f"\nLibrary totals: DMMA {tot['DMMA']}, DFMA {tot['DFMA']}, UTMALDG {tot['UTMALDG']}, UBLKCP {tot['UBLKCP']}, "
          f"UTCMMA (tcgen05.mma) {tot['UTCMMA']}.")


if __name__ == "__main__":
    main()
