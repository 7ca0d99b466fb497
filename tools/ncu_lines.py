"""Join an `ncu --page source --csv` dump (per-SASS-instruction counters) with `nvdisasm --print-line-info`
output of the same cubin: instructions executed / stall samples per CUDA source line.
Usage: python tools/ncu_lines.py src.csv disasm.txt kernel_substring source.cu [top_n]"""
import collections
import csv
import re
import sys


def main():
    src_csv, dis, kern, cu = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ia, isrc, ii = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    isamp = hdr.index("# Samples")
    base = None
    counts = {}
    for r in rows[2:]:
        if len(r) <= ii or not r[ii].isdigit():
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        counts[a - base] = (int(r[ii]), int(r[isamp]) if r[isamp].isdigit() else 0, r[isrc])
    # line info
    line_of = {}
    cur, active = None, False
    for ln in open(dis):
        if ln.startswith(".text.") and ln.rstrip().endswith(":"):
            active = kern in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File ".*", line (\d+)', ln)
        if m:
            cur = int(m.group(1))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    per_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot = sum(c[0] for c in counts.values())
    tots = sum(c[1] for c in counts.values())
    for off, (n, smp, sass) in counts.items():
        L = line_of.get(off)
        per_line[L][0] += n
        per_line[L][1] += smp
        op = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", sass)
        per_line[L][2][op.group(2) if op else "?"] += n
    text = open(cu).read().split("\n")
    print(f"total warp instructions {tot}, samples {tots}")
    for L, (n, smp, ops) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        srcl = text[L - 1].strip()[:90] if L else "?"
        opstr = " ".join(f"{k}:{100 * v / n:.0f}" for k, v in ops.most_common(4))
        print(f"{L!s:>5} {100 * n / tot:5.1f}% inst {100 * smp / max(tots, 1):5.1f}% smp | {srcl} | {opstr}")


if __name__ == "__main__":
    main()
