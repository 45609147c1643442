"""Join an `ncu --page source --csv` export (per-SASS-instruction counts) with `nvdisasm -g -c` line info of the same
cubin: executed warp-instructions and stall samples per SOURCE line (innermost inlined location).

usage: line_hist.py <ncu_source.csv> <nvdisasm.txt> <mangled-function-substring> [top]
"""
import collections
import csv
import re
import sys


def parse_disasm(path, func):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and func in l and l.endswith(":"))
    loc, cur = {}, ("?", 0)
    for l in lines[start + 1:]:
        if l.startswith("//-----"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).rsplit("/", 1)[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            loc[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return loc


def main(csv_path, dis_path, func, top=60):
    loc = parse_disasm(dis_path, func)
    rows = list(csv.reader(open(csv_path)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[hi]
    i_a, i_s, i_e, i_n = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = []
    for r in rows[hi + 1:]:
        try:
            data.append((int(r[i_a], 16), r[i_s].strip(), int(r[i_e]), int(r[i_n])))
        except (ValueError, IndexError):
            continue
    base = data[0][0]
    ex, sm, mism = collections.Counter(), collections.Counter(), 0
    for a, s, e, n in data:
        l = loc.get(a - base)
        if l is None:
            mism += 1
            continue
        ex[l[0]] += e
        sm[l[0]] += n
    tot_e, tot_s = sum(ex.values()), sum(sm.values())
    print(f"instructions {len(data)} (unmatched {mism}), executed {tot_e}, samples {tot_s}")
    for k, e in ex.most_common(top):
        print(f"{k[0]}:{k[1]:<5d} exec {100.0 * e / tot_e:5.1f}%  samples {100.0 * sm[k] / max(tot_s, 1):5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 60)
