"""Aggregate an `ncu --page source --csv` export by opcode: executed warp-instructions per launched warp."""
import collections
import csv
import sys


def main(path, top=45):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[hi]
    i_s, i_e, i_n = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= max(i_s, i_e, i_n):
            continue
        try:
            data.append((r[i_s].strip(), int(r[i_e]), int(r[i_n])))
        except ValueError:
            continue
    warps = max(e for _, e, _ in data)
    tot = sum(e for _, e, _ in data)
    print(f"sass instructions {len(data)}, executed {tot}, launched warps ~{warps}, per warp {tot / warps:.1f}")
    ops, samp = collections.Counter(), collections.Counter()
    for s, e, n in data:
        if s.startswith("@"):
            s = s.split(None, 1)[1]
        op = s.split()[0].split(".")[0]
        ops[op] += e
        samp[op] += n
    tot_s = sum(samp.values())
    for op, c in ops.most_common(top):
        print(f"{op:10s} {c / warps:8.1f} /warp   stall samples {100.0 * samp[op] / max(tot_s, 1):5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
