"""Top stalled SASS instructions (with a few lines of context) from an `ncu --page source --csv` export."""
import csv
import sys


def main(path, top=12, ctx=3):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[hi]
    i_s, i_e, i_n = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.startswith("Stall")]
    data = [r for r in rows[hi + 1:] if len(r) > i_n and r[i_n].isdigit()]
    tot = sum(int(r[i_n]) for r in data)
    order = sorted(range(len(data)), key=lambda k: -int(data[k][i_n]))[:top]
    for k in order:
        r = data[k]
        reasons = sorted(((int(r[c]) if r[c].isdigit() else 0, hdr[c]) for c in stall_cols), reverse=True)[:3]
        print(f"--- #{k} samples {100.0 * int(r[i_n]) / tot:.1f}%  exec {r[i_e]}  top reasons {[(n, h) for n, h in reasons if n]}")
        for j in range(max(0, k - ctx), min(len(data), k + 2)):
            mark = ">>" if j == k else "  "
            print(f"   {mark} {data[j][i_s].strip()[:110]}   [{data[j][i_n]}]")


if __name__ == "__main__":
    main(sys.argv[1], *(int(a) for a in sys.argv[2:]))
