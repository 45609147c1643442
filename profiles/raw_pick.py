"""Print selected metrics per kernel launch from an `ncu --page raw --csv` export."""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def main(path, extra=()):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "smsp__average_warp" in h and "issue_stalled" in h and h.endswith("_per_warp_active.pct")]
    stall2 = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and "not_issued" not in h]
    for r in rows[2:]:
        print("=====", r[idx["Kernel Name"]])
        for w in list(WANT) + list(extra):
            if w in idx:
                print(f"  {w:72s} {r[idx[w]]} {units[idx[w]]}")
        st = []
        for h in stall or stall2:
            try:
                st.append((float(r[idx[h]]), h))
            except ValueError:
                pass
        for v, h in sorted(st, reverse=True)[:8]:
            print(f"  STALL {h:80s} {v:.2f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
