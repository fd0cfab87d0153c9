"""Summarise an `ncu --page raw --csv` export: one block per profiled kernel with the metrics the
roofline discussion needs (duration, DRAM bytes, throughputs, occupancy, top stall reasons)."""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    # raw page: header row, units row, then one row per kernel
    hdr = rows[0]
    units = rows[1]
    return hdr, units, rows[2:]


KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def main(path):
    hdr, units, rows = load(path)
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and not h.endswith("_not_issued")]
    for r in rows:
        print("=" * 100)
        for k in KEYS:
            if k in idx:
                print(f"{k:75s} {r[idx[k]]:>18s} {units[idx[k]]}")
        st = []
        for h in stall:
            try:
                st.append((float(r[idx[h]].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
        tot = sum(v for v, _ in st) or 1.0
        print("stall samples: " + ", ".join(f"{n}={v / tot:.0%}" for v, n in sorted(st, reverse=True)[:8]))


if __name__ == "__main__":
    main(sys.argv[1])
