"""Summarise `ncu -i X.ncu-rep --page raw --csv` into a small metric table (one column per captured launch).
usage: python tools/ncu_summary.py raw.csv out.csv "title line" [launch indices...]"""
import csv, sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]

def main():
    raw, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    pick = [int(x) for x in sys.argv[4:]]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if pick:
        data = [data[i] for i in pick]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# {title}\n")
        f.write("metric,unit," + ",".join(f"launch{i}" for i in (pick or range(len(data)))) + "\n")
        for k in KEYS:
            if k in col:
                f.write(f"{k},{units[col[k]]}," + ",".join(r[col[k]].replace(",", "") for r in data) + "\n")

if __name__ == "__main__":
    main()
