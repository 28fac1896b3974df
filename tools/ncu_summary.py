"""Text summary of an ``.ncu-rep`` capture (run where ``ncu`` is installed; no GPU needed):

    python tools/ncu_summary.py profiles/ncu/attn_bwd_r2.ncu-rep > profiles/ncu/attn_bwd_r2_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "smsp__cycles_elapsed.avg.per_second", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum")


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    rows = [r for r in rows if len(r) > 5]
    hdr, units, vals = rows[0], rows[1], rows[2]
    print(f"# {path}")
    for i, h in enumerate(hdr):
        if h in ("Kernel Name", "Block Size", "Grid Size") or h in KEYS or ("tensor" in h and ("pct" in h or h.endswith(".sum"))):
            if vals[i] not in ("0", "0.000000", ""):          # dozens of unused tensor sub-pipes read zero
                print(f"{h} [{units[i]}] = {vals[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
