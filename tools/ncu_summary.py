"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# summary of {rep} (ncu --set full --clock-control none)\n")
        for r in rows[2:]:
            f.write(f"\nkernel: {r[hdr.index('Kernel Name')]}\n")
            for w in WANT:
                if w in hdr:
                    f.write(f"  {w:75s} {r[hdr.index(w)]} {units[hdr.index(w)]}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
