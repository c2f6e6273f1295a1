"""Extracts the judged metrics from an .ncu-rep (raw page) into a short text table."""
import csv, subprocess, sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "derived__lts__lts2xbar_bytes.sum.per_second"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    name_i = hdr.index("Kernel Name")
    for row in r[2:]:
        print(row[name_i][:110])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"    {w:75s} {row[i]:>16s} {units[i]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
