"""Print selected metrics of every kernel in an `ncu -i X.ncu-rep --page raw --csv` dump (read from a file)."""
import csv
import sys

PICK = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_uniform",
        "sm__warps_active.avg.pct", "launch__registers_per_thread", "sm__throughput.avg.pct", "gpu__dram_throughput.avg.pct", "lts__t_bytes.sum ",
        "lts__t_sectors_srcunit_tex_op_read.sum", "launch__waves_per_multiprocessor", "launch__occupancy_limit", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "smsp__cycles_active.avg", "lts__t_sector_hit_rate", "sm__ctas_launched", "tensor"]


def main(path, extra):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    print("kernels:", [(r[ki][:30], r[gi]) for r in data])
    for i, h in enumerate(hdr):
        if any(p in h for p in PICK + extra):
            print(f"{h[:95]:95s} [{units[i]:>12s}] " + " | ".join(r[i] for r in data))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
