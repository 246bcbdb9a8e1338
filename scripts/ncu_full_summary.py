"""Per-kernel digest of an `ncu --set full` report: python scripts/ncu_full_summary.py X.ncu-rep > profiles/....txt
(one block per captured launch: duration, tensor-pipe / issue / occupancy counters, DRAM and L2 traffic)."""
import csv, subprocess, sys
PICK = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
        ("smsp__inst_executed.sum", "warp instructions"), ("sm__cycles_elapsed.max", "SM cycles elapsed")]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
print(f"# {rep}: {len(data)} captured launches (ncu --set full --clock-control none; cold-cache, serialised replays)")
for r in data:
    print(f"\n== {r[ki][:110]}")
    for key, label in PICK:
        if key in hdr:
            i = hdr.index(key)
            print(f"   {label:44s} {r[i]:>16s} {units[i]}")
    dur_i = hdr.index("gpu__time_duration.sum")
    try:
        dur = float(r[dur_i].replace(",", "")); du = units[dur_i]
        dur_us = dur / 1e3 if du in ("ns", "nsecond") else dur if du in ("us", "usecond") else dur * 1e3
        rd, wr = float(r[hdr.index("dram__bytes_read.sum")].replace(",", "")), float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rdb = rd * mult.get(units[hdr.index("dram__bytes_read.sum")], 1); wrb = wr * mult.get(units[hdr.index("dram__bytes_write.sum")], 1)
        print(f"   {'DRAM bytes / duration':44s} {(rdb + wrb) / dur_us / 1e3:16.1f} GB/s  ({(rdb + wrb) / 1e6:.2f} MB in {dur_us:.2f} us)")
    except Exception as e:
        print("   (could not derive bandwidth:", e, ")")
