"""Summarise .ncu-rep files (raw page) into a small table: python scripts/ncu_summary.py rep1 [rep2 ...]"""
import csv, subprocess, sys
WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("## " + rep)
    print("| kernel | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|" + "---|" * len(WANT))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0][:48]
        cells = []
        for key, _ in WANT:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                cells.append(v + (" " + units[i] if units[i] not in ("", "%") and _ not in ("regs", "grid", "block") else ""))
            else:
                cells.append("-")
        print("| " + name + " | " + " | ".join(cells) + " |")
    print()
