"""One row per kernel launch of the one-step ncu launch list (scripts/r2_capture.sh: time, DRAM bytes and the busiest units of every
launch, `--clock-control none`, serialised cold-cache replays: shares, not absolute times).
    python scripts/ncu_step_table.py gpurun_out/r2_step_launches.csv > profiles/r2_ncu_step_table.txt"""
import collections
import csv
import io
import re
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd_MB"), ("dram__bytes_write.sum", "wr_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tc%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%")]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3, "%": 1.0}


def main():
    rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
    per = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(rows))):
        d = per.setdefault(r["ID"], {"kernel": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("ldc::", "")})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * SCALE.get(r["Metric Unit"].lower(), 1.0)
        except ValueError:
            pass
    print("%-58s %8s %8s %8s %6s %6s %6s %6s %6s %6s" % ("kernel", *[c[1] for c in COLS]))
    tot_us = tot_mb = 0.0
    for d in per.values():
        cells = [d.get(m, float("nan")) for m, _ in COLS]
        tot_us += cells[0]
        tot_mb += cells[1] + cells[2]
        print("%-58s %8.1f %8.1f %8.1f %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f" % (d["kernel"][:58], *cells))
    print("%-58s %8.1f us, %.1f MB of DRAM traffic in %d launches" % ("total", tot_us, tot_mb, len(per)))


if __name__ == "__main__":
    main()
