"""One row per kernel launch of an `ncu --set full` capture of a whole step: duration, DRAM bytes and the busiest units.
    python scripts/ncu_step_table.py gpurun_out/prof_step_full.ncu-rep > profiles/r2_ncu_step_full.txt"""
import csv
import io
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "rd_MB", None), ("dram__bytes_write.sum", "wr_MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1.0),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 1.0),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1.0),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tc%", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1.0)]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("%-58s %8s %8s %8s %6s %6s %6s %6s %6s %6s" % ("kernel", *[c[1] for c in COLS]))
    tot_us = tot_mb = 0.0
    for vals in rows[2:]:
        name = re.sub(r"\(.*", "", vals[ix["Kernel Name"]]).replace("void ", "").replace("ldc::", "")
        cells = []
        for metric, label, scale in COLS:
            try:
                v = float(vals[ix[metric]].replace(",", ""))
            except (KeyError, ValueError):
                v = float("nan")
            if scale is None:       # bytes in whatever unit ncu chose
                u = units[ix[metric]].lower()
                v *= {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1.0)
            elif label == "us":
                u = units[ix[metric]].lower()
                v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
            cells.append(v)
        tot_us += cells[0]
        tot_mb += cells[1] + cells[2]
        print("%-58s %8.1f %8.1f %8.1f %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f" % (name[:58], *cells))
    print("%-58s %8.1f %8.1f MB of DRAM traffic in %d launches (serialised, cold-cache replays: shares, not absolute times)" %
          ("total", tot_us, tot_mb, len(rows) - 2))


if __name__ == "__main__":
    main()
