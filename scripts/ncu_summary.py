"""Text summary of an `ncu --set full` capture (the files kept under profiles/): selected raw metrics per kernel and, with
--lines N, the N SASS lines with the most warp-stall samples.
    python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep [--lines 25] > profiles/r1_ncu_X.txt
"""
import argparse
import csv
import io
import subprocess

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--lines", type=int, default=0)
    args = ap.parse_args()
    rows = page(args.rep, "raw")
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("--- " + name[:110])
        for i, h in enumerate(hdr):
            stall = h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
            if h in WANT or stall:
                print("  %-92s %-16s %s" % (h, units[i], vals[i]))
    if args.lines:
        rows = page(args.rep, "source")
        hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
        hdr, data = rows[hi], rows[hi + 1:]
        ix = {h: i for i, h in enumerate(hdr)}

        def f(r, k):
            try:
                return float(r[ix[k]])
            except (ValueError, IndexError, KeyError):
                return 0.0
        tot = sum(f(r, "# Samples") for r in data) or 1.0
        print("\nSASS lines with the most warp-stall samples (of %d):" % tot)
        keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
        for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:args.lines]:
            why = {k[6:]: int(f(r, k)) for k in keys if f(r, k) > 0.05 * f(r, "# Samples")}
            print("  %5.1f%%  %-72s %s" % (100 * f(r, "# Samples") / tot, r[1][:72], why))


if __name__ == "__main__":
    main()
