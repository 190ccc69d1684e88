"""Builds profiles/r2_ncu_traffic.json from two ncu captures of THE CURRENT library build (stamped with its sha256[:16]; bench.py
reports `roofline.traffic` / `step_roofline` only when the stamp equals the library it times):
  gpurun_out/r2_step_launches.csv   ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,
                                    dram__bytes_write.sum --clock-control none --csv of benchmarks/profile_step.py (one step)
  gpurun_out/prof_onepassL1.ncu-rep ncu --set full --clock-control none of the one-pass kernel at layer 1, batch 64
    python scripts/ncu_traffic.py   (after scripts/r2_capture.sh ran on the GPU box)
"""
import collections
import csv
import hashlib
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    import sys
    sys.path.insert(0, ROOT)
    import bench
    sha = bench.lib_sha16()            # sha256 over the library's sources (the binary is not reproducible bit for bit)
    stamp_file = os.path.join(ROOT, "gpurun_out", "r2_capture_lib_sha16.txt")
    captured = open(stamp_file).read().strip() if os.path.exists(stamp_file) else None
    out = {"lib_sha16": captured or sha, "lib_sha16_now": sha}
    # ---- one-step launch list -----------------------------------------------------------------------------------------------
    rows = [l for l in open(os.path.join(ROOT, "gpurun_out", "r2_step_launches.csv")) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    per = collections.OrderedDict()
    for r in rd:
        d = per.setdefault(r["ID"], {"kernel": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("ldc::", "")})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    launches = list(per.values())
    step_bytes = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in launches)
    step_ns = sum(d.get("gpu__time_duration.sum", 0) for d in launches)
    by_kernel = collections.OrderedDict()
    for d in launches:
        k = re.sub(r"<.*", "", d["kernel"])
        a = by_kernel.setdefault(k, {"launches": 0, "us": 0.0, "dram_MB": 0.0})
        a["launches"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0) / 1e3
        a["dram_MB"] += (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / 1e6
    for a in by_kernel.values():
        a["us"], a["dram_MB"] = round(a["us"], 1), round(a["dram_MB"], 1)
        a["share_of_step_time"] = round(a["us"] * 1e3 / max(step_ns, 1), 4)
    out.update({"step_dram_bytes": int(step_bytes), "step_launches": len(launches), "step_sum_of_kernel_us_under_ncu": round(step_ns / 1e3, 1),
                "step_by_kernel": by_kernel,
                "step_source": "ncu launch list of one eager step (serialised, cold cache): profiles/r2_step_launches.csv"})
    # ---- the roofline kernel: one --set full capture ---------------------------------------------------------------------------
    rep = os.path.join(ROOT, "gpurun_out", "prof_onepassL1.ncu-rep")
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(io.StringIO(raw)))
        hdr, vals = rr[0], rr[2]
        g = lambda k: float(vals[hdr.index(k)].replace(",", ""))
        unit = lambda k: rr[1][hdr.index(k)]
        def nbytes(k):
            v, u = g(k), unit(k).lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        rd_b, wr_b = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
        out.update({"roofline_kernel_bytes": int(rd_b + wr_b), "roofline_kernel_read_bytes": int(rd_b), "roofline_kernel_write_bytes": int(wr_b),
                    "roofline_kernel_us_under_ncu": round(g("gpu__time_duration.sum") / (1e3 if unit("gpu__time_duration.sum") in ("ns", "nsecond") else 1), 1),
                    "roofline_kernel": vals[hdr.index("Kernel Name")][:60],
                    "roofline_source": "ncu --set full --clock-control none, ldconv_onepass_kernel at layer 1 (B=64, 16->32, N=3, s=2, 320x320); writes still "
                                       "resident in the 126 MB L2 at kernel end are not counted by the DRAM counters; summary profiles/r2_ncu_onepassL1.txt"})
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "step_by_kernel"}, indent=1))


if __name__ == "__main__":
    main()
