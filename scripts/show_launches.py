"""print an ncu --csv launch list (gpu__time_duration.sum, dram bytes) as one line per launch: python scripts/show_launches.py FILE [min_us]"""
import collections, csv, io, re, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
per = collections.OrderedDict()
for r in csv.DictReader(io.StringIO("".join(rows))):
    d = per.setdefault(r["ID"], {"k": re.sub(r"\(.*", "", r["Kernel Name"])[:70]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
lo = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
tot = 0.0
for d in per.values():
    t = d.get("gpu__time_duration.sum", 0) / 1e3
    tot += t
    if t >= lo:
        print(f"{t:8.1f} us  {(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)) / 1e6:8.1f} MB  {d['k']}")
print(f"{tot:8.1f} us total, {len(per)} launches")
