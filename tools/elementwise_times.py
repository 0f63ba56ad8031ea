"""Summarise an ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum CSV of
tools/elementwise_probe.py: median time and DRAM GB/s per davo kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hd = rows[h]; ki = hd.index("Kernel Name"); mi = hd.index("Metric Name"); vi = hd.index("Metric Value"); ui = hd.index("Metric Unit")
d = collections.defaultdict(lambda: collections.defaultdict(list))
scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows[h + 2:]:
    if len(r) > vi and ("davo" in r[ki] or "_kernel<" in r[ki]) and "at::" not in r[ki]:
        d[r[ki][:64]][r[mi]].append(float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0))
med = lambda L: sorted(L)[len(L) // 2]
for k, m in d.items():
    t, rd, wr = med(m["gpu__time_duration.sum"]), med(m["dram__bytes_read.sum"]), med(m["dram__bytes_write.sum"])
    print(f"{k:66s} {t*1e6:8.1f} us  {(rd+wr)/1e6:8.1f} MB  {(rd+wr)/t/1e9:7.0f} GB/s")
