"""Per-kernel mean duration of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  python tools/launch_summary.py file.csv [...]"""
import collections, csv, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(l for l in open(f) if not l.startswith("=="))]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    d = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi: continue
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        d.setdefault(r[ki].split("(")[0].replace("void ", ""), []).append(v)
    print(f)
    for k, v in d.items(): print(f"   {k:28s} n={len(v):3d}  mean {sum(v) / len(v):9.1f} us")
