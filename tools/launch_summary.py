"""python tools/launch_summary.py LAUNCHES.csv: per-kernel count / total time / share of an
`ncu --metrics gpu__time_duration.sum --csv --log-file` launch list."""
import csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = {}
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik])
    ns = float(r[iv].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iu], 1.0)
    c, t = tot.get(name, (0, 0.0))
    tot[name] = (c + 1, t + ns)
total = sum(t for _, t in tot.values())
print("%-72s %6s %12s %7s" % ("kernel", "count", "total ms", "share"))
for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %6d %12.3f %6.2f%%" % (name[:72], c, t / 1e6, 100.0 * t / total))
