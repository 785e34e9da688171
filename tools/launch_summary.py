"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share, mean."""
import csv, sys, re, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iu = hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(u, 1e-6)
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "")
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + v)
T = sum(t for _, t in tot.values())
print("# %s" % " ".join(sys.argv[2:]))
print("# per-launch times are cold-cache/serialised under ncu: compare SHARES with bench.py's kernel_breakdown, not absolutes")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-60s n=%4d total %9.1f ms (%5.1f%%) avg %.3f ms" % (name[:60], n, t, 100 * t / T, t / n))
