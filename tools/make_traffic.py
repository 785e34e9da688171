"""Write profiles/traffic.json from an ncu capture of the 2-D transform kernels of a workload:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:"k_fft_pass|k_fft_colsub2|k_split_p|k_col_radix|k_fft_colsub" -s <warm-up> -c <n> --csv --log-file X.csv <cmd>
    python tools/make_traffic.py X.csv coupled8192 "<where the capture is kept>"

dram_bytes_per_pass = (DRAM bytes of all captured transform launches) / (2 * number of 2-D transforms captured), i.e. the
measured counterpart of the 32 B/point a pass moves algorithmically.  The record carries the hash of csrc/ it was taken
on; bench.py reports `traffic: null` when the sources have changed since."""
import csv, json, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import csrc_hash, WORKLOADS

path, workload, source = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[iid], r[ik].split("(")[0].replace("void ", "")), {})[r[im]] = float(r[iv].replace(",", ""))
kinds = collections.OrderedDict()
for (_, name), d in per.items():
    k = kinds.setdefault(name, dict(n=0, bytes=0.0, ns=0.0))
    k["n"] += 1
    k["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    k["ns"] += d.get("gpu__time_duration.sum", 0.0)
rows_n = sum(k["n"] for n, k in kinds.items() if n.startswith("k_fft_pass"))
cols_n = sum(k["n"] for n, k in kinds.items() if n.startswith("k_fft_colsub"))
n2d = max(rows_n, cols_n)
# row passes with a loader (last template argument 1: k_fft_pass<M, W, 1, 0, 1, 1>) read three fields instead of one: they
# count as one pass more (64 B per point instead of 32)
def _ld(n):
    parts = n.rstrip(">").split(",")
    return int(parts[-1]) if n.startswith("k_fft_pass") and len(parts) == 6 else 0
n_ld = sum(k["n"] for n, k in kinds.items() if _ld(n) == 1)       # three operands: +32 B per point
n_ld2 = sum(k["n"] for n, k in kinds.items() if _ld(n) == 2)      # two operands: +16 B per point
total = sum(k["bytes"] for k in kinds.values())
model, nx, batch = WORKLOADS[workload]
alg = 32.0 * nx * nx * batch
npass = 2 * n2d + n_ld + 0.5 * n_ld2
rec = {"dram_bytes_per_pass": total / npass, "algorithmic_bytes_per_pass": alg, "ratio": total / npass / alg,
       "transforms_captured": n2d, "loader_row_passes_captured": n_ld + n_ld2, "csrc_hash": csrc_hash(), "source": source,
       "kernels": {n: {"launches": k["n"], "dram_bytes_per_launch": k["bytes"] / k["n"], "avg_us": k["ns"] / k["n"] / 1e3}
                   for n, k in kinds.items()}}
out = os.path.join(ROOT, "profiles", "traffic.json")
try:
    allrec = json.load(open(out))
    if not isinstance(allrec.get(workload, {}), dict):
        allrec = {}
except Exception:
    allrec = {}
allrec = {k: v for k, v in allrec.items() if isinstance(v, dict)}
allrec[workload] = rec
json.dump(allrec, open(out, "w"), indent=1)
print(json.dumps(rec, indent=1))
