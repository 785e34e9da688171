"""Hot spots of an .ncu-rep captured with --import-source on: top SASS instructions by stall samples."""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if rows[j]: body.append(rows[j])
            j += 1
        ci = {h: k for k, h in enumerate(hdr)}
        tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
        print("=== %s: %d SASS instrs, %d samples" % (name, len(body), tot))
        stall_cols = [h for h in hdr if h.startswith("stall_")]
        agg = {h: sum(int(r[ci[h]] or 0) for r in body) for h in stall_cols}
        print("   stall totals:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(tot, 1)) for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
        order = sorted(range(len(body)), key=lambda k: -int(body[k][ci["# Samples"]] or 0))[:top]
        for k in sorted(order):
            r = body[k]
            st = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
            print("   %5d %6.2f%%  %-70s %s" % (k, 100.0 * int(r[ci["# Samples"]] or 0) / max(tot, 1), r[ci["Source"]][:70], st))
        i = j
    else:
        i += 1
