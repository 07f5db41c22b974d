"""Hot SASS lines of an `ncu --set full --import-source on` report (run where ncu is installed, no GPU needed):
python scripts/ncu_hot.py gpurun_out/x.ncu-rep [top N] -> sample count, dominant stall reasons, SASS, wavefront excess."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k = 0
while k < len(rows):
    if rows[k] and rows[k][0] == "Kernel Name":
        name = rows[k][1]
        hdr = rows[k + 1]
        body = []
        k += 2
        while k < len(rows) and not (rows[k] and rows[k][0] == "Kernel Name"):
            if len(rows[k]) == len(hdr):
                body.append(rows[k])
            k += 1
        ci = {h: i for i, h in enumerate(hdr)}
        sa = ci["Warp Stall Sampling (All Samples)"]
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(float(r[sa] or 0) for r in body)
        print("== %s: %d samples, %d SASS lines" % (name, tot, len(body)))
        agg = {s: sum(float(r[ci[s]] or 0) for r in body) for s in stalls}
        print("   stall totals:", ", ".join("%s %.0f" % (s[6:], v) for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        order = sorted(range(len(body)), key=lambda i: -float(body[i][sa] or 0))[:top_n]
        for i in sorted(order):
            r = body[i]
            st = sorted(((float(r[ci[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
            exc = r[ci["L1 Wavefronts Shared Excessive"]]
            print("%5d %6s  %-70s %s  exc=%s" % (i, r[sa], r[ci["Source"]].strip()[:70], " ".join("%s:%.0f" % (n, v) for v, n in st if v > 0), exc))
    else:
        k += 1
