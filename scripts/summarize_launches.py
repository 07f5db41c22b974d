"""Markdown summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total us and share per kernel.
usage: python scripts/summarize_launches.py profiles/launches_lstm_r1.csv [--second-half]   (--second-half: the file holds two steps, report the second)"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
    hdr, recs = None, []
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr, ki, vi = r, r.index("Kernel Name"), r.index("Metric Value")
            continue
        if len(r) <= vi:
            continue
        try:
            recs.append((r[ki], float(r[vi].replace(",", ""))))
        except ValueError:
            pass
    if "--second-half" in sys.argv:
        recs = recs[len(recs) // 2:]
    cnt, tm = collections.Counter(), collections.Counter()
    for name, v in recs:
        nm = re.sub(r"\(.*", "", re.sub(r"^void ", "", name))[:70]
        cnt[nm] += 1
        tm[nm] += v
    tot = sum(tm.values())
    print("%d launches, %.2f ms (cold-cache, serialised)\n" % (len(recs), tot / 1e6))
    print("| kernel | launches | us | share |\n|---|---|---|---|")
    for k, v in tm.most_common():
        print("| `%s` | %d | %.0f | %.1f %% |" % (k, cnt[k], v / 1e3, 100 * v / tot))


if __name__ == "__main__":
    main()
