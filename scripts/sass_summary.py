"""SASS evidence without a GPU: counts of the Blackwell-native mnemonics per object file of the library (cuobjdump -sass), written to
profiles/sass_r2.txt together with one excerpt of the CTA-pair MMA loop."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
build = os.path.join(ROOT, "rust-autograd_b200", "csrc", "build")
keys = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG.2CTA", "UTMALDG", "UTMASTG", "UTCBAR.2CTA.MULTICAST", "UTCBAR", "LDTM", "UTCATOMSWS", "HMMA"]
out = ["# cuobjdump -sass of rust-autograd_b200/csrc/build/*.o (sm_100a): instruction counts per file\n"]
excerpt = None
for f in sorted(os.listdir(build)):
    if not f.startswith("tc_") or not f.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(build, f)], capture_output=True, text=True).stdout
    cnt = collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"\b(UTCHMMA(?:\.2CTA)?|UTMALDG(?:\.\dD)?(?:\.2CTA)?|UTMASTG\S*|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|LDTM\S*|UTCATOMSWS\S*|HMMA\S*)", line)
        if m:
            cnt[m.group(1)] += 1
        if excerpt is None and "UTCHMMA.2CTA" in line:
            excerpt = (f, line.strip())
    out.append("%-28s %s" % (f, "  ".join("%s=%d" % kv for kv in sorted(cnt.items()))))
if excerpt:
    out.append("\nfirst UTCHMMA.2CTA in %s:\n  %s" % excerpt)
open(os.path.join(ROOT, "profiles", "sass_r2.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
