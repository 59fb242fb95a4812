"""Development tool: instruction counts of one kernel launch grouped by source-line ranges of raster.cu.
   python scripts/ncu_ranges.py report kernel launch  "name:lo-hi,name:lo-hi,..." """
import csv, io, os, subprocess, sys
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as m
rep, kernel, launch, spec = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
ranges = []
for part in spec.split(","):
    n, r = part.split(":"); lo, hi = r.split("-"); ranges.append((n, int(lo), int(hi)))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
blk = [b for b in blocks if kernel in b["name"]][launch]
hdr = blk["rows"][0]
ci, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
data = blk["rows"][1:]; base = int(data[0][0], 16); lines = m.sass_lines(kernel)
agg = defaultdict(lambda: [0, 0, 0]); tot = tots = 0
for r in data:
    try:
        off = int(r[0], 16) - base; n = float(r[ci] or 0); t = float(r[ti] or 0); s_ = float(r[si] or 0)
    except ValueError:
        continue
    key = lines.get(off) or ("?", 0)
    name = "(headers:%s)" % key[0]
    if key[0] == "raster.cu":
        name = "(other)"
        for nm, lo, hi in ranges:
            if lo <= key[1] <= hi:
                name = nm; break
    agg[name][0] += n; agg[name][1] += t; agg[name][2] += s_; tot += n; tots += s_
print(blk["name"][:50], "launch", launch, f"{tot / 1e6:.1f} M warp instructions")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:34s} {a[0] / tot * 100:5.1f}% inst {a[0] / 1e6:7.1f}M  {a[1] / max(a[0], 1):5.1f} thr/inst  {a[2] / max(tots, 1) * 100:5.1f}% samples")
