"""Development tool: instruction counts of one kernel launch in an ncu report, grouped by named source regions
(markers `// @region name` in raster.cu start a region)."""
import csv, io, os, re, subprocess, sys
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as m

rep, kernel, launch = sys.argv[1], sys.argv[2], int(sys.argv[3])
src = open(os.path.join(m.ROOT, "homan_b200", "csrc", "raster.cu")).read().splitlines()
marks = [(i + 1, re.search(r"@region (\S+)", l).group(1)) for i, l in enumerate(src) if "@region" in l]
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
    name = "(headers)"
    if key[0] == "raster.cu":
        name = "(before)"
        for ln, nm in marks:
            if ln <= key[1]:
                name = nm
    agg[name][0] += n; agg[name][1] += t; agg[name][2] += s_; tot += n; tots += s_
print(blk["name"][:50], "launch", launch, f"{tot / 1e6:.1f} M warp instructions")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:24s} {a[0] / tot * 100:5.1f}% inst {a[0] / 1e6:7.1f}M  {a[1] / max(a[0], 1):5.1f} thr/inst  {a[2] / max(tots, 1) * 100:5.1f}% samples")
