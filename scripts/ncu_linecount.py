"""Development tool: absolute warp / thread instruction counts of source lines matching substrings.
   python scripts/ncu_linecount.py report kernel launch 'substr1|substr2|...'"""
import csv, io, os, subprocess, sys
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as m
rep, kernel, launch, pats = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4].split("|")
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
ci, ti = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
data = blk["rows"][1:]; base = int(data[0][0], 16); lines = m.sass_lines(kernel)
src = open(os.path.join(m.ROOT, "homan_b200", "csrc", "raster.cu")).read().splitlines()
agg = defaultdict(lambda: [0, 0, 0, 0])
for r in data:
    try:
        off = int(r[0], 16) - base; n = float(r[ci] or 0); t = float(r[ti] or 0)
    except ValueError:
        continue
    key = lines.get(off)
    if key and key[0] == "raster.cu":
        a = agg[key[1]]; a[0] += n; a[1] += t; a[2] += 1; a[3] = max(a[3], t)
for ln, a in sorted(agg.items()):
    text = src[ln - 1].strip()
    if any(p in text for p in pats):
        print(f"{ln:5d} sass={a[2]:3.0f} warp={a[0]/1e6:8.2f}M thr={a[1]/1e6:8.2f}M max-thr-per-sass={a[3]/1e6:8.2f}M  {text[:90]}")
