"""Development tool: per-source-line instruction counts of one kernel from an ncu report.

    python scripts/ncu_lines.py report.ncu-rep kernel_substring [launch_index] [top]

ncu's CSV source page lists SASS instructions with their counters but without source lines; `nvdisasm -g` on the
cubin of homan_b200/libhoman_b200.so gives the line of every instruction. The two are joined by instruction offset."""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "homan_b200", "libhoman_b200.so")], cwd=tmp,
                          stdout=subprocess.DEVNULL)
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        try:
            txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        except OSError:
            continue
        if kernel not in txt:
            continue
        out, cur, inside = {}, None, False
        for line in txt.splitlines():
            if line.startswith("\t.section\t.text.") or line.startswith(".text."):
                inside = kernel in line
            if not inside:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*)", line)
            if m:
                out[int(m.group(1), 16)] = cur
        if out:
            return out
    return {}


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    blocks = [b for b in blocks if kernel in b["name"]]
    blk = blocks[launch]
    hdr = blk["rows"][0]
    ci, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    data = blk["rows"][1:]
    base = int(data[0][0], 16)
    lines = sass_lines(kernel)
    src = open(os.path.join(ROOT, "homan_b200", "csrc", "raster.cu")).read().splitlines()
    agg = defaultdict(lambda: [0.0, 0.0, 0.0])
    tot = tot_s = 0.0
    for r in data:
        try:
            off = int(r[0], 16) - base
            n, t, s = float(r[ci] or 0), float(r[ti] or 0), float(r[si] or 0)
        except ValueError:
            continue
        key = lines.get(off) or ("?", 0)
        a = agg[key]
        a[0] += n; a[1] += t; a[2] += s
        tot += n; tot_s += s
    print(f"{blk['name'][:60]}  launch {launch}: {tot / 1e6:.1f} M warp instructions, {tot_s:.0f} samples")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = src[key[1] - 1].strip()[:100] if key[0] == "raster.cu" and 0 < key[1] <= len(src) else ""
        print(f"{key[0]}:{key[1]:<5} {a[0] / tot * 100:5.1f}% inst  {a[1] / max(a[0], 1):5.1f} thr/inst  {a[2] / max(tot_s, 1) * 100:5.1f}% samples  {text}")


if __name__ == "__main__":
    main()
