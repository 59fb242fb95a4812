"""Copies the round's measurements from gpurun_out/ (scripts/collect_round_profiles.sh) into profiles/ and derives the
tables from the full ncu capture: per-launch summary (JSON + one line per launch), instruction counts of the raster
forward by source region, hottest lines of both raster kernels, SASS mnemonic histogram.

    python scripts/refresh_profiles.py [round-prefix, default r02]
"""
import csv
import json
import os
import re
import shutil
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rp = sys.argv[1] if len(sys.argv) > 1 else "r02"
COPY = {"bench_cfg3_n1.json": "bench_cfg3_n1.json", "bench_cfg2_n1.json": "bench_cfg2_n1.json",
        "bench_cfg4_n1.json": "bench_cfg4_n1.json", "bench_cfg5_n1.json": "bench_cfg5_n1.json",
        "bench_cfg3_polar.json": "bench_cfg3_polar_hand.json", "bench_reference.json": "bench_reference_cfg3.json",
        "ncu_launches.csv": "ncu_launches.csv", "raster_vs_nmr_style.json": "raster_vs_nmr_style.json",
        "iteration_vs_nmr_style.json": "iteration_vs_nmr_style.json", "pose_init.json": "pose_init.json",
        "bench_cfg3_n2.json": "bench_cfg3_n2.json", "bench_cfg4_n2.json": "bench_cfg4_n2.json",
        "bench_cfg3_n8.json": "bench_cfg3_n8.json", "bench_cfg4_n8.json": "bench_cfg4_n8.json"}
for src, dst in COPY.items():
    if os.path.exists(os.path.join(G, src)) and os.path.getsize(os.path.join(G, src)) > 0:
        shutil.copy(os.path.join(G, src), os.path.join(P, f"{rp}_{dst}"))

rep = os.path.join(G, "prof_final.ncu-rep")
if os.path.exists(rep):
    run = lambda *a: subprocess.run(a, capture_output=True, text=True, cwd=ROOT).stdout  # noqa: E731
    open(os.path.join(P, f"{rp}_ncu_full_summary.txt"), "w").write(run(sys.executable, "scripts/ncu_summary.py", rep))
    rows = list(csv.reader(run("ncu", "-i", rep, "--page", "raw", "--csv").splitlines()))
    h, units = rows[0], rows[1]
    keep = [c for c in h if re.match(r"(Kernel Name|gpu__time_duration|smsp__inst_executed\.sum$|smsp__thread_inst_executed_per_inst|"
                                     r"sm__warps_active|smsp__issue_active|launch__(registers|occupancy_limit|grid_size|block_size|"
                                     r"shared_mem_per_block)|dram__bytes_(read|write)\.sum$|gpu__dram_throughput|sm__throughput|"
                                     r"lts__t_sector_hit_rate|l1tex__t_sector_hit_rate|smsp__average_warps_issue_stalled)", c)]
    launches = [{c: r[h.index(c)] for c in keep} for r in rows[2:]]
    json.dump({"command": "ncu --set full --clock-control none --import-source on -k regex:raster_bwd|raster_fwd|sdf_pair|"
                          "mano_bwd|sil_loss_prep -s 16 -c 9 python bench.py --steps 1 --warmup 3 --no-cpu-baseline "
                          "(cfg3, 480 images per launch)",
               "units": {c: units[h.index(c)] for c in keep}, "launches": launches},
              open(os.path.join(P, f"{rp}_ncu_full_cfg3.json"), "w"), indent=1)
    src = open(os.path.join(ROOT, "homan_b200", "csrc", "raster.cu")).read().splitlines()
    ln = lambda pat: next(i + 1 for i, l in enumerate(src) if pat in l)  # noqa: E731
    marks = [("binning", ln("constexpr int SCAN")), ("row_clipping_(clip_edge)", ln("struct RowCtx")),
             ("untouched_tiles", ln("void write_untouched_tile")), ("kernel_entry+first_scan", ln("raster_fwd_kernel(")),
             ("hi-z_summary+hidden-layer_filter", ln("if (pass == 1) {")), ("face_dealing+prefetch", ln("const int np = cnt[pass];")),
             ("per-face_head", ln("const float4 *rp = wrec;")), ("row_intervals+prefix_sum", ln("// Covered pixels:")),
             ("pixel_loop_(depth+z-buffer)", ln("// barycentric matrix in pixel coordinates and corner 1/z")),
             ("ambiguous_pixels_(exact)", ln("// ---- ambiguous pixels:")), ("write-out", ln("// ---- write-out:")),
             ("(after)", ln("// ------------------------------------------------------------------------------------------ sweep masks"))]
    spec = ",".join(f"{n}:{a}-{b - 1}" for (n, a), (_, b) in zip(marks[:-1], marks[1:]))
    out = ["# raster_fwd_kernel, cfg3 (480 images per launch): warp instructions by source region",
           "# from gpurun_out/prof_final.ncu-rep (ncu --set full --import-source on), joined to the source lines of",
           "# homan_b200/csrc/raster.cu through nvdisasm -g (scripts/ncu_ranges.py, scripts/ncu_lines.py)", ""]
    for title, idx in (("object launch (500 faces)", 0), ("hand launch (1538 faces)", 2)):
        out += [f"## {title}", run(sys.executable, "scripts/ncu_ranges.py", rep, "raster_fwd", str(idx), spec)]
    out += ["## hottest source lines, hand launch", run(sys.executable, "scripts/ncu_lines.py", rep, "raster_fwd", "2", "30")]
    open(os.path.join(P, f"{rp}_raster_fwd_hotspots.txt"), "w").write("\n".join(out))
    out = ["# raster_bwd_kernel, cfg3 (480 images per launch): hottest source lines", "# (same capture and tools)", ""]
    for title, idx in (("hand launch (1538 faces)", 0), ("object launch (500 faces)", 2)):
        out += [f"## {title}", run(sys.executable, "scripts/ncu_lines.py", rep, "raster_bwd", str(idx), "30")]
    open(os.path.join(P, f"{rp}_raster_bwd_lines.txt"), "w").write("\n".join(out))

# SASS mnemonic histogram of the raster / sdf kernels
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "homan_b200", "libhoman_b200.so")], capture_output=True,
                      text=True).stdout
out = ["# SASS of homan_b200/libhoman_b200.so (cuobjdump -sass, sm_100a): instruction mix of the heavy kernels",
       "# (mnemonic histogram; the full listing is reproducible with `cuobjdump -sass homan_b200/libhoman_b200.so`)", ""]
cur, hist = None, Counter()


def flush():
    if cur and any(k in cur for k in ("raster_fwd", "raster_bwd", "sil_loss_prep", "sdf_pair")):
        out.append(f"## {cur}  ({sum(hist.values())} instructions)")
        out.extend(f"   {m:28s} {n}" for m, n in hist.most_common(24))
        tma = sum(n for m, n in hist.items() if m.startswith(("UBLKCP", "UTMA", "SYNCS")))
        tc = sum(n for m, n in hist.items() if m.startswith(("UTC", "HMMA", "LDTM", "STTM")))
        out.append(f"   (TMA / mbarrier mnemonics: {tma}; tensor-core / TMEM mnemonics: {tc})")
        out.append("")


for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        cur, hist = m.group(1), Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[m.group(1)] += 1
flush()
open(os.path.join(P, f"{rp}_sass_excerpt.txt"), "w").write("\n".join(out))
print("refreshed", sorted(f for f in os.listdir(P) if f.startswith(rp)))
