#!/usr/bin/env python
"""Aggregates the per-line output of ncu_lines.py into the phases of enhance_dense_kernel (by source markers)."""
import re, sys
rows = []
for ln in open(sys.argv[1]):
    m = re.match(r"\s*(\d+)\s+([\d.]+)%\s+samples\s+(\d+)\s+smem_wf\s+(\d+)\s+\('([^']+)', (\d+)\)", ln)
    if m: rows.append((int(m.group(1)), int(m.group(3)), int(m.group(4)), m.group(5), int(m.group(6))))
src = open('/root/repo/yolo-mslesseg_b200/csrc/msl_enhance_dense.cu').read().splitlines()
def find(txt):
    return next((i + 1 for i, l in enumerate(src) if txt in l), None)
cands = [('clip / CDF (clip_cdf_tile)', 'OpenCV CLAHE_CalcLut_Body for one tile whose 256 L-bins'),
         ('plane tables kernel', 'Everything that depends only on the plane (geometry + LUT_L)'),
         ('prologue', 'template <bool DO_CLAHE>'),
         ('load + blank check', '// ---------------------------------------------------------------- load'),
         ('tile loop: queue + clear', 'tile LUTs, one warp per tile, no block barrier inside'),
         ('tile loop: real pixels', '// real pixels'),
         ('tile loop: HE add', "HE's histogram = sum of the tile histograms of the REAL pixels"),
         ('tile loop: padding', 'BORDER_REFLECT_101 padding (OpenCV pads bottom / right'),
         ('tile loop: fold', 'fold the u-bins into L-bins, 8 L-bins per lane'),
         ('HE-only histogram', 'HE without CLAHE'),
         ('HE CDF + HE|GC|LT map', 'HE CDF -> LUT; packed HE | GC | LT table'),
         ('pair tables', 'Pair tables: PT[ty][u][j]'),
         ('blend', 'CLAHE: bilinear blend + LUT_OUT')]
marks = sorted([(n, find(t)) for n, t in cands if find(t) is not None], key=lambda m: m[1]) + [('end', 10 ** 6)]
agg = {}
tot_i = sum(r[0] for r in rows); tot_s = sum(r[1] for r in rows)
for n, s_, w, f, l in rows:
    ph = f
    if f == 'msl_enhance_dense.cu':
        ph = 'prologue / helpers'
        for (name, start), (_, nxt) in zip(marks, marks[1:]):
            if start <= l < nxt: ph = name
    a = agg.setdefault(ph, [0, 0, 0]); a[0] += n; a[1] += s_; a[2] += w
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} instructions {100 * v[0] / tot_i:5.1f}%   stall samples (time) {100 * v[1] / tot_s:5.1f}%   smem wavefronts {v[2] / 1e6:6.2f}M")
print("total warp instructions", tot_i, " samples", tot_s)
