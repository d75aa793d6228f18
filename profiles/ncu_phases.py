#!/usr/bin/env python
"""Aggregates the per-line output of ncu_lines.py into the phases of enhance_dense_kernel (by source markers)."""
import re, sys
rows = []
for ln in open(sys.argv[1]):
    m = re.match(r"\s*(\d+)\s+([\d.]+)%\s+samples\s+(\d+)\s+smem_wf\s+(\d+)\s+\('([^']+)', (\d+)\)", ln)
    if m: rows.append((int(m.group(1)), int(m.group(3)), int(m.group(4)), m.group(5), int(m.group(6))))
src = open('/root/repo/yolo-mslesseg_b200/csrc/msl_enhance_dense.cu').read().splitlines()
def find(txt): return next(i + 1 for i, l in enumerate(src) if txt in l)
marks = [('load + blank check', find('// ---------------------------------------------------------------- load')),
         ('tile histograms', find('tile histograms over u (real pixels)')),
         ('HE sum + padding', find("HE's histogram = sum of the 64 tile histograms")),
         ('HE-only histogram', find('HE without CLAHE')),
         ('HE CDF + HE|GC|LT map', find('HE CDF -> LUT; packed HE | GC | LT table')),
         ('fold / clip / CDF', find('CLAHE: fold u-bins into L-bins')),
         ('weight tables', find('interpolation tables (OpenCV CLAHE_Interpolation_Body)')),
         ('pair tables', find('Pair tables: PT[ty][u][j]')),
         ('blend', find('CLAHE: bilinear blend + LUT_OUT')),
         ('end', 10 ** 6)]
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
