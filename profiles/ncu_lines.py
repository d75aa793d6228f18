#!/usr/bin/env python
"""Joins an `ncu --page source --csv` SASS export with `nvdisasm -g -c` line info: instructions executed,
stall samples and shared-memory wavefronts per CUDA source line of one kernel.
usage: ncu_lines.py <sass_csv> <nvdisasm_output> <kernel substring> [top]"""
import csv, re, sys, collections
sass_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: track current line for each instruction offset inside the wanted function
lines = open(dis, errors='replace').read().splitlines()
cur_fn, cur_line, off2line, in_fn = None, None, {}, False
for ln in lines:
    m = re.search(r'\.section\s+\.text\.(\S+)', ln)
    if m: in_fn = kname in m.group(1); continue
    if not in_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur_line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: off2line[int(m.group(1), 16)] = (cur_line, m.group(2))
r = list(csv.reader(open(sass_csv)))
h = next(i for i, row in enumerate(r) if 'Address' in row)
hdr = r[h]
ai, ii, si = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
wi = hdr.index('L1 Wavefronts Shared') if 'L1 Wavefronts Shared' in hdr else None
rows = []
for row in r[h + 1:]:
    if not row or not row[ai].startswith('0x'): break
    rows.append(row)
base = int(rows[0][ai], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for row in rows:
    off = int(row[ai], 16) - base
    n, s = int(row[ii] or 0), int(row[si] or 0)
    w = int(row[wi] or 0) if wi is not None else 0
    key = off2line.get(off, (None, ''))[0]
    agg[key][0] += n; agg[key][1] += s; agg[key][2] += w
    tot += n
src_cache = {}
def src(key):
    if not key: return ''
    f, l = key
    if f not in src_cache:
        try: src_cache[f] = open('/root/repo/yolo-mslesseg_b200/csrc/' + f).read().splitlines()
        except OSError: src_cache[f] = []
    s = src_cache[f]
    return s[l - 1].strip() if 0 < l <= len(s) else ''
print('total warp instructions', tot)
for key, (n, s, w) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n:>11d} {100 * n / tot:5.1f}%  samples {s:6d}  smem_wf {w:9d}  {key}  {src(key)[:100]}")
