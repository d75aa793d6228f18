#!/usr/bin/env python
"""Measurement of the two widened rows (SURVEY 8f-3 first half, 8f-4) on one B200, CPU restatement timed beside.
Prints one JSON object; CUDA events after warm-up, inputs larger than L2 or L2 flushed between iterations.
usage (GPU box): python profiles/bench_next_rows.py > gpurun_out/next_rows.json"""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "yolo-mslesseg_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from mslesseg_b200 import ops, synthetic as S
from oracle import oracle as O
from oracle.make_golden_pred import instance_masks

dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

out = {"peak_gbs": peak}
# ---- R0: 120 slices (one patient, three planes x 40) x 3 instances at 640 x 544 -> (218, 182) images
mh, mw, h, w, nsl, ninst = 640, 544, 218, 182, 120, 3
base = torch.from_numpy(instance_masks(3, ninst, mh, mw)).to(dev)
masks = base.repeat(nsl, 1, 1).contiguous()
off = torch.arange(0, nsl * ninst + 1, ninst, dtype=torch.int32, device=dev)
q = torch.empty((nsl, w, h), dtype=torch.uint8, device=dev)
ms = timed(lambda: ops.combine_predictions(masks, off, rows=w, cols=h, layout="G", out=q))
alg = nsl * (ninst * h * w * 4 + h * w)
t0 = time.perf_counter()
hm = masks[:ninst].cpu().numpy()
for _ in range(8): O.normalizar_prediccion(O.combinar_predicciones(list(hm), (h, w)))
cpu_ms = (time.perf_counter() - t0) / 8 * 1e3 * nsl
out["combine_predictions"] = {"slices": nsl, "instances_per_slice": ninst, "mask": [mh, mw], "image": [h, w], "ms": ms,
                              "algorithmic_bytes": alg, "gb_s": alg / ms / 1e6, "frac_of_peak": alg / ms / 1e6 / peak,
                              "mask_bytes_resident": int(masks.numel() * 4),
                              "note": "algorithmic = sampled mask values (4 B each) + 1 B written per pixel; nearest sampling touches ~1/3 of the mask rows",
                              "cpu_oracle_ms_same_work": cpu_ms, "cpu_cores": 1}
# ---- 8f-4: per-slice counts of 32 volumes
B = 32
pats = [S.make_patient(i + 1, 4, 40) for i in range(4)]
gt = torch.from_numpy(np.stack([p.gt for p in pats])).to(dev).repeat(B // 4, 1, 1, 1).contiguous()
pred = (gt.roll(1, 3) | gt.roll(2, 2)).contiguous()
ms = timed(lambda: ops.slice_counts(gt, pred))
alg = 2 * gt.numel()
g0, p0 = pats[0].gt.transpose(2, 1, 0), pred[0].cpu().numpy().transpose(2, 1, 0)
t0 = time.perf_counter()
for plano in ("axial", "coronal", "sagital"):
    for i in range(g0.shape[O.plane_axis(plano)]): O.confusion_counts(O.slice_of(g0, plano, i), O.slice_of(p0, plano, i))
cpu_ms = (time.perf_counter() - t0) * 1e3 * B
out["slice_counts"] = {"volumes": B, "ms": ms, "algorithmic_bytes": alg, "gb_s": alg / ms / 1e6, "frac_of_peak": alg / ms / 1e6 / peak,
                       "gvoxel_s": gt.numel() / ms / 1e6, "cpu_oracle_ms_same_work": cpu_ms, "cpu_cores": 1}
# ---- 8f-1 (encode side): PNG files for 1184 RGBA slices of 218 x 182 (eight per SM)
NPNG = 1184
rgba = torch.randint(0, 256, (NPNG, 218, 182, 4), dtype=torch.uint8, device=dev)
ms = timed(lambda: ops.png_pack(rgba))
files, size = ops.png_pack(rgba)
alg = rgba.numel() + NPNG * size
import io
try:
    from PIL import Image
    h0 = rgba[0].cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(5):
        Image.fromarray(h0, mode="RGBA").save(io.BytesIO(), format="PNG")
    cpu_ms = (time.perf_counter() - t0) / 5 * 1e3 * NPNG
except ImportError:
    cpu_ms = None
out["png_pack"] = {"images": NPNG, "shape": [218, 182, 4], "file_bytes": size, "ms": ms, "algorithmic_bytes": alg,
                   "gb_s": alg / ms / 1e6, "frac_of_peak": alg / ms / 1e6 / peak,
                   "cpu_pillow_deflate_ms_same_images": cpu_ms, "cpu_cores": 1,
                   "note": "pixels read + file written; the CPU figure is Pillow's PNG encoder (deflate) on random pixels"}
print(json.dumps(out))
