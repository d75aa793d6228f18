#!/usr/bin/env python
"""Stand-alone timings of the device codec kernels on the bench's data (one stream, CUDA events, L2 flushed by a scratch write).
usage: python profiles/bench_codec.py [--vols 4]"""
import argparse, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "yolo-mslesseg_b200")]
import numpy as np, torch
from mslesseg_b200 import _lib, ops, codec, synthetic as S

ap = argparse.ArgumentParser(); ap.add_argument("--vols", type=int, default=4); ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
_lib.load()
dev = torch.device("cuda:0")
pat = [S.make_patient(1 + b, config_id=4, num_cortes=40) for b in range(min(4, args.vols))]
flair = torch.stack([torch.from_numpy(pat[b % len(pat)].flair) for b in range(args.vols)]).to(dev)
gt = torch.stack([torch.from_numpy(pat[b % len(pat)].gt) for b in range(args.vols)]).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, reps=args.reps):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps

res = {}
outs = ops.enhance_volumes(flair, ("HE", "CLAHE"), ("axial",))
for m in ("HE", "CLAHE"):
    px = outs[(m, "axial")].reshape(-1, 218, 182)
    wsb = int(_lib.load().msl_deflate_workspace_bytes(px.shape[0], _lib.Z_PNG, 218 * 183))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev); out = torch.empty(int(_lib.load().msl_deflate_bound(px.shape[0], _lib.Z_PNG, 218 * 183)), dtype=torch.uint8, device=dev)
    _lib.profile_enable(True)
    t = timeit(lambda: ops.png_encode(px, out=out, workspace=ws))
    prof = _lib.profile_collect()
    ps = ops.png_encode(px, out=out, workspace=ws)
    size = int(ps.off[-1].item())
    res[f"png_encode_{m}_axial"] = {"images": int(px.shape[0]), "raw_mb": px.numel() / 1e6, "file_mb": size / 1e6, "ms": t, "raw_gb_s": px.numel() / t / 1e6,
                                    "kernels_ms": {k: v[0] / (args.reps + 1) for k, v in prof.items()}}
# nifti gz: float32 volumes, uint8 masks as float32
aff = np.eye(4)
for name, vol, kw in (("flair_f32", flair, {}), ("mask_as_f32", gt, {"como_float32": True}), ("mask_u8", gt, {})):
    t = timeit(lambda: codec.nifti_gz_device(vol, aff, **kw))
    ps = codec.nifti_gz_device(vol, aff, **kw)
    raw = vol.numel() * (4 if (vol.dtype == torch.float32 or kw) else 1)
    res[f"nifti_gz_{name}"] = {"raw_mb": raw / 1e6, "file_mb": int(ps.off[-1].item()) / 1e6, "ms": t, "raw_gb_s": raw / t / 1e6}
# inflate of those files (members in parallel)
for name, vol, kw, dt in (("flair_f32", flair, {}, torch.float32), ("mask_as_f32", gt, {"como_float32": True}, torch.uint8)):
    ps = codec.nifti_gz_device(vol, aff, **kw)
    data, off = ps.to_host()
    meta = ps.meta.cpu().numpy().astype(np.int64) & 0xffffffff
    src = ps.data[:int(off[-1]) + 8].clone()
    so = torch.from_numpy(off.copy()).to(dev)
    do_ = np.zeros(len(off), np.int64); np.cumsum((meta[:, 1] + 15) & ~15, out=do_[1:])
    dst = torch.empty(int(do_[-1]) + 64, dtype=torch.uint8, device=dev); do = torch.from_numpy(do_).to(dev)
    st = torch.empty((len(off) - 1, 4), dtype=torch.int32, device=dev)
    t = timeit(lambda: codec.inflate_device(src, so, dst, do, "gzip", status=st))
    assert int(st[:, 0].abs().sum().item()) == 0
    res[f"inflate_{name}"] = {"streams": len(off) - 1, "file_mb": int(off[-1]) / 1e6, "raw_mb": int(meta[:, 1].sum()) / 1e6, "ms": t, "raw_gb_s": int(meta[:, 1].sum()) / t / 1e6}
# predicted-mask PNGs written by cv2
import cv2
files = [cv2.imencode(".png", q, [cv2.IMWRITE_PNG_COMPRESSION, 3])[1].tobytes() for b in range(args.vols) for q in pat[b % len(pat)].pred_slices["axial"]]
t0 = time.perf_counter(); out = codec.png_decode_first_channel(files, dev); torch.cuda.synchronize()
res["png_decode_masks_axial_host_api"] = {"files": len(files), "file_mb": sum(map(len, files)) / 1e6, "wall_ms": (time.perf_counter() - t0) * 1e3}
print(json.dumps(res, indent=1))
