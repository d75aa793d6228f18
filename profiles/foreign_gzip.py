"""One-warp inflate of a single-member .nii.gz written by the host's gzip (what the dataset ships), against the host decoder."""
import sys, time, gzip, zlib
sys.path[:0] = ['/root/repo', '/root/repo/yolo-mslesseg_b200']
import numpy as np, torch
from mslesseg_b200 import _lib, codec, nifti, synthetic as S
_lib.load()
dev = torch.device('cuda:0')
pat = S.make_patient(1, config_id=4, num_cortes=40)
hdr = codec.nifti_header_bytes((182, 218, 182), np.float32, np.eye(4))
raw = hdr + pat.flair.astype(np.float32).tobytes()
for lvl in (1, 6, 9):
    t0 = time.perf_counter(); blob = gzip.compress(raw, lvl); tc = time.perf_counter() - t0
    t0 = time.perf_counter(); back = gzip.decompress(blob); th = time.perf_counter() - t0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dst, off = codec.inflate([blob], [len(raw)], 'gzip', dev)
    torch.cuda.synchronize(); tg = time.perf_counter() - t0
    ok = dst[:len(raw)].cpu().numpy().tobytes() == raw
    print(f"level {lvl}: file {len(blob)/1e6:.1f} MB, host compress {tc:.2f}s decompress {th:.3f}s, GPU one-warp inflate {tg:.3f}s ok={ok}")
