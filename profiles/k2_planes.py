#!/usr/bin/env python
"""Which output of norm_scatter costs what: times the kernel (library-side CUDA events) with subsets of the three planes
enabled.  usage (GPU box): python profiles/k2_planes.py [nvol]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "yolo-mslesseg_b200"))
import torch
from mslesseg_b200 import ops, _lib as L, synthetic as S

nvol = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
vols = torch.stack([torch.from_numpy(S.make_patient(i + 1, 1, 40, with_predictions=False).flair) for i in range(4)]).to(dev)
vol = vols.repeat((nvol + 3) // 4, 1, 1, 1)[:nvol].contiguous()
for planos in (("axial",), ("coronal",), ("sagital",), ("axial", "coronal"), ("axial", "coronal", "sagital")):
    for it in range(3):
        if it == 2:
            L.profile_enable(True)
        ops.enhance_volumes(vol, mejoras=("GC",), planos=planos)
    torch.cuda.synchronize()
    prof = L.profile_collect()
    L.profile_enable(False)
    print(f"{'+'.join(planos):24s} norm_scatter {prof['norm_scatter'][0]:.4f} ms   plane_stats {prof['plane_stats_f32'][0]:.4f} ms   ({nvol} volumes)")
