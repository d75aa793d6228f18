"""TEST INFRASTRUCTURE - container only.  Freezes outputs of the reference's own prediction post-processing
(scripts/generar_predicciones.py: combinar_predicciones + normalizar_prediccion, which call cv2.resize / cv2.flip)
on seeded instance masks -> tests/golden/golden_pred_v1.json (sha256 of the uint8 results).

usage: python oracle/make_golden_pred.py        (needs /root/reference; see oracle/ref_import.py)"""
import hashlib
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import as ri  # noqa: E402

CASES = [  # (seed, n_instances, mask h, mask w, image height, image width)
    (1, 3, 640, 544, 218, 182), (2, 1, 640, 640, 182, 182), (3, 5, 544, 640, 182, 218), (4, 0, 640, 544, 218, 182),
    (5, 2, 160, 160, 218, 182), (6, 4, 97, 131, 45, 37), (7, 2, 20, 14, 64, 80), (8, 7, 320, 288, 218, 182),
]


def instance_masks(seed, n, mh, mw):
    """Blobby float masks in [0, 1] like ultralytics' `masks.data` (values exactly 0 / 1 plus a few soft ones)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, mh, mw), dtype=np.float32)
    yy, xx = np.mgrid[0:mh, 0:mw]
    for i in range(n):
        cy, cx = rng.uniform(0.2, 0.8) * mh, rng.uniform(0.2, 0.8) * mw
        ry, rx = rng.uniform(0.02, 0.15) * mh + 1, rng.uniform(0.02, 0.15) * mw + 1
        d = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
        out[i] = (d <= 1.0).astype(np.float32)
        soft = rng.random((mh, mw)) < 0.01
        out[i][soft] = rng.random(int(soft.sum()), dtype=np.float32)       # includes values around the 0.5 threshold
    return out


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ri.install_stubs()
    sys.path.insert(0, ri.REFERENCE_ROOT)
    os.makedirs("/tmp/mslesseg_ref_cwd", exist_ok=True)
    os.chdir("/tmp/mslesseg_ref_cwd")
    gp = importlib.import_module("yolo_mslesseg.scripts.generar_predicciones")
    out = {"cases": []}
    for seed, n, mh, mw, h, w in CASES:
        masks = instance_masks(seed, n, mh, mw)
        comb = gp.combinar_predicciones(list(masks), (h, w))
        norm = gp.normalizar_prediccion(comb.copy())
        out["cases"].append({"seed": seed, "n": n, "mask_shape": [mh, mw], "image_shape": [h, w],
                             "combined_sha": sha(comb), "combined_sum": int(comb.sum()),
                             "normalised_sha": sha(norm), "normalised_shape": list(norm.shape)})
    dst = os.path.join(ROOT, "tests", "golden", "golden_pred_v1.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", dst, len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
