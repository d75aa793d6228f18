"""TEST INFRASTRUCTURE - not product code.

Golden vectors for the YOLO label polygons: the ground-truth masks of the two demo patients of the
reference checkout (demo/MSLesSeg-Dataset/train/P{18,39}/T1/*_MASK.nii.gz), every lesion slice of the three planes in
the orientation guardar_cortes saves (mask.T, origin="lower"), packed to bits, together with the sha256 of the label
text that ultralytics' converter (restated in oracle/ref_stubs.py on the installed cv2.findContours) writes for it.

    python oracle/make_golden_labels.py     # -> tests/golden/demo_label_masks.npz, tests/golden/demo_labels_v1.json
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import oracle as O, ref_import, ref_stubs   # noqa: E402


def main():
    arrays, meta = {}, {}
    for pid in ("P18", "P39"):
        gt = ref_import.demo_volume(pid, "MASK")
        for plano in O.PLANOS:
            idx = O.indices_cortes_con_lesion(gt, plano)
            masks = np.stack([O.png_orient((O.slice_of(gt, plano, i) > 0).astype(np.uint8)) for i in idx])
            key = f"{pid}_{plano}"
            arrays[key + "_bits"] = np.packbits(masks, axis=None)
            arrays[key + "_shape"] = np.asarray(masks.shape, np.int32)
            arrays[key + "_idx"] = np.asarray(idx, np.int32)
            shas, ncont = [], 0
            for m in masks:
                lines = ref_stubs.yolo_seg_lines(m, 1)
                ncont += len(lines)
                shas.append(hashlib.sha256(("".join(ln + "\n" for ln in lines)).encode()).hexdigest())
            meta[key] = {"n": len(idx), "contours": ncont, "sha": shas}
            print(key, masks.shape, ncont)
    np.savez_compressed(ROOT / "tests" / "golden" / "demo_label_masks.npz", **arrays)
    (ROOT / "tests" / "golden" / "demo_labels_v1.json").write_text(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
