"""TEST INFRASTRUCTURE - not product code.

Runs the UNMODIFIED reference stage scripts (the original checkout at /root/reference when present, else the copy
staged by oracle/build_ref.py) on the synthetic dataset tree of oracle/stock_pipeline.py and freezes a digest of every
artefact they write (decoded PNG pixels, NIfTI data, JSON values, label text) plus the stages' tri-state results:

    python oracle/make_golden_stock.py        # -> tests/golden/stock_pipeline_v1.json

Third-party stand-ins: oracle/ref_stubs.py (nibabel file format; matplotlib imsave = oracle restatement, E8 UNPINNED;
ultralytics label text restated on the installed cv2.findContours).
"""
import hashlib
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import build_ref, stock_pipeline as SP   # noqa: E402

CASES = {
    "CLAHE_P50": dict(mejora="CLAHE", num_cortes="P50"),
    "HE_12": dict(mejora="HE", num_cortes=12),
    "Base_P50": dict(mejora=None, num_cortes="P50"),
}
IDS = ("P3", "P9", "P20", "P30", "P41")


def canonical(snap):
    return json.dumps(snap, sort_keys=True, default=str)


def run_case(ns, name, **kw):
    pats = SP.synthetic_patients(ids=IDS)
    root = tempfile.mkdtemp(prefix=f"stock_{name}_")
    SP.make_tree(root, pats)
    states = SP.run_all(ns, root, pats, k_folds=3, **kw)
    snap = SP.snapshot(root)
    return root, pats, states, snap


def digest_of(snap, states):
    groups = {}
    for k, v in snap.items():
        kind = "png_images" if "/images/" in k else "png_gt_masks" if "/GT_masks/" in k else "png_pred_masks" if "/pred_masks/" in k else \
            "labels" if k.endswith(".txt") else "nifti" if k.endswith(".nii.gz") else "json" if k.endswith(".json") else "other"
        groups.setdefault(kind, {})[k] = v
    return {"n_files": len(snap),
            "sha": {g: hashlib.sha256(canonical(d).encode()).hexdigest() for g, d in sorted(groups.items())},
            "count": {g: len(d) for g, d in sorted(groups.items())},
            "states": {"|".join(map(str, k)): v for k, v in sorted(states.items(), key=lambda kv: str(kv[0]))},
            "json": groups.get("json", {})}


def main():
    ns = build_ref.load()
    out = {"reference_root": ns.root, "ids": list(IDS), "cases": {}}
    for name, kw in CASES.items():
        _, _, states, snap = run_case(ns, name, **kw)
        out["cases"][name] = digest_of(snap, states)
        print(name, out["cases"][name]["count"])
    path = ROOT / "tests" / "golden" / "stock_pipeline_v1.json"
    path.write_text(json.dumps(out, indent=1, sort_keys=True, default=str))
    print("wrote", path)


if __name__ == "__main__":
    main()
