"""TEST INFRASTRUCTURE - generates tests/golden/* by running the REAL reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

It imports the reference's own modules (oracle/ref_import.py), feeds them
  (a) selected slices of the two demo volumes shipped with the reference
      (demo/MSLesSeg-Dataset/train/P{18,39}/T1), and
  (b) seeded synthetic volumes that any machine can regenerate bit-identically
      (yolo-mslesseg_b200/mslesseg_b200/synthetic.py),
and freezes what the reference returns:

  tests/golden/demo_slices.npz   raw slices (int16, integer-valued in the source files) and
                                 the reference's gray outputs for HE / CLAHE / GC / LT
  tests/golden/golden_v1.json    sha256 digests of whole-volume reference outputs on the
                                 synthetic patients, lesion-slice known answers, recon /
                                 consensus digests, confusion counts and metric dicts,
                                 fold statistics, fold assignment.

The reference source is never copied: only its outputs are stored.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "yolo-mslesseg_b200"))

from oracle import ref_import as ri            # noqa: E402
from mslesseg_b200 import synthetic as S       # noqa: E402

PLANOS = ("axial", "coronal", "sagital")
MEJORAS = ("HE", "CLAHE", "GC", "LT")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_slice(v, plano, i):
    if plano == "axial":
        return v[:, :, i]
    if plano == "coronal":
        return v[:, i, :]
    return v[i, :, :]


def ref_enhance_all(ref, vol_xyz_f64, plano, mejora):
    """[n][rows][cols] uint8: verificar_grises(<Mejora>().aplicar(slice)) for EVERY slice index."""
    cls = getattr(ref.mejora, mejora)
    n = vol_xyz_f64.shape[{"axial": 2, "coronal": 1, "sagital": 0}[plano]]
    return np.stack([ref.utils.verificar_grises(cls().aplicar(ref_slice(vol_xyz_f64, plano, i))) for i in range(n)])


def noise_volume(seed: int, shape_xyz=(40, 52, 36)) -> np.ndarray:
    """Small non-integer float volume [Z][Y][X] that exercises float32 rounding in E1."""
    rng = np.random.default_rng(seed)
    X, Y, Z = shape_xyz
    v = rng.standard_normal((Z, Y, X)).astype(np.float32) * np.float32(37.25) + np.float32(11.5)
    v[rng.random((Z, Y, X)) < 0.3] = np.float32(-3.0)
    v[0] = np.float32(7.0)          # one constant axial slice
    return v


def main() -> None:
    warnings.simplefilter("ignore")
    ref = ri.load_reference()
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    G = {"generator": "oracle/make_golden.py", "numpy": np.__version__}
    import cv2
    G["cv2"] = cv2.__version__

    # ---------------- (a) demo slices ----------------
    npz = {}
    demo_meta = {}
    picks = {"P18": {"axial": [60, 5], "coronal": [100, 212], "sagital": [95, 3]},
             "P39": {"axial": [84, 150], "coronal": [115, 30], "sagital": [88, 170]}}
    for pid, per_plane in picks.items():
        v = ri.demo_volume(pid, "FLAIR")
        m = ri.demo_volume(pid, "MASK")
        assert np.array_equal(v, np.round(v)) and v.max() < 32767 and v.min() >= 0
        demo_meta[pid] = {}
        for plano, idxs in per_plane.items():
            p = ref.Paciente(pid, plano, modalidad=["FLAIR"], mejora=None, gt_mask=m)
            p._volumenes["FLAIR"] = v
            lesion = p.indices_cortes_con_lesion()
            demo_meta[pid][plano] = {
                "n_lesion": len(lesion),
                "lesion_sha": sha(np.asarray(lesion, dtype=np.int32)),
                "usar20": p.indices_a_usar(20),
                "usar7": p.indices_a_usar(7),
            }
            for i in idxs:
                s = ref_slice(v, plano, i)
                npz[f"{pid}_{plano}_{i}_raw"] = np.ascontiguousarray(s).astype(np.int16)
                for mej in MEJORAS:
                    g = ref.utils.verificar_grises(getattr(ref.mejora, mej)().aplicar(s))
                    npz[f"{pid}_{plano}_{i}_{mej}"] = np.ascontiguousarray(g)
    np.savez_compressed(out_dir / "demo_slices.npz", **npz)
    G["demo"] = demo_meta

    # ---------------- (b) synthetic patients, input side ----------------
    G["synthetic_enhance"] = {}
    for pn in (1, 2):
        pat = S.make_patient(pn, config_id=1, num_cortes=20)
        v = np.asfortranarray(S.as_xyz(pat.flair).astype(np.float64))
        gt = np.asfortranarray(S.as_xyz(pat.gt).astype(np.float64))
        entry = {"seed": pat.seed, "flair_sha": sha(pat.flair), "gt_sha": sha(pat.gt), "planes": {}}
        for plano in PLANOS:
            p = ref.Paciente(pat.id, plano, modalidad=["FLAIR"], mejora="GC", gt_mask=gt)
            p._volumenes["FLAIR"] = v
            lesion = p.indices_cortes_con_lesion()
            pe = {"n_lesion": len(lesion), "lesion_sha": sha(np.asarray(lesion, dtype=np.int32)),
                  "usar20": p.indices_a_usar(20), "usar7": p.indices_a_usar(7)}
            # cortes_con_lesion_img through the reference's own class (GC, 7 slices)
            cl = p.cortes_con_lesion_img(7)["FLAIR"]
            pe["cortes_img_GC7_sha"] = sha(np.stack([ref.utils.verificar_grises(c) for _, c in cl]))
            cm = p.cortes_con_lesion_mask(7)
            pe["cortes_mask7_sha"] = sha(np.stack([c for _, c in cm]).astype(np.uint8))
            for mej in MEJORAS:
                pe[mej] = sha(ref_enhance_all(ref, v, plano, mej))
            entry["planes"][plano] = pe
        G["synthetic_enhance"][pat.id] = entry

    # small non-integer float volume
    nv = noise_volume(777)
    nvx = np.asfortranarray(S.as_xyz(nv).astype(np.float64))
    G["noise_enhance"] = {"seed": 777, "shape_xyz": [40, 52, 36], "in_sha": sha(nv),
                          "planes": {pl: {mej: sha(ref_enhance_all(ref, nvx, pl, mej)) for mej in MEJORAS}
                                     for pl in PLANOS}}

    # ---------------- (c) output side: recon -> consensus -> eval ----------------
    from PIL import Image
    G["synthetic_eval"] = {}
    fold_metricas = {}
    # stub the two NIfTI helpers the reference's reconstruir_volumen() calls
    ref.recon.cargar_referencia_nifti = lambda p: (S.SHAPE_XYZ, np.eye(4))
    ref.recon.guardar_volumen = lambda volumen, affine, output_path: None
    for pn in (54, 55, 56):
        pat = S.make_patient(pn, config_id=2, num_cortes=20)
        gt = np.asfortranarray(S.as_xyz(pat.gt).astype(np.float64))
        entry = {"seed": pat.seed, "gt_sha": sha(pat.gt), "planes": {}}
        vols = {}
        for plano in PLANOS:
            with tempfile.TemporaryDirectory() as td:
                for i, q in zip(pat.pred_indices[plano], pat.pred_slices[plano]):
                    Image.fromarray(q).save(os.path.join(td, f"{pat.id}_FLAIR_{i}.png"))
                vol = ref.recon.reconstruir_volumen(Path(td), "gt.nii.gz", "out.nii.gz", plano)
            assert vol.dtype == np.float32 and vol.shape == S.SHAPE_XYZ
            vols[plano] = vol.astype(np.float64)      # what cargar_volumen(get_fdata) hands to the next stage
            u8 = np.ascontiguousarray(vol.transpose(2, 1, 0)).astype(np.uint8)   # [Z][Y][X]
            met = ref.eval.generar_diccionario_metricas(gt, vols[plano])
            tp = int(np.sum((gt == 1) & (vols[plano] == 1))); fp = int(np.sum((gt == 0) & (vols[plano] == 1)))
            fn = int(np.sum((gt == 1) & (vols[plano] == 0))); tn = int(np.sum((gt == 0) & (vols[plano] == 0)))
            entry["planes"][plano] = {"indices": pat.pred_indices[plano], "slices_sha": sha(pat.pred_slices[plano]),
                                      "recon_u8_sha": sha(u8), "recon_f32_sha": sha(np.ascontiguousarray(vol.transpose(2, 1, 0))),
                                      "counts": [tp, fp, fn, tn], "metricas": met}
        for umbral in (2, 3):
            c = ref.consenso.combinar_volumenes(vols["axial"], vols["coronal"], vols["sagital"], umbral)
            assert c.dtype == np.uint8
            cf = c.astype(np.float64)
            met = ref.eval.generar_diccionario_metricas(gt, cf)
            tp = int(np.sum((gt == 1) & (cf == 1))); fp = int(np.sum((gt == 0) & (cf == 1)))
            fn = int(np.sum((gt == 1) & (cf == 0))); tn = int(np.sum((gt == 0) & (cf == 0)))
            entry[f"consenso{umbral}"] = {"sha": sha(np.ascontiguousarray(c.transpose(2, 1, 0))),
                                          "counts": [tp, fp, fn, tn], "metricas": met}
        G["synthetic_eval"][pat.id] = entry
        for k, val in entry["consenso2"]["metricas"].items():
            fold_metricas.setdefault(k, []).append(val)

    # empty-GT and empty-prediction edge cases
    gt0 = np.zeros((6, 5, 4)); pr = np.zeros((6, 5, 4)); pr[1, 2, 3] = 1
    gt1 = np.zeros((6, 5, 4)); gt1[0, 0, 0] = 1; gt1[1, 2, 3] = 1
    nan2none = lambda d: {k: (None if (isinstance(v, float) and np.isnan(v)) else v) for k, v in d.items()}
    G["edge_metricas"] = {
        "empty_gt": nan2none(ref.eval.generar_diccionario_metricas(gt0, pr)),
        "empty_pred": nan2none(ref.eval.generar_diccionario_metricas(gt1, np.zeros((6, 5, 4)))),
        "perfect": nan2none(ref.eval.generar_diccionario_metricas(gt1, gt1.copy())),
    }

    # ---------------- (d) fold statistics / fold assignment ----------------
    G["promedio_fold"] = {"in": fold_metricas, "out": ref.eval.calcular_promedio(fold_metricas)}
    folds_in = {"DSC": [0.612, 0.587, 0.655, 0.601, 0.59], "AUC": [0.81, 0.79, 0.83, 0.8, 0.805],
                "Precision": [0.7, 0.66, 0.71, 0.69, 0.72], "Recall": [0.55, 0.53, 0.61, 0.54, 0.5]}
    G["resumen_experimento"] = {"in": folds_in, "out": ref.promediar.calcular_resumen_experimento(folds_in)}
    G["calcular_fold"] = {str(k): {f"P{n}": ref.utils.calcular_fold(f"P{n}", k) for n in range(1, 54)} for k in (5, 3)}
    G["percentil"] = {"in": [28, 21, 23, 101, 147, 113, 40], "P50": int(np.percentile([28, 21, 23, 101, 147, 113, 40], 50)),
                      "P25": int(np.percentile([28, 21, 23, 101, 147, 113, 40], 25))}

    # random-count AUC known answers through the real sklearn call the reference makes
    rng = np.random.default_rng(5)
    auc_cases = []
    for _ in range(40):
        n = int(rng.integers(20, 400))
        yt = (rng.random(n) < rng.uniform(0.05, 0.6)).astype(np.float64)
        yp = (rng.random(n) < rng.uniform(0.05, 0.6)).astype(np.float64)
        if len(np.unique(yt)) < 2:
            continue
        tp = int(np.sum((yt == 1) & (yp == 1))); fp = int(np.sum((yt == 0) & (yp == 1)))
        fn = int(np.sum((yt == 1) & (yp == 0))); tn = int(np.sum((yt == 0) & (yp == 0)))
        auc_cases.append({"counts": [tp, fp, fn, tn], "metricas": ref.eval.generar_diccionario_metricas(yt, yp)})
    G["random_metricas"] = auc_cases

    with open(out_dir / "golden_v1.json", "w") as f:
        json.dump(G, f, indent=1, sort_keys=True)
    print("wrote", out_dir / "golden_v1.json", os.path.getsize(out_dir / "golden_v1.json"), "bytes;",
          "demo_slices.npz", os.path.getsize(out_dir / "demo_slices.npz"), "bytes")


if __name__ == "__main__":
    main()
