"""TEST INFRASTRUCTURE - not product code.

Drives the reference's own, unmodified stage entry points on a small synthetic dataset tree:

    ejecutar_dataset_pipeline            scripts/extraer_dataset.py:486
    (stage 3, YOLO prediction, is replaced by fabricated pred_masks PNGs written with cv2.imwrite exactly like
     guardar_prediccion, scripts/generar_predicciones.py:143-153)
    ejecutar_reconstrucciones_pipeline   scripts/reconstruir_volumen.py:486
    ejecutar_eval_pipeline               scripts/eval.py:417
    ejecutar_consenso_pipeline           scripts/generar_consenso.py:378   (+ eval with plano="consenso")
    ejecutar_promediar_folds_pipeline    scripts/promediar_folds.py:293

in the order of ejecutar_pipeline.py:385-442 (--completo mode).  The same driver runs twice in the tests: once on
the stock modules (third-party stand-ins of oracle/ref_stubs.py) and once after
`mslesseg_b200.compat.install.install()` rebinds the hot-path functions to the CUDA library; `snapshot()` decodes
every artefact (PNG pixels, NIfTI dtype / shape / data, JSON values, label text) so the two trees can be compared.
"""
from __future__ import annotations

import contextlib
import gzip
import hashlib
import json
import os
import struct
from pathlib import Path

import numpy as np

PLANOS = ("axial", "coronal", "sagital")


@contextlib.contextmanager
def chdir(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def synthetic_patients(ids=("P3", "P20", "P41"), shape=(62, 74, 58), seed=7):
    """Small skull-stripped-like volumes (integer-valued float32, ~a third brain) with blob lesions;
    ids fall into different folds of calcular_fold(k_folds=3)."""
    X, Y, Z = shape
    out = {}
    xs, ys, zs = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij")
    for n, pid in enumerate(ids):
        rng = np.random.default_rng(seed + n)
        brain = (((xs - X / 2) / (0.40 * X)) ** 2 + ((ys - Y / 2) / (0.42 * Y)) ** 2 + ((zs - 0.45 * Z) / (0.43 * Z)) ** 2) <= 1.0
        scale = rng.uniform(150, 600)
        flair = np.where(brain, np.round(np.clip(scale * (1 + 0.25 * rng.standard_normal(shape)), 1, 2.4 * scale)), 0).astype(np.float32)
        gt = np.zeros(shape, np.uint8)
        for _ in range(int(rng.integers(4, 9))):
            c = [int(rng.integers(int(0.3 * d), int(0.7 * d))) for d in shape]
            r = int(rng.integers(2, 6))
            gt |= ((xs - c[0]) ** 2 + (ys - c[1]) ** 2 + (zs - c[2]) ** 2 <= r * r).astype(np.uint8)
        gt &= brain.astype(np.uint8)
        flair = np.where(gt > 0, np.round(flair * 1.4), flair).astype(np.float32)
        out[pid] = (np.asfortranarray(flair), np.asfortranarray(gt))
    return out


def write_nifti(path, vol, affine=None, dtype=np.float32):
    """Plain NIfTI-1 .nii.gz writer for the input tree (float32 data like the MSLesSeg files)."""
    vol = np.asarray(vol).astype(dtype)
    affine = np.diag([1.0, 1.0, 1.0, 1.0]) if affine is None else affine
    code = {"float32": 16, "uint8": 2, "float64": 64, "int16": 4}[vol.dtype.name]
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, *vol.shape, 1, 1, 1, 1)
    struct.pack_into("<2h", hdr, 70, code, vol.dtype.itemsize * 8)
    struct.pack_into("<8f", hdr, 76, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, 1.0, 0.0)
    struct.pack_into("<2h", hdr, 252, 0, 2)
    for r in range(3):
        struct.pack_into("<4f", hdr, 280 + 16 * r, *[float(x) for x in affine[r]])
    hdr[344:348] = b"n+1\x00"
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    with gzip.open(str(path), "wb", compresslevel=1) as f:
        f.write(bytes(hdr) + b"\x00" * 4 + np.asfortranarray(vol).tobytes(order="F"))


def make_tree(root, patients):
    """MSLesSeg-Dataset/train/PX/T1/PX_T1_{FLAIR,MASK}.nii.gz + GT/train/PX/PX_MASK.nii.gz (scripts/setup.py:217)."""
    root = Path(root)
    aff = np.array([[1.0, 0, 0, -90], [0, 1.0, 0, -126], [0, 0, 1.0, -72], [0, 0, 0, 1.0]])
    for pid, (flair, gt) in patients.items():
        d = root / "MSLesSeg-Dataset" / "train" / pid / "T1"
        write_nifti(d / f"{pid}_T1_FLAIR.nii.gz", flair, aff)
        write_nifti(d / f"{pid}_T1_MASK.nii.gz", gt, aff)
        write_nifti(root / "GT" / "train" / pid / f"{pid}_MASK.nii.gz", gt, aff)


def fabricate_pred_masks(root, modelo, k_folds, patients, calcular_fold, seed=11):
    """Stage 3 stand-in: for every image PNG of the dataset stage write a predicted mask (GT with seeded drop /
    add noise) in slice orientation, uint8 {0, 255}, single channel, cv2.imwrite(..., PNG_COMPRESSION 3), named
    like the image (scripts/generar_predicciones.py:143-153, 225-236)."""
    import cv2
    root = Path(root)
    plano = modelo.plano
    for pid, (_, gt) in patients.items():
        fold = calcular_fold(paciente_id=pid, k_folds=k_folds)
        pdir = root / "datasets" / modelo.base_path / f"fold{fold}" / pid / plano
        (pdir / "pred_masks").mkdir(parents=True, exist_ok=True)
        rng = np.random.default_rng(seed + int(pid[1:]) * 3 + PLANOS.index(plano))
        for img in sorted((pdir / "images").glob("*.png")):
            i = int(img.stem.split("_")[-1])
            sl = {"axial": gt[:, :, i], "coronal": gt[:, i, :], "sagital": gt[i, :, :]}[plano]
            keep = rng.random(sl.shape) > 0.25
            add = rng.random(sl.shape) < 0.002
            pred = (((sl > 0) & keep) | add).astype(np.uint8) * 255
            cv2.imwrite(str(pdir / "pred_masks" / img.name), pred, [cv2.IMWRITE_PNG_COMPRESSION, 3])


def run_all(ns, root, patients, mejora="CLAHE", num_cortes="P50", epochs=50, k_folds=3, umbral=2, planos=PLANOS,
            repeat_stage=None):
    """Runs the stage entry points for every plane, then consensus and fold averaging.  Returns the tri-state
    results the stages computed (evaluar_resultados, utils/utils.py:435-447), keyed by (stage, plano, fold)."""
    states = {}
    mods = {"dataset": ns.extraer, "recon": ns.recon, "consenso": ns.consenso}
    originals = {k: m.evaluar_resultados for k, m in mods.items()}
    current = {"key": None}

    def recorder(stage):
        def rec(resultados):
            r = originals[stage](resultados)
            states[(stage,) + tuple(current["key"])] = r
            return r
        return rec

    for k, m in mods.items():
        m.evaluar_resultados = recorder(k)
    try:
        with chdir(root):
            for plano in planos:
                modelo = ns.Modelo(plano=plano, num_cortes=num_cortes, modalidad=["FLAIR"], k_folds=k_folds, mejora=mejora)
                current["key"] = (plano, 0)
                ns.extraer.ejecutar_dataset_pipeline(modelo=modelo, k_folds=k_folds)
                if repeat_stage == "dataset":
                    current["key"] = (plano, -1)
                    ns.extraer.ejecutar_dataset_pipeline(modelo=modelo, k_folds=k_folds)     # skip-if-exists
                fabricate_pred_masks(root, modelo, k_folds, patients, ns.utils.calcular_fold)
                for fold in range(1, k_folds + 1):
                    current["key"] = (plano, fold)
                    ns.recon.ejecutar_reconstrucciones_pipeline(modelo=modelo, fold_test=fold, epochs=epochs, k_folds=k_folds)
                    ns.eval.ejecutar_eval_pipeline(modelo=modelo, fold_test=fold, epochs=epochs, k_folds=k_folds)
                ns.promediar.ejecutar_promediar_folds_pipeline(modelo=modelo, epochs=epochs, k_folds=k_folds)
            if set(planos) == set(PLANOS):
                modelo = ns.Modelo(plano="axial", num_cortes=num_cortes, modalidad=["FLAIR"], k_folds=k_folds, mejora=mejora)
                for fold in range(1, k_folds + 1):
                    current["key"] = ("consenso", fold)
                    ns.consenso.ejecutar_consenso_pipeline(modelo=modelo, epochs=epochs, k_folds=k_folds, umbral=umbral, fold_test=fold)
                    ns.eval.ejecutar_eval_pipeline(modelo=modelo, plano="consenso", epochs=epochs, k_folds=k_folds, fold_test=fold)
                ns.promediar.ejecutar_promediar_folds_pipeline(modelo=modelo, plano="consenso", epochs=epochs, k_folds=k_folds)
    finally:
        for k, m in mods.items():
            m.evaluar_resultados = originals[k]
    return states


def _nifti_content(path):
    with gzip.open(str(path), "rb") as f:
        raw = f.read()
    dim = struct.unpack("<8h", raw[40:56])
    datatype = struct.unpack("<h", raw[70:72])[0]
    off = int(struct.unpack("<f", raw[108:112])[0])
    dt = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8"}[datatype]
    shape = tuple(dim[1:1 + dim[0]])
    data = np.frombuffer(raw, dtype="<" + dt, count=int(np.prod(shape)), offset=off)
    sform = [struct.unpack("<4f", raw[280 + 16 * r:296 + 16 * r]) for r in range(3)]
    return {"dtype": dt, "shape": shape, "sha": hashlib.sha256(data.tobytes()).hexdigest(), "nonzero": int(np.count_nonzero(data)),
            "affine": [[round(float(x), 4) for x in row] for row in sform]}


def snapshot(root, skip=("MSLesSeg-Dataset", "GT")):
    """{relative path: decoded content} of every artefact the stages wrote."""
    from PIL import Image
    root = Path(root)
    snap = {}
    for p in sorted(root.rglob("*")):
        if not p.is_file():
            continue
        rel = p.relative_to(root).as_posix()
        if rel.split("/")[0] in skip or p.name.endswith(".log"):
            continue
        if p.suffix == ".png":
            im = Image.open(p)
            a = np.array(im)
            snap[rel] = {"mode": im.mode, "shape": a.shape, "sha": hashlib.sha256(a.tobytes()).hexdigest()}
        elif p.name.endswith(".nii.gz"):
            snap[rel] = _nifti_content(p)
        elif p.suffix == ".json":
            snap[rel] = json.loads(p.read_text())
        else:
            snap[rel] = p.read_text()
    return snap
