"""CPU BASELINE PORT of the reference's voxel path.  TEST / BENCH INFRASTRUCTURE - NOT PRODUCT CODE.

A restated port of the reference's Python call sequence that makes the SAME third-party calls the
reference makes (cv2.cvtColor / equalizeHist / createCLAHE / LUT, NumPy full-array reductions,
sklearn.metrics.roc_auc_score), so that its run time is what the reference's own CPU path costs
on the same host.  Used by bench.py for the `cpu_baseline` object and the `--impl reference` arm
(kind "port": /root/reference is pure Python and cannot travel to the GPU box), and by
tests/test_oracle_golden.py, which pins it against the golden vectors of the real reference.

When cv2 / scikit-learn are not importable it degrades to the NumPy restatement of
oracle/oracle.py (and says so through `BACKEND`).
"""
from __future__ import annotations

import warnings

import numpy as np

from . import oracle as O

try:  # the reference's own native dependencies
    import cv2
    cv2.setNumThreads(1)          # one worker process per core in bench.py; no nested thread pools
    _HAVE_CV2 = True
except Exception:  # pragma: no cover
    cv2 = None
    _HAVE_CV2 = False
try:
    from sklearn.metrics import roc_auc_score
    _HAVE_SK = True
except Exception:  # pragma: no cover
    roc_auc_score = None
    _HAVE_SK = False

BACKEND = ("cv2" if _HAVE_CV2 else "numpy-restatement") + "+" + ("sklearn" if _HAVE_SK else "counts-auc")


# ---- utils/utils.py:396-427 ----------------------------------------------------------------
def convertir_a_bgr(imagen):
    u = O.normalizar_a_uint8(imagen)
    return cv2.cvtColor(u, cv2.COLOR_GRAY2BGR) if u.ndim == 2 else cv2.cvtColor(u, cv2.COLOR_RGB2BGR)


def verificar_grises(imagen):
    if imagen.ndim == 3 and imagen.shape[2] == 3:
        return cv2.cvtColor(imagen, cv2.COLOR_BGR2GRAY)
    return imagen


# ---- utils/mejora_imagen.py ----------------------------------------------------------------
def aplicar_mejora(imagen, mejora):
    """<HE|CLAHE|GC|LT>().aplicar(imagen), call for call (utils/mejora_imagen.py:52-184)."""
    if not _HAVE_CV2:
        g = O.enhance_slice(imagen, mejora)
        return np.repeat(g[:, :, None], 3, axis=2)
    img_bgr = convertir_a_bgr(imagen)
    if mejora == "HE":
        yuv = cv2.cvtColor(img_bgr, cv2.COLOR_BGR2YUV)
        yuv[:, :, 0] = cv2.equalizeHist(yuv[:, :, 0])
        return cv2.cvtColor(yuv, cv2.COLOR_YUV2RGB)
    if mejora == "CLAHE":
        lab = cv2.cvtColor(img_bgr, cv2.COLOR_BGR2LAB)
        l, a, b = cv2.split(lab)
        l2 = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(l)
        return cv2.cvtColor(cv2.merge((l2, a, b)), cv2.COLOR_LAB2BGR)
    if mejora == "GC":
        table = np.array((np.linspace(0, 1, 256) ** 2.0) * 255, dtype=np.uint8)
        return cv2.LUT(img_bgr, table)
    if mejora == "LT":
        img16 = img_bgr.astype(np.uint16)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            c = 255 / np.log(1 + img16.max())
            return np.clip(c * np.log(1 + img16), 0, 255).astype(np.uint8)
    raise ValueError(f"Mejora no reconocida: {mejora}.")


def enhance_plane(vol_xyz, plano, mejora, indices=None):
    """guardar_cortes' image loop without the PNG encode (scripts/extraer_dataset.py:187-192):
    verificar_grises(aplicar(slice)) then the .T / origin='lower' orientation."""
    n = vol_xyz.shape[O.plane_axis(plano)]
    out = []
    for i in (range(n) if indices is None else indices):
        g = verificar_grises(aplicar_mejora(O.slice_of(vol_xyz, plano, i), mejora))
        out.append(np.ascontiguousarray(g.T[::-1]))
    return out


# ---- scripts/reconstruir_volumen.py:179-213 --------------------------------------------------
def reconstruir(slices, indices, shape_xyz, plano):
    return O.reconstruir(slices, indices, shape_xyz, plano)


# ---- scripts/generar_consenso.py:106-109 -----------------------------------------------------
def combinar_volumenes(ax, co, sa, umbral=2):
    return ((ax + co + sa) >= umbral).astype(np.uint8)


# ---- utils/utils.py:455-495, scripts/eval.py:115-128 -----------------------------------------
def generar_diccionario_metricas(gt_vol, pred_vol):
    if not _HAVE_SK:
        return O.generar_diccionario_metricas(gt_vol, pred_vol)
    yt, yp = gt_vol.flatten(), pred_vol.flatten()
    auc = float("nan") if len(np.unique(yt)) < 2 else float(np.round(roc_auc_score(yt, yp), 3))
    return {"DSC": O.DSC(gt_vol, pred_vol), "AUC": auc,
            "Precision": O.precision(gt_vol, pred_vol), "Recall": O.recall(gt_vol, pred_vol)}


def patient_chain(flair_zyx, gt_zyx, pred_slices, pred_indices, umbral=2, mejoras=O.MEJORAS):
    """The whole per-patient hot path as the reference runs it, in memory (float64 volumes as
    nibabel's get_fdata() yields them): enhance->slice for every enhancement and plane (ALL slices),
    recon of the three planes, consensus, metrics of the three planes and the consensus.
    Returns (number of enhanced slices, metrics dict of the consensus)."""
    vol = np.asfortranarray(flair_zyx.transpose(2, 1, 0).astype(np.float64))
    gt = np.asfortranarray(gt_zyx.transpose(2, 1, 0).astype(np.float64))
    nsl = 0
    for mej in mejoras:
        for plano in O.PLANOS:
            nsl += len(enhance_plane(vol, plano, mej))
    vols = {}
    met = {}
    for plano in O.PLANOS:
        vols[plano] = reconstruir(pred_slices[plano], pred_indices[plano], vol.shape, plano).astype(np.float64)
        met[plano] = generar_diccionario_metricas(gt, vols[plano])
    cons = combinar_volumenes(vols["axial"], vols["coronal"], vols["sagital"], umbral).astype(np.float64)
    met["consenso"] = generar_diccionario_metricas(gt, cons)
    return nsl, met
