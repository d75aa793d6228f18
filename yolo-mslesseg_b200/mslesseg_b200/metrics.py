"""Host side of R4-R7: voxel confusion counts -> DSC / AUC / precision / recall and the fold /
experiment statistics.  The counts come from the CUDA kernels (ops.consensus_eval /
ops.confusion_counts); turning four integers into four rounded floats stays in NumPy float64 so
that every intermediate rounds exactly like the reference.

Reference: utils/utils.py:455-495 (DSC, precision, recall, AUC), scripts/eval.py:115-160,
scripts/promediar_folds.py:126-134, utils/utils.py:299-316 (calcular_fold).
"""
from __future__ import annotations

import logging

import numpy as np

logger = logging.getLogger(__name__)


def auc_binario(tp: int, fp: int, fn: int, tn: int) -> float:
    """sklearn.metrics.roc_auc_score(y_true, y_pred) for {0,1}-valued y_pred, from the counts.
    scikit-learn's path (_binary_clf_curve -> roc_curve -> auc/np.trapezoid) sees the ROC points
    (0,0), (fp/(fp+tn), tp/(tp+fn)), (1,1); the middle one is absent when y_pred has one value."""
    fps = np.array([0.0, float(fp), float(fp + tn)])
    tps = np.array([0.0, float(tp), float(tp + fn)])
    if (tp + fp) == 0 or (fn + tn) == 0:
        fps, tps = fps[[0, 2]], tps[[0, 2]]
    fpr = fps / fps[-1]
    tpr = tps / tps[-1]
    return float((np.diff(fpr) * (tpr[1:] + tpr[:-1]) / 2.0).sum())


def iou_desde_conteos(tp: int, fp: int, fn: int) -> float:
    """Jaccard index tp / (tp + fp + fn) of two binary arrays from their counts.  NOT a reference metric (the reference
    computes DSC / AUC / precision / recall only, scripts/eval.py:121-126); BASELINE.json's north_star names it, so it is
    offered with the reference's conventions: + 1e-8 in the denominator, float64, np.round(., 3)."""
    tp64, fp64, fn64 = np.int64(int(tp)), np.int64(int(fp)), np.int64(int(fn))
    return float(np.round(tp64 / (tp64 + fp64 + fn64 + 1e-8), 3))


def calcular_rango_global(rangos, cortes=None):
    """extras/generar_gif_predicciones.py:141-148 on per-slice (min, max) pairs [n, 2] (ops.slice_ranges of one patient
    and plane): (global minimum, global maximum) over the listed slices (default: all)."""
    r = np.asarray(rangos)
    if cortes is not None:
        r = r[list(cortes)]
    if r.shape[0] == 0:
        raise ValueError("min() arg is an empty sequence")
    return r[:, 0].min(), r[:, 1].max()


def metricas_desde_conteos(tp: int, fp: int, fn: int, tn: int, con_iou: bool = False) -> dict:
    """generar_diccionario_metricas (scripts/eval.py:115-128) for binary volumes.
    sum(gt*pred) = tp, sum(gt) = tp+fn, sum(pred) = tp+fp are exact in float64.
    con_iou=True adds an "IoU" key (iou_desde_conteos); off by default so that the JSON files equal the reference's."""
    tp, fp, fn, tn = int(tp), int(fp), int(fn), int(tn)
    tp64, fp64, fn64 = np.int64(tp), np.int64(fp), np.int64(fn)
    dsc = (2.0 * np.float64(tp)) / (np.float64(tp + fn) + np.float64(tp + fp) + 1e-8)
    prec = tp64 / (tp64 + fp64 + 1e-8)
    rec = tp64 / (tp64 + fn64 + 1e-8)
    if (tp + fn) == 0 or (fp + tn) == 0:
        logger.warning("⚠️ AUC no definido: y_true contiene una sola clase.")
        auc = float("nan")
    else:
        auc = float(np.round(auc_binario(tp, fp, fn, tn), 3))
    out = {
        "DSC": float(np.round(dsc, 3)),
        "AUC": auc,
        "Precision": float(np.round(prec, 3)),
        "Recall": float(np.round(rec, 3)),
    }
    if con_iou:
        out["IoU"] = iou_desde_conteos(tp, fp, fn)
    return out


def dsc_desde_conteos(tp: int, fp: int, fn: int) -> float:
    """DSC (utils/utils.py:455-460) of two binary arrays from their counts, rounded to 3 decimals like the reference."""
    tp, fp, fn = int(tp), int(fp), int(fn)
    return float(np.round((2.0 * np.float64(tp)) / (np.float64(tp + fn) + np.float64(tp + fp) + 1e-8), 3))


def seleccionar_mejor_corte(conteos_plano, cortes=None):
    """The selection loop of extras/visualizar_prediccion_corte.py:150-182 on per-slice counts [n_plane, 4] of one
    patient and plane: DSC of every listed slice (default: all), the first strictly larger value wins.
    Returns (corte, dsc); (None, -1.0) for an empty list."""
    conteos = np.asarray(conteos_plano)
    mejor_corte, mejor_dsc = None, -1.0
    for corte in (range(conteos.shape[0]) if cortes is None else cortes):
        tp, fp, fn = (int(v) for v in conteos[int(corte), :3])
        dsc = dsc_desde_conteos(tp, fp, fn)
        if dsc > mejor_dsc:
            mejor_dsc, mejor_corte = dsc, int(corte)
    return mejor_corte, mejor_dsc


def calcular_promedio(metricas_dic: dict) -> dict:
    """scripts/eval.py:144-160: mean and population std of the patients' (already rounded) metrics."""
    if not metricas_dic:
        raise ValueError("El diccionario de métricas está vacío.")
    return {m: {"media": float(np.round(np.mean(v), 3)), "std": float(np.round(np.std(v), 3))}
            for m, v in metricas_dic.items()}


def calcular_resumen_experimento(metricas_fold: dict) -> dict:
    """scripts/promediar_folds.py:126-134: mean and SAMPLE std (ddof=1) of the fold means."""
    return {m: {"media": float(np.round(np.mean(v), 3)), "std": float(np.round(np.std(v, ddof=1), 3))}
            for m, v in metricas_fold.items()}


def calcular_fold(paciente_id: str, k_folds: int = 5, n_ids: int = 53) -> int:
    """utils/utils.py:299-316.  n_ids = 53 is the reference behaviour (P1..P53, raises beyond);
    larger synthetic cohorts pass their own size."""
    numero = int(paciente_id[1:])
    folds = np.array_split(list(range(1, n_ids + 1)), k_folds)
    for i, fold in enumerate(folds, 1):
        if numero in fold:
            return i
    raise ValueError(f"No se puede calcular el fold del paciente {paciente_id}.")


def ventana_central(indices_validos, num_cortes):
    """List arithmetic of Paciente.indices_a_usar (utils/Paciente.py:261-275)."""
    indices_validos = list(indices_validos)
    if num_cortes is None or len(indices_validos) <= num_cortes:
        return indices_validos
    centro = len(indices_validos) // 2
    mitad = num_cortes // 2
    start = max(0, centro - mitad)
    return indices_validos[start:start + num_cortes]


def num_cortes_percentil(conteos, percentil: int = 50) -> int:
    """scripts/extraer_dataset.py:110-135: int(np.percentile(lesion-slice counts, p))."""
    conteos = list(conteos)
    if not conteos:
        raise ValueError("No se encontraron cortes con lesión válidos para calcular el percentil.")
    try:
        return int(np.percentile(conteos, percentil))
    except Exception as e:  # same wrapping as the reference
        raise ValueError(f"Percentil no válido ({percentil}): {e}")
