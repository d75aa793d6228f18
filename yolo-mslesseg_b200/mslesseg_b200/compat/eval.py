"""Mirror of the numeric part of yolo_mslesseg/scripts/eval.py (:115-160) and promediar_folds.py (:87-134)."""
from __future__ import annotations

import logging
from pathlib import Path

from .. import metrics as _M
from .utils import cargar_volumen, cargar_volumen_dispositivo, leer_json, metricas, reconstruccion_valida

logger = logging.getLogger(__name__)


def generar_diccionario_metricas(gt_vol, pred_vol):
    """{"DSC", "AUC", "Precision", "Recall"} rounded to 3 dp: one pass of voxel counts on the GPU, the float64
    formulas of utils/utils.py:455-495 on the host."""
    return metricas(gt_vol, pred_vol)


def calcular_metricas(gt_vol_path, pred_vol_path):
    if not reconstruccion_valida(pred_vol_path, gt_vol_path):
        logger.warning(f"⚠️ Reconstrucción inválida: {Path(pred_vol_path).name}")
        return {}
    # both files are inflated on the GPU as uint8 masks (non-integral / out-of-range voxels raise) and counted there
    import torch
    from .. import ops
    gt = cargar_volumen_dispositivo(gt_vol_path, torch.uint8).reshape(1, -1)
    pred = cargar_volumen_dispositivo(pred_vol_path, torch.uint8).reshape(1, -1)
    c = ops.confusion_counts(gt, pred)[0].cpu().numpy()
    if int(c.sum()) != gt.numel():
        return generar_diccionario_metricas(cargar_volumen(gt_vol_path), cargar_volumen(pred_vol_path))   # non-binary masks: exact == 1 / == 0 predicates
    return _M.metricas_desde_conteos(*(int(x) for x in c))


calcular_promedio = _M.calcular_promedio
calcular_resumen_experimento = _M.calcular_resumen_experimento


def agregar_metricas_fold(dic_total, archivo):
    for k, v in leer_json(archivo).items():
        if isinstance(v, dict) and "media" in v:
            dic_total.setdefault(k, []).append(v["media"])
        elif isinstance(v, (int, float)):
            dic_total.setdefault(k, []).append(float(v))
        else:
            logger.warning(f"⚠️ Formato inesperado para la métrica '{k}' en {archivo}: {v}")
