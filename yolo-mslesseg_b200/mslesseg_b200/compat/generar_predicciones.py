"""Mirror of the numeric part of yolo_mslesseg/scripts/generar_predicciones.py (:123-140): the post-processing of the
YOLO instance masks.  The network call itself (`ejecutar_prediccion`, :111-120) stays with ultralytics."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from . import device


def _stack(predicciones):
    if isinstance(predicciones, torch.Tensor):
        return predicciones.to(device=device(), dtype=torch.float32).contiguous()
    preds = [np.asarray(p, dtype=np.float32) for p in predicciones]
    if not preds:
        return None
    shapes = {p.shape for p in preds}
    if len(shapes) != 1 or len(preds[0].shape) != 2:
        raise ValueError("todas las predicciones deben ser máscaras 2D de la misma resolución")
    return torch.from_numpy(np.stack(preds)).to(device())


def combinar_predicciones(predicciones, shape):
    """OR of the instance masks (> 0.5), nearest-resized to `shape` = (height, width): uint8 {0, 1}."""
    height, width = (int(d) for d in shape)
    masks = _stack(predicciones)
    n_inst = 0 if masks is None else int(masks.shape[0])
    dev = device()
    out = torch.empty((1, height, width), dtype=torch.uint8, device=dev)
    # the image the model saw is PNG-oriented: (height, width) = (cols, rows) of the slice
    ops.combine_predictions(masks, [0, n_inst], rows=width, cols=height, layout="P", out=out)
    return out[0].cpu().numpy()


def normalizar_prediccion(pred):
    """cv2.flip(pred.T, 1) * 255 for the {0, 1} mask `combinar_predicciones` returns: (height, width) -> (width, height),
    values {0, 255}.  Runs through the same kernel (identity resize, slice-oriented output)."""
    pred = np.asarray(pred)
    if pred.ndim != 2 or pred.max(initial=0) > 1 or pred.min(initial=0) < 0:
        raise ValueError("normalizar_prediccion: se espera la máscara 2D {0, 1} de combinar_predicciones")
    height, width = pred.shape
    masks = torch.from_numpy(np.ascontiguousarray(pred, dtype=np.float32))[None].to(device())
    return ops.combine_predictions(masks, [0, 1], rows=width, cols=height, layout="G")[0].cpu().numpy()


def predicciones_a_cortes(masks, inst_offset, rows, cols):
    """Batched, fused form used by the accelerated pipeline: all slices of a patient / plane at once, straight to the
    {0, 255} slice-oriented masks `reconstruir_volumen` consumes (device tensor [n, rows, cols])."""
    return ops.combine_predictions(masks, inst_offset, rows, cols, layout="G")
