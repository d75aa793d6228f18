"""Mirror of the numeric part of yolo_mslesseg/scripts/generar_consenso.py (:106-127)."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from . import device
from .utils import cargar_referencia_nifti, cargar_volumen, cargar_volumen_dispositivo, guardar_volumen


def _u8_dev(vol, nombre):
    a = np.asarray(vol)
    u = a.astype(np.uint8)
    if a.dtype != np.uint8 and not np.array_equal(u, a):
        raise ValueError(f"{nombre}: la ruta acelerada combina volúmenes con valores enteros 0..255 (máscaras)")
    return torch.from_numpy(np.ascontiguousarray(u.transpose(2, 1, 0) if u.ndim == 3 else u)).to(device())[None]


def combinar_volumenes(axial_vol, coronal_vol, sagital_vol, umbral=2):
    """((axial + coronal + sagital) >= umbral).astype(uint8), same shape as the inputs."""
    a = np.asarray(axial_vol)
    cons, _ = ops.consensus_eval(_u8_dev(axial_vol, "axial"), _u8_dev(coronal_vol, "coronal"), _u8_dev(sagital_vol, "sagital"),
                                 None, umbral)
    out = cons[0].cpu().numpy()
    return np.asfortranarray(out.transpose(2, 1, 0)) if a.ndim == 3 else out


def generar_consenso(axial_path, coronal_path, sagital_path, output_path, umbral=2):
    """scripts/generar_consenso.py:112-127 without a host round trip: the three .nii.gz are inflated on the GPU, voted
    there, and the consensus .nii.gz is deflated there (uint8 volume, affine of the axial file)."""
    vols = [cargar_volumen_dispositivo(p, torch.uint8)[None] for p in (axial_path, coronal_path, sagital_path)]
    affine = cargar_referencia_nifti(axial_path)[1]
    consenso, _ = ops.consensus_eval(vols[0], vols[1], vols[2], None, umbral)
    guardar_volumen(volumen=consenso[0], affine=affine, output_path=output_path)
