"""Mirror of the numeric part of yolo_mslesseg/scripts/extraer_dataset.py (:110-197)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from .. import metrics as _M
from .. import ops
from . import device
from .Paciente import Paciente
from .utils import listar_pacientes


def calcular_num_cortes_percentil(input_dir, plano, modalidad, percentil=50):
    """int(np.percentile(lesion-slice counts of every patient, percentil)); each patient's counts for the three
    planes come from one GPU pass over its mask."""
    conteos = [len(Paciente(id=pid, plano=plano, modalidad=modalidad).indices_a_usar()) for pid in listar_pacientes(input_dir)]
    if not conteos:
        raise ValueError(f"No se encontraron cortes con lesión válidos para calcular el percentil en {input_dir}.")
    return _M.num_cortes_percentil(conteos, percentil)


def resolver_num_cortes(num_cortes, input_dir, plano, modalidad):
    if isinstance(num_cortes, int) or num_cortes is None:
        return num_cortes, None
    if isinstance(num_cortes, str) and num_cortes.startswith("P"):
        percentil = int(num_cortes[1:])
        return calcular_num_cortes_percentil(input_dir=input_dir, plano=plano, modalidad=modalidad, percentil=percentil), percentil
    raise ValueError(f"Formato de num_cortes no válido: {num_cortes}.")


def _escribir_pngs(rgba_dev, rutas):
    """One PNG file per image of the device stack [n, H, W, 4]: the files are deflated and assembled on the GPU
    (ops.png_encode: fixed-Huffman deflate with run matches, Adler-32 / CRC-32 in the kernels), the host only writes
    the bytes."""
    data, off = ops.png_encode(rgba_dev).to_host()
    for n, ruta in enumerate(rutas):
        with open(ruta, "wb") as f:
            f.write(data[off[n]:off[n + 1]].tobytes())


def guardar_cortes(paciente, images_dir, gt_masks_dir, num_cortes):
    """Writes the image and mask PNGs of guardar_cortes (:174-197).  The RGBA pixels are what
    plt.imsave(path, corte.T, cmap="gray", origin="lower") produces (orientation, second normalisation and gray
    colormap computed on the GPU, layout PNG_RGBA), and so are the PNG files themselves."""
    images_dir, gt_masks_dir = Path(images_dir), Path(gt_masks_dir)
    indices = paciente.indices_a_usar(num_cortes)
    if not indices:
        raise ValueError(f"No se encontraron cortes válidos para el paciente {paciente.id}.")
    for modalidad, (idx, rgba) in paciente.cortes_con_lesion_gris(num_cortes, layout="PNG_RGBA", en_dispositivo=True).items():
        _escribir_pngs(rgba, [images_dir / f"{paciente.id}_{modalidad}_{i}.png" for i in idx])
    masks = ops.enhance_slices(paciente._gt_dev(), None, paciente.plano, [0] * len(indices), indices, layout="PNG_RGBA")
    _escribir_pngs(masks, [gt_masks_dir / f"{paciente.id}_{i}.png" for i in indices])
