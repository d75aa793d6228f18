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


def lineas_yolo(contornos, img_width, img_height, clase=0):
    """The label lines ultralytics' convert_segment_masks_to_yolo_seg writes for the contours of one mask: contours with
    at least 3 points, `class x1 y1 x2 y2 ...`, coordinates divided by the image size and rounded to 6 decimals."""
    lineas = []
    for contour in contornos:
        if len(contour) >= 3:
            item = [clase]
            for point in contour:
                item.append(round(point[0] / img_width, 6))
                item.append(round(point[1] / img_height, 6))
            lineas.append(" ".join(map(str, item)))
    return lineas


def anotar_mascaras(gt_masks_dir, labels_dir):
    """scripts/extraer_dataset.py:215-227 in one batch per directory: every mask PNG is decoded on the GPU, rewritten as a
    binary {0, 1} gray PNG (normalizar_mascaras / normalizar_mascara_binaria, utils/utils.py:387-393) and its external
    contours (cv2.findContours RETR_EXTERNAL / CHAIN_APPROX_SIMPLE of mask == 1, classes = 1 -> class 0) become the
    YOLO polygon lines of labels/<stem>.txt."""
    gt_masks_dir, labels_dir = Path(gt_masks_dir), Path(labels_dir)
    archivos = list(gt_masks_dir.glob("*.png"))
    if not archivos:
        raise FileNotFoundError(f"No se encontraron máscaras .png en {gt_masks_dir}")
    from .. import codec as _codec
    grupos = {}
    for ruta in archivos:
        try:
            datos = ruta.read_bytes()
            grupos.setdefault(_codec.png_parse(datos)[:3], []).append((ruta, datos))
        except Exception as e:
            raise OSError(f"Error al normalizar {ruta.name}: {e}")
    for (w, h, _), grupo in grupos.items():
        px = _codec.png_decode_first_channel([d for _, d in grupo], device())
        binaria = (px > 0).to(torch.uint8)
        data, off = ops.png_encode(binaria).to_host()
        contornos = ops.mask_contours(binaria, value=1)
        for n, (ruta, _) in enumerate(grupo):
            with open(ruta, "wb") as f:
                f.write(data[off[n]:off[n + 1]].tobytes())
            with open(labels_dir / f"{ruta.stem}.txt", "w") as f:
                for linea in lineas_yolo(contornos[n], w, h):
                    f.write(linea + "\n")
