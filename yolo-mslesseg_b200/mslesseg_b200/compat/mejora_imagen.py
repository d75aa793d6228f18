"""Mirror of yolo_mslesseg/utils/mejora_imagen.py: Algoritmo, HE, CLAHE, GC, LT with `.aplicar(imagen)`.

`aplicar` takes one 2-D image (float: normalised per image like normalizar_a_uint8; uint8: used as is) and
returns the (rows, cols, 3) uint8 array the reference returns ("RGB" for HE, "BGR" for the others): the three
channels are equal for HE / GC / LT; CLAHE's differ by at most one level, exactly like cv2's LAB2BGR of
(L', 128, 128).  `aplicar_gris` / `aplicar_lote` return what every caller in the reference actually uses,
verificar_grises(aplicar(...)) - the single gray channel - for one image or a batch.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from . import device


def _to_device_batch(imagenes) -> torch.Tensor:
    arr = np.asarray(imagenes)
    if arr.ndim == 2:
        arr = arr[None]
    if arr.ndim != 3:
        raise ValueError("Se esperaba una imagen 2D (o un lote [n, filas, columnas]); las imágenes RGB de entrada "
                         "no forman parte de la ruta acelerada.")
    if arr.dtype != np.uint8:
        arr = arr.astype(np.float32)          # normalizar_a_uint8 casts to float32 first (utils/utils.py:400)
    return torch.from_numpy(np.ascontiguousarray(arr)).to(device())


class Algoritmo:
    """Clase base: interfaz común de las técnicas de mejora (utils/mejora_imagen.py:22-35)."""
    nombre = None

    def aplicar(self, imagen):
        raise NotImplementedError("El método aplicar debe ser implementado por la clase hija.")

    def aplicar_lote(self, imagenes) -> np.ndarray:
        """[n, filas, columnas] uint8: verificar_grises(aplicar(imagen)) for every image of the batch."""
        return ops.enhance_images(_to_device_batch(imagenes), self.nombre, layout="G").cpu().numpy()

    def aplicar_gris(self, imagen) -> np.ndarray:
        return self.aplicar_lote(imagen)[0]

    def __repr__(self):
        return self.nombre


class _TresCanalesIguales(Algoritmo):
    def aplicar(self, imagen):
        g = self.aplicar_gris(imagen)
        return np.repeat(g[:, :, None], 3, axis=2)


class HE(_TresCanalesIguales):
    """Ecualización de histograma (utils/mejora_imagen.py:43-70)."""
    nombre = "HE"


class GC(_TresCanalesIguales):
    """Corrección gamma, gamma = 2.0 (utils/mejora_imagen.py:121-154)."""
    nombre = "GC"

    def __init__(self, gamma=2.0):
        if gamma != 2.0:
            raise ValueError("La ruta acelerada fija gamma = 2.0 (valor por defecto de la referencia).")
        self.gamma = gamma


class LT(_TresCanalesIguales):
    """Transformación logarítmica (utils/mejora_imagen.py:157-187)."""
    nombre = "LT"


class CLAHE(Algoritmo):
    """CLAHE sobre el canal L, clip_limit = 2.0, tile_grid_size = (8, 8) (utils/mejora_imagen.py:73-118)."""
    nombre = "CLAHE"

    def __init__(self, clip_limit=2.0, tile_grid_size=(8, 8)):
        if clip_limit != 2.0 or tuple(tile_grid_size) != (8, 8):
            raise ValueError("La ruta acelerada fija clip_limit = 2.0 y tile_grid_size = (8, 8) (valores por defecto).")
        self.clip_limit = clip_limit
        self.tile_grid_size = tile_grid_size

    def aplicar(self, imagen):
        d = _to_device_batch(imagen)
        canales = [ops.enhance_images(d, "CLAHE", layout="G", lut_out=c)[0] for c in ("B", "G", "R")]
        return torch.stack(canales, dim=-1).cpu().numpy()
