"""Mirror of the hot-path helpers of yolo_mslesseg/utils/utils.py (image, metric and fold parts)."""
from __future__ import annotations

import argparse
import json
import logging
import os
import struct
import zlib
from pathlib import Path

import numpy as np
import torch

from .. import metrics as _M
from .. import codec as _codec
from .. import ops
from . import device

logger = logging.getLogger(__name__)


# ---- NIfTI / JSON (utils/utils.py:153-181, 259-269) ------------------------------------------------
def ruta_existente(path):
    return Path(path).exists()


def cargar_volumen_dispositivo(vol_path, dtype=torch.float32):
    """The volume of a .nii / .nii.gz as a device tensor [Z][Y][X] (float32, uint8 or float64): the FILE bytes are
    uploaded and inflated / converted on the GPU (mslesseg_b200.codec).  uint8 / float32 raise ValueError when the file
    holds values they cannot represent exactly."""
    return _codec.nifti_load_device(vol_path, device(), dtype)[0]


def cargar_volumen(vol_path):
    """nib.load(vol_path).get_fdata(): float64 (X, Y, Z), Fortran order.  Decoded on the GPU, then copied back."""
    try:
        v = cargar_volumen_dispositivo(vol_path, torch.float64)
        return np.asfortranarray(v.cpu().numpy().transpose(2, 1, 0))
    except Exception as e:
        logger.error(f"❌ Error al cargar el volumen desde {vol_path}: {e}")
        raise


def cargar_referencia_nifti(referencia_path):
    if not ruta_existente(referencia_path):
        raise FileNotFoundError(f"Archivo no encontrado: {referencia_path}")
    try:
        return _codec.nifti_read_header(referencia_path)
    except (_codec.CodecError, zlib.error, OSError, struct.error) as e:
        raise ValueError(f"Archivo no válido: {referencia_path}") from e


_TORCH_DT = {np.dtype(np.float32): torch.float32, np.dtype(np.uint8): torch.uint8, np.dtype(np.float64): torch.float64,
             np.dtype(np.int16): torch.int16, np.dtype(np.int32): torch.int32, np.dtype(np.int8): torch.int8}


def guardar_volumen(volumen, affine, output_path):
    """nib.save(nib.Nifti1Image(volumen, affine), output_path): the .nii.gz is deflated on the GPU (64 KB gzip members,
    readable by any gzip / nibabel); `volumen` may be a NumPy array (X, Y, Z) or a device tensor [Z][Y][X]."""
    try:
        if isinstance(volumen, torch.Tensor):
            dev = volumen
        else:
            a = np.asarray(volumen)
            if a.dtype == np.bool_:
                a = a.astype(np.uint8)
            if a.dtype not in _TORCH_DT:
                raise ValueError(f"no se puede guardar un volumen de tipo {a.dtype} en NIfTI-1")
            if a.ndim != 3:
                raise ValueError(f"se esperaba un volumen 3D, forma {a.shape}")
            dev = torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).to(device())
        _codec.nifti_save_device(dev, affine, output_path)
    except Exception as e:
        logger.error(f"❌ Error al guardar el volumen en {output_path}: {e}")
        raise


def reconstruccion_valida(pred_vol_path, gt_vol_path):
    """Shape equality of the two volumes (utils/utils.py:183-194); only the headers are read."""
    ps, gs = _codec.nifti_read_header(pred_vol_path)[0], _codec.nifti_read_header(gt_vol_path)[0]
    if ps != gs:
        logger.warning(f"⚠️ Dimensiones distintas: {ps} vs {gs}")
        return False
    return True


def escribir_json(dic, json_path):
    with open(json_path, "w") as f:
        json.dump(dic, f)


def leer_json(json_path):
    if os.path.exists(json_path):
        with open(json_path, "r") as f:
            return json.load(f)
    raise FileNotFoundError(f"Archivo no encontrado: {json_path}")


# ---- patients / folds (utils/utils.py:286-316, 343-358, 435-447) ------------------------------------
def archivo_ignorable(nombre):
    return nombre.startswith(".") or nombre.startswith("~") or nombre.lower().endswith(".tmp")


def listar_pacientes(input_dir):
    pacientes = [d.name for d in Path(input_dir).iterdir() if not archivo_ignorable(d.name)]
    if not pacientes:
        raise FileNotFoundError(f"No se encontraron pacientes en {input_dir}.")
    return sorted(pacientes, key=lambda p: int(p[1:]) if p[1:].isdigit() else 1_000_000)


calcular_fold = _M.calcular_fold


def int_o_percentil(valor):
    try:
        return int(valor)
    except ValueError:
        if isinstance(valor, str) and valor.upper().startswith("P") and valor[1:].isdigit():
            return valor.upper()
        raise argparse.ArgumentTypeError(
            "El valor debe ser un entero o un string de formato 'PX' (ejemplo: P10 para percentil 10).")


def evaluar_resultados(resultados):
    if not resultados:
        return None
    if all(r is None for r in resultados):
        return None
    if all(r is True for r in resultados):
        return True
    return "parcial"


# ---- image helpers (utils/utils.py:396-427) ---------------------------------------------------------
def normalizar_a_uint8(imagen):
    """Per-image float32 min / ptp normalisation to uint8 (E1); uint8 input is returned unchanged."""
    imagen = np.asarray(imagen)
    if imagen.dtype == np.uint8:
        return imagen
    if imagen.ndim != 2:
        raise ValueError("normalizar_a_uint8 acelerado espera una imagen 2D")
    d = torch.from_numpy(np.ascontiguousarray(imagen.astype(np.float32))).to(device())[None]
    return ops.enhance_images(d, None, layout="G")[0].cpu().numpy()


def convertir_a_bgr(imagen):
    u = normalizar_a_uint8(imagen)
    if u.ndim == 2:
        return np.repeat(u[:, :, None], 3, axis=2)
    return np.ascontiguousarray(u[:, :, ::-1])


def verificar_grises(imagen):
    """cv2.COLOR_BGR2GRAY for 3-channel images (OpenCV's 15-bit fixed point: (9798 R + 19235 G + 3735 B + 2^14) >> 15),
    identity for 2-D images.  The enhancement shims already return images whose gray value is the reference's."""
    imagen = np.asarray(imagen)
    if imagen.ndim == 3 and imagen.shape[2] == 3:
        if imagen.dtype != np.uint8:
            raise ValueError("verificar_grises: se espera una imagen uint8")
        return ops.bgr_to_gray(torch.from_numpy(np.ascontiguousarray(imagen)).to(device())).cpu().numpy()
    return imagen


def normalizar_mascara_binaria(mask_path):
    """utils/utils.py:387-393: cv2.imread(GRAYSCALE) -> (mask > 0) -> cv2.imwrite, the file rewritten in place as an 8-bit
    gray PNG with values {0, 1}.  Decode, threshold and encode run on the GPU (the gray value of the gray-colormapped RGBA
    files guardar_cortes writes is their first channel)."""
    from pathlib import Path as _P
    px = _codec.png_decode_first_channel([_P(mask_path).read_bytes()], device())
    data, off = ops.png_encode((px > 0).to(torch.uint8)).to_host()
    with open(mask_path, "wb") as f:
        f.write(data[:int(off[1])].tobytes())


# ---- metrics (utils/utils.py:455-495) ---------------------------------------------------------------
def _as_mask_u8(a, nombre):
    a = np.asarray(a)
    if a.dtype != np.uint8:
        u = a.astype(np.uint8)
        if not np.array_equal(u, a):
            raise ValueError(f"{nombre}: la ruta acelerada evalúa máscaras con valores enteros 0..255 (binarias en la práctica)")
        a = u
    return torch.from_numpy(np.ascontiguousarray(a).reshape(1, -1)).to(device())


def conteos(y_true, y_pred):
    """(tp, fp, fn, tn) with the reference's ==1 / ==0 predicates, counted on the GPU."""
    y_true, y_pred = np.asarray(y_true), np.asarray(y_pred)
    if y_true.shape != y_pred.shape:
        raise ValueError(f"Dimensiones distintas: {y_true.shape} vs {y_pred.shape}")
    c = ops.confusion_counts(_as_mask_u8(y_true, "y_true"), _as_mask_u8(y_pred, "y_pred"))[0].cpu().numpy()
    if int(c.sum()) != y_true.size:
        raise ValueError("La ruta acelerada calcula las métricas de máscaras binarias {0, 1}; "
                         "se encontraron otros valores.")
    return tuple(int(x) for x in c)


def metricas(y_true, y_pred) -> dict:
    return _M.metricas_desde_conteos(*conteos(y_true, y_pred))


def DSC(y_true, y_pred):
    return metricas(y_true, y_pred)["DSC"]


def precision(y_true, y_pred):
    return metricas(y_true, y_pred)["Precision"]


def recall(y_true, y_pred):
    return metricas(y_true, y_pred)["Recall"]


def AUC(y_true, y_pred):
    try:
        return metricas(y_true, y_pred)["AUC"]
    except Exception as e:
        logger.warning(f"⚠️ No se pudo calcular AUC: {e}")
        return np.nan
