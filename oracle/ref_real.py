"""CPU BASELINE through the REAL reference.  TEST / BENCH INFRASTRUCTURE - NOT PRODUCT CODE.

Same interface as oracle/ref_path.py (the port), but every call lands in the reference's own, unmodified functions,
imported from the copy staged by oracle/build_ref.py (or from /root/reference in the build container):

    enhance_plane    Paciente.obtener_corte_imagen -> aplicar_mejora -> <HE|CLAHE|GC|LT>().aplicar   utils/Paciente.py:195-222
                     + verificar_grises (scripts/extraer_dataset.py:190) + the .T / origin="lower" orientation (:192)
    reconstruir      the (> 0).astype(float32) of cargar_y_preprocesar_imagen (scripts/reconstruir_volumen.py:146-148)
                     + insertar_corte (:179-186) into np.zeros(shape, float32) (:202)
    combinar_volumenes            scripts/generar_consenso.py:106
    generar_diccionario_metricas  scripts/eval.py:115 (DSC / AUC / precision / recall of utils/utils.py:455-495)

bench.py reports this arm as cpu_baseline.kind = "reference".
"""
from __future__ import annotations

import warnings

import numpy as np

from . import build_ref
from . import oracle as O

try:
    import cv2
    cv2.setNumThreads(1)          # one worker process per core in bench.py; no nested thread pools
except Exception:  # pragma: no cover
    cv2 = None

BACKEND = "yolo_mslesseg (unmodified reference, staged copy)"


def available() -> bool:
    return build_ref.available() and cv2 is not None


def _ns():
    return build_ref.load()


def enhance_plane(vol_xyz, plano, mejora, indices=None):
    ns = _ns()
    pac = ns.Paciente(id="P1", plano=plano, modalidad=["FLAIR"], mejora=mejora, gt_mask=np.zeros((1, 1, 1)))
    pac._volumenes["FLAIR"] = vol_xyz                      # the cache nib.load(...).get_fdata() fills (utils/Paciente.py:164-168)
    n = vol_xyz.shape[O.plane_axis(plano)]
    out = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                    # LT on blank slices divides by log(1) (RuntimeWarning in the reference too)
        for i in (range(n) if indices is None else indices):
            g = ns.utils.verificar_grises(pac.obtener_corte_imagen(i, "FLAIR"))
            out.append(np.ascontiguousarray(g.T[::-1]))
    return out


def reconstruir(slices, indices, shape_xyz, plano):
    ns = _ns()
    vol = np.zeros(tuple(int(d) for d in shape_xyz), dtype=np.float32)
    for q, i in sorted(zip(slices, indices), key=lambda t: t[1]):
        img = np.asarray(q)
        if img.max() > 1:
            img = (img > 0).astype(np.float32)
        ns.recon.validar_corte(indice=int(i), img_array=img, shape_original=vol.shape, plano=plano)
        ns.recon.insertar_corte(vol, img, int(i), plano)
    return vol


def combinar_volumenes(ax, co, sa, umbral=2):
    return _ns().consenso.combinar_volumenes(ax, co, sa, umbral)


def generar_diccionario_metricas(gt_vol, pred_vol):
    return _ns().eval.generar_diccionario_metricas(gt_vol, pred_vol)
