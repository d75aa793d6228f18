"""The import swap of INTEGRATION.md section 1 as one call.

    import mslesseg_b200.compat.install as fast
    fast.install()            # before or after the reference's scripts were imported
    ...                       # run yolo_mslesseg.ejecutar_pipeline / the stage scripts unchanged
    fast.uninstall()          # (tests) put the reference's own functions back

`install()` imports the reference's modules and rebinds, in every loaded `yolo_mslesseg.*` module, each global
that IS one of the reference's hot-path functions / classes to its GPU mirror in `mslesseg_b200.compat`
(identity match, so `from yolo_mslesseg.utils.utils import cargar_volumen` copies inside the stage scripts are
caught too).  Nothing else is touched: Config*, CLIs, logging, skip-if-exists and the per-patient try / except
loops are the reference's own code.  There is no CPU fallback: `install()` raises if the
CUDA library or a GPU is missing.
"""
from __future__ import annotations

import importlib
import sys

_PKG = "yolo_mslesseg"
_saved: list = []          # (module, name, original object)

# reference module -> {name: (compat module, compat name)}
_SWAPS = {
    "utils.utils": {n: ("utils", n) for n in (
        "cargar_volumen", "cargar_referencia_nifti", "guardar_volumen", "reconstruccion_valida",
        "normalizar_a_uint8", "convertir_a_bgr", "verificar_grises", "normalizar_mascara_binaria",
        "DSC", "precision", "recall", "AUC")},
    "utils.mejora_imagen": {n: ("mejora_imagen", n) for n in ("HE", "CLAHE", "GC", "LT")},
    "utils.Paciente": {"Paciente": ("Paciente", "Paciente")},
    "scripts.extraer_dataset": {n: ("extraer_dataset", n) for n in (
        "calcular_num_cortes_percentil", "resolver_num_cortes", "guardar_cortes", "anotar_mascaras")},
    "scripts.reconstruir_volumen": {n: ("reconstruir_volumen", n) for n in (
        "extraer_indices_png", "validar_corte", "reconstruir_volumen")},
    "scripts.generar_predicciones": {n: ("generar_predicciones", n) for n in (
        "combinar_predicciones", "normalizar_prediccion")},
    "scripts.generar_consenso": {n: ("generar_consenso", n) for n in ("combinar_volumenes", "generar_consenso")},
    "scripts.eval": {n: ("eval", n) for n in ("generar_diccionario_metricas", "calcular_metricas", "calcular_promedio")},
    "scripts.promediar_folds": {n: ("eval", n) for n in ("agregar_metricas_fold", "calcular_resumen_experimento")},
}


def install(paquete: str = _PKG) -> int:
    """Rebinds the reference's hot-path names to their GPU mirrors; returns the number of bindings replaced."""
    from .. import _lib
    from . import device
    _lib.load()                 # raises when libmslesseg.so is missing
    device()                    # raises without a CUDA device
    if _saved:
        return len(_saved)
    replacement = {}            # id(original) -> (original, new)
    for ref_mod, names in _SWAPS.items():
        try:
            mod = importlib.import_module(f"{paquete}.{ref_mod}")
        except ImportError:
            continue             # e.g. generar_predicciones needs ultralytics
        for name, (cmod, cname) in names.items():
            orig = getattr(mod, name, None)
            cm = importlib.import_module(f"{__package__}.{cmod}")
            new = getattr(cm, cname, None)
            if orig is None or new is None or orig is new:
                continue
            replacement[id(orig)] = (orig, new)
    for mname, mod in list(sys.modules.items()):
        if mod is None or not (mname == paquete or mname.startswith(paquete + ".")):
            continue
        for name, val in list(vars(mod).items()):
            hit = replacement.get(id(val))
            if hit is not None and hit[0] is val:
                _saved.append((mod, name, val))
                setattr(mod, name, hit[1])
    return len(_saved)


def uninstall() -> None:
    while _saved:
        mod, name, orig = _saved.pop()
        setattr(mod, name, orig)


def installed() -> bool:
    return bool(_saved)
