"""TEST INFRASTRUCTURE - not product code.

Imports the *real* reference (srozenblum/YOLO-MSLesSeg, mounted read-only at
/root/reference) inside the build container so that `oracle/make_golden.py` can
run the reference's own functions and freeze their outputs as golden vectors.

The reference imports `nibabel`, `matplotlib` and `ultralytics` at module top
level; none of them is installed here and none of them takes part in the
arithmetic of the hot path, so empty `sys.modules` stand-ins are enough
(SURVEY.md section 8c).  A tiny NIfTI-1 reader stands in for `nib.load` for the
two demo volumes.

/root/reference does not exist on the GPU box: nothing under tests/ marked
`gpu`, nothing in bench.py and nothing in the product may import this module.
"""
from __future__ import annotations

import gzip
import logging
import os
import struct
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MSLESSEG_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "yolo_mslesseg"))


def install_stubs() -> None:
    """Register stand-ins for the three absent third-party packages."""
    names = [
        "nibabel", "nibabel.filebasedimages",
        "ultralytics", "ultralytics.utils", "ultralytics.data", "ultralytics.data.converter",
        "matplotlib", "matplotlib.pyplot",
    ]
    for n in names:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["nibabel"].filebasedimages = sys.modules["nibabel.filebasedimages"]
    if not hasattr(sys.modules["nibabel.filebasedimages"], "ImageFileError"):
        sys.modules["nibabel.filebasedimages"].ImageFileError = type("ImageFileError", (Exception,), {})
    sys.modules["ultralytics"].YOLO = object
    sys.modules["ultralytics"].utils = sys.modules["ultralytics.utils"]
    sys.modules["ultralytics"].data = sys.modules["ultralytics.data"]
    sys.modules["ultralytics.data"].converter = sys.modules["ultralytics.data.converter"]
    sys.modules["ultralytics.utils"].LOGGER = logging.getLogger("ultralytics-stub")
    sys.modules["ultralytics.data.converter"].convert_segment_masks_to_yolo_seg = lambda **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


_REF = None


def load_reference():
    """Returns a namespace with the reference modules used by the hot path."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    cwd = os.getcwd()
    # configurar_logging opens a log file relative to CWD on import; keep it out of the repo.
    os.makedirs("/tmp/mslesseg_ref_cwd", exist_ok=True)
    os.chdir("/tmp/mslesseg_ref_cwd")
    try:
        import importlib
        ns = types.SimpleNamespace()
        ns.utils = importlib.import_module("yolo_mslesseg.utils.utils")
        ns.mejora = importlib.import_module("yolo_mslesseg.utils.mejora_imagen")
        ns.Paciente = importlib.import_module("yolo_mslesseg.utils.Paciente").Paciente
        ns.recon = importlib.import_module("yolo_mslesseg.scripts.reconstruir_volumen")
        ns.consenso = importlib.import_module("yolo_mslesseg.scripts.generar_consenso")
        ns.eval = importlib.import_module("yolo_mslesseg.scripts.eval")
        ns.promediar = importlib.import_module("yolo_mslesseg.scripts.promediar_folds")
        ns.extraer = importlib.import_module("yolo_mslesseg.scripts.extraer_dataset")
    finally:
        os.chdir(cwd)
    logging.getLogger().setLevel(logging.ERROR)
    _REF = ns
    return ns


def read_nifti(path: str) -> np.ndarray:
    """Minimal NIfTI-1 single-file reader -> float64 array (X, Y, Z), Fortran order,
    i.e. what `nib.load(path).get_fdata()` returns for the MSLesSeg files."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    hdr = raw[:348]
    endian = "<" if struct.unpack("<i", hdr[:4])[0] == 348 else ">"
    dim = struct.unpack(endian + "8h", hdr[40:56])
    datatype, bitpix = struct.unpack(endian + "2h", hdr[70:74])
    vox_offset = int(struct.unpack(endian + "f", hdr[108:112])[0])
    slope, inter = struct.unpack(endian + "2f", hdr[112:120])
    dt = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4"}[datatype]
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    n = int(np.prod(shape))
    arr = np.frombuffer(raw, dtype=np.dtype(endian + dt), count=n, offset=vox_offset)
    arr = arr.reshape(shape, order="F").astype(np.float64)
    if slope not in (0.0, 1.0) and np.isfinite(slope):
        arr = arr * slope + inter
    elif inter != 0.0 and np.isfinite(inter):
        arr = arr + inter
    return np.asfortranarray(arr)


def demo_volume(pid: str, kind: str) -> np.ndarray:
    """kind in {"FLAIR", "T1", "MASK"}; pid in {"P18", "P39"}."""
    p = os.path.join(REFERENCE_ROOT, "demo", "MSLesSeg-Dataset", "train", pid, "T1", f"{pid}_T1_{kind}.nii.gz")
    return read_nifti(p)
