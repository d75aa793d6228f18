"""TEST INFRASTRUCTURE - not product code.

Recipe that stages the UNMODIFIED reference package (`/root/reference/yolo_mslesseg`, pure Python, 29 files)
into the git-ignored `oracle/_ref/` so that the real reference - not a port - can be executed where
`/root/reference` does not exist (the GPU box receives `oracle/_ref/` with the rest of the working tree):

    python oracle/build_ref.py            # copies the package, writes oracle/_ref/MANIFEST.json (sha256 per file)

Nothing is copied into tracked paths; `oracle/_ref/` is listed in `.gitignore` (and not in `.gpurunignore`).
`__graft_entry__.build()` runs this whenever `/root/reference` is present.

`load()` imports the staged (or, in the build container, the original) package with stand-ins for the three
third-party packages that are absent from this image (`oracle/ref_stubs.py`): nibabel (a small NIfTI-1
reader / writer), matplotlib.pyplot.imsave (the restated E8 of oracle.py - UNPINNED, matplotlib is not
installable here) and ultralytics' mask -> YOLO polygon converter (restated on cv2.findContours).

Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline arms may use this module.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import logging
import os
import shutil
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
STAGED = HERE / "_ref"
SOURCE = Path(os.environ.get("MSLESSEG_REFERENCE_ROOT", "/root/reference"))
PKG = "yolo_mslesseg"


def source_available() -> bool:
    return (SOURCE / PKG / "utils" / "utils.py").is_file()


def staged_available() -> bool:
    return (STAGED / PKG / "utils" / "utils.py").is_file() and (STAGED / "MANIFEST.json").is_file()


def stage(force: bool = False) -> Path:
    """Copy the reference package verbatim (only *.py) into oracle/_ref/ and record a manifest."""
    if not source_available():
        raise RuntimeError(f"reference checkout not found at {SOURCE}")
    dst = STAGED / PKG
    if dst.exists():
        if not force and staged_available() and _manifest(dst) == json.loads((STAGED / "MANIFEST.json").read_text())["files"] \
                and _manifest(SOURCE / PKG) == _manifest(dst):
            return STAGED
        shutil.rmtree(dst)
    STAGED.mkdir(parents=True, exist_ok=True)
    shutil.copytree(SOURCE / PKG, dst, ignore=lambda d, names: [n for n in names
                                                              if not (n.endswith(".py") or (Path(d) / n).is_dir())])
    (STAGED / "MANIFEST.json").write_text(json.dumps({"source": str(SOURCE / PKG), "files": _manifest(dst)}, indent=1))
    return STAGED


def _manifest(root: Path) -> dict:
    return {str(p.relative_to(root)): hashlib.sha256(p.read_bytes()).hexdigest()
            for p in sorted(root.rglob("*.py"))}


def root() -> Path:
    """Directory to put on sys.path: the staged copy when it exists, else the original checkout."""
    if staged_available():
        return STAGED
    if source_available():
        return SOURCE
    raise RuntimeError("the reference is neither staged under oracle/_ref nor present at /root/reference; "
                       "run `python oracle/build_ref.py` where /root/reference exists")


def available() -> bool:
    return staged_available() or source_available()


_NS = None


def load(functional_stubs: bool = True):
    """Import the reference's hot-path modules; returns a namespace (utils, mejora, Paciente, Modelo, extraer, recon,
    consenso, eval, promediar, Config*).  The current directory during import is a scratch directory
    (configurar_logging opens pipeline.log relative to CWD at import time)."""
    global _NS
    if _NS is not None:
        return _NS
    from oracle import ref_stubs
    ref_stubs.install(functional=functional_stubs)
    r = str(root())
    if r not in sys.path:
        sys.path.insert(0, r)
    cwd = os.getcwd()
    os.makedirs("/tmp/mslesseg_ref_cwd", exist_ok=True)
    os.chdir("/tmp/mslesseg_ref_cwd")
    try:
        ns = types.SimpleNamespace()
        ns.root = r
        ns.utils = importlib.import_module(f"{PKG}.utils.utils")
        ns.mejora = importlib.import_module(f"{PKG}.utils.mejora_imagen")
        ns.paciente_mod = importlib.import_module(f"{PKG}.utils.Paciente")
        ns.Paciente = ns.paciente_mod.Paciente
        ns.Modelo = importlib.import_module(f"{PKG}.utils.Modelo").Modelo
        ns.extraer = importlib.import_module(f"{PKG}.scripts.extraer_dataset")
        ns.recon = importlib.import_module(f"{PKG}.scripts.reconstruir_volumen")
        ns.consenso = importlib.import_module(f"{PKG}.scripts.generar_consenso")
        ns.eval = importlib.import_module(f"{PKG}.scripts.eval")
        ns.promediar = importlib.import_module(f"{PKG}.scripts.promediar_folds")
        ns.predicciones = importlib.import_module(f"{PKG}.scripts.generar_predicciones")
    finally:
        os.chdir(cwd)
    logging.getLogger().setLevel(logging.ERROR)
    _NS = ns
    return ns


if __name__ == "__main__":
    out = stage(force="--force" in sys.argv)
    man = json.loads((out / "MANIFEST.json").read_text())
    print(f"staged {len(man['files'])} files of {man['source']} into {out}")
