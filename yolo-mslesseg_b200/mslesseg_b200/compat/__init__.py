"""Drop-in mirrors of the reference's Python call sites for the voxel path.

Same names, argument meaning and error behaviour as `yolo_mslesseg.utils.{mejora_imagen,utils,Paciente}`
and `yolo_mslesseg.scripts.{extraer_dataset,reconstruir_volumen,generar_consenso,eval,promediar_folds}`;
the bodies upload to the current CUDA device and call libmslesseg.so through mslesseg_b200.ops.
The reference's Config* classes, CLIs and logging are reused unchanged.
"""
import os

import torch


def device() -> torch.device:
    """CUDA device used by the shims (env MSLESSEG_GPU, default the current device).  No CPU fallback."""
    if not torch.cuda.is_available():
        raise RuntimeError("mslesseg_b200 needs a CUDA device: the voxel path has no CPU fallback")
    idx = os.environ.get("MSLESSEG_GPU")
    return torch.device("cuda", int(idx)) if idx is not None else torch.device("cuda", torch.cuda.current_device())
