"""Mirror of yolo_mslesseg/utils/Paciente.py: volume loader, plane slicer, lesion-slice selection and
enhancement dispatch - the natural batch point of the input side: one GPU call per
(patient, modality, plane) instead of one OpenCV round trip per slice.

The volumes live on the GPU as float32 [Z][Y][X] (the on-disk float32 is lossless); the
lesion flags of all three planes come from one pass over the mask (E0)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from .. import metrics as _M
from .. import codec as _codec
from .. import ops
from . import device
from .utils import ruta_existente


class Paciente:
    DATASET_DIR = Path("MSLesSeg-Dataset/train")
    MODALIDADES = ("T1", "T2", "FLAIR")
    MEJORAS = ("HE", "CLAHE", "GC", "LT")
    PLANOS = ("axial", "coronal", "sagital", "consenso")
    TIMEPOINTS = ("T1", "T2", "T3", "T4")

    def __init__(self, id, plano, timepoint="T1", modalidad=None, mejora=None, gt_mask=None):
        self._validar_argumentos(id, plano, timepoint, mejora, modalidad)
        self.id = id
        self.base_dir = self.DATASET_DIR / id
        self.plano = plano
        self.timepoint = timepoint
        self.sin_timepoints = not any((self.base_dir / tp).exists() for tp in self.TIMEPOINTS)
        self.mejora = mejora
        self._gt_mask = gt_mask
        self._volumenes = {}          # modality -> host array (X, Y, Z), as the reference caches them
        self._dev = {}                # modality / "gt" -> device tensor [1, Z, Y, X]
        self._flags = None
        self.modalidad = list(dict.fromkeys(modalidad))
        self.modalidad_str = "".join([m for m in self.MODALIDADES if m in set(self.modalidad)])

    # ---- constructor checks (utils/Paciente.py:89-117) ----
    def _validar_argumentos(self, id, plano, timepoint, mejora, modalidad):
        if not id.startswith("P"):
            raise ValueError(f"ID de paciente no válido: '{id}'. Debe seguir el formato 'P#' (por ejemplo: P1, P12, P53).")
        if plano not in self.PLANOS:
            raise ValueError(f"Plano {plano} no válido.")
        if timepoint not in self.TIMEPOINTS:
            raise ValueError(f"Timepoint {timepoint} no válido.")
        if mejora is not None and mejora not in self.MEJORAS:
            raise ValueError(f"Algoritmo de mejora '{mejora}' no válido. Opciones: {self.MEJORAS}")
        if not isinstance(modalidad, list) or not modalidad:
            raise TypeError("Modalidad debe ser una lista no vacía (por ejemplo, ['T1', 'T2'] o ['T1','T2','FLAIR'])")
        invalidas = [m for m in modalidad if m not in self.MODALIDADES]
        if invalidas:
            raise ValueError(f"Modalidades no reconocidas: {invalidas}")

    # ---- paths (utils/Paciente.py:139-157) ----
    def volumen_path(self, modalidad):
        if self.sin_timepoints:
            return self.base_dir / f"{self.id}_{modalidad}.nii.gz"
        return self.base_dir / self.timepoint / f"{self.id}_{self.timepoint}_{modalidad}.nii.gz"

    @property
    def gt_mask_path(self):
        if self.sin_timepoints:
            return self.base_dir / f"{self.id}_MASK.nii.gz"
        return self.base_dir / self.timepoint / f"{self.id}_{self.timepoint}_MASK.nii.gz"

    # ---- loading (utils/Paciente.py:159-193) ----
    def cargar_volumen(self, modalidad):
        if modalidad not in self._volumenes:
            vol_path = self.volumen_path(modalidad)
            if not ruta_existente(vol_path):
                raise FileNotFoundError(f"No se encontró el volumen {modalidad}.")
            # decoded on the GPU (the device copy is kept: it is what the enhancement kernels read)
            self._volumenes[modalidad] = np.asfortranarray(self._vol_dev(modalidad)[0].cpu().numpy().astype(np.float64).transpose(2, 1, 0))
        return self._volumenes[modalidad]

    @property
    def gt_mask(self):
        if self._gt_mask is None:
            if not ruta_existente(self.gt_mask_path):
                raise FileNotFoundError(f"No se encontró la máscara en {self.gt_mask_path}")
            g = _codec.nifti_load_device(self.gt_mask_path, device(), torch.float64)[0]
            self._gt_mask = np.asfortranarray(g.cpu().numpy().transpose(2, 1, 0))
        return self._gt_mask

    @property
    def num_cortes(self):
        mapping = {"axial": 2, "coronal": 1, "sagital": 0}
        if self.plano not in mapping:
            raise ValueError(f"Plano no reconocido: {self.plano}")
        if self._gt_mask is None and ruta_existente(self.gt_mask_path):
            return _codec.nifti_read_header(self.gt_mask_path)[0][mapping[self.plano]]
        return self.gt_mask.shape[mapping[self.plano]]

    # ---- device residency ----
    @staticmethod
    def _upload(vol_xyz, dtype):
        """(X, Y, Z) host array -> [1, Z, Y, X] device tensor (the same bytes when the array is Fortran-ordered)."""
        a = np.ascontiguousarray(np.asarray(vol_xyz).transpose(2, 1, 0).astype(dtype, copy=False))
        return torch.from_numpy(a).to(device())[None]

    def _vol_dev(self, modalidad):
        """float32 [1, Z, Y, X] on the device.  From the file: its bytes are uploaded and inflated on the GPU (no host
        decode); from an array the caller put into `_volumenes`: uploaded."""
        if modalidad not in self._dev:
            if modalidad in self._volumenes:
                self._dev[modalidad] = self._upload(self._volumenes[modalidad], np.float32)
            else:
                vol_path = self.volumen_path(modalidad)
                if not ruta_existente(vol_path):
                    raise FileNotFoundError(f"No se encontró el volumen {modalidad}.")
                self._dev[modalidad] = _codec.nifti_load_device(vol_path, device(), torch.float32)[0][None]
        return self._dev[modalidad]

    def _gt_dev(self):
        """uint8 [1, Z, Y, X]: (mask > 0), the predicate of indices_cortes_con_lesion (utils/Paciente.py:256)."""
        if "gt" not in self._dev:
            if self._gt_mask is not None:
                self._dev["gt"] = self._upload(np.asarray(self._gt_mask) > 0, np.uint8)
            else:
                if not ruta_existente(self.gt_mask_path):
                    raise FileNotFoundError(f"No se encontró la máscara en {self.gt_mask_path}")
                g = _codec.nifti_load_device(self.gt_mask_path, device(), torch.float32)[0]
                self._dev["gt"] = (g > 0).to(torch.uint8)[None]
        return self._dev["gt"]

    # ---- processing (utils/Paciente.py:195-246) ----
    def aplicar_mejora(self, imagen):
        if self.mejora is None:
            return imagen
        from . import mejora_imagen as MI
        cls = {"HE": MI.HE, "CLAHE": MI.CLAHE, "GC": MI.GC, "LT": MI.LT}.get(self.mejora)
        if cls is None:
            raise ValueError(f"Mejora no reconocida: {self.mejora}.")
        return cls().aplicar(imagen)

    def indice_plano(self, i):
        if self.plano == "consenso":
            raise ValueError("El plano 'consenso' no es un plano anatómico y no admite extracción de índices.")
        return {"axial": (slice(None), slice(None), i), "coronal": (slice(None), i, slice(None)),
                "sagital": (i, slice(None), slice(None))}[self.plano]

    def obtener_corte_imagen(self, i, modalidad):
        return self.aplicar_mejora(imagen=self.cargar_volumen(modalidad)[self.indice_plano(i)])

    def obtener_corte_mascara(self, i):
        return self.gt_mask[self.indice_plano(i)]

    # ---- lesion slices (utils/Paciente.py:252-275) ----
    def indices_cortes_con_lesion(self):
        if self.plano == "consenso":
            raise ValueError("El plano 'consenso' no es un plano anatómico y no admite extracción de índices.")
        if self._flags is None:
            self._flags = [f[0].cpu().numpy() for f in ops.lesion_slices(self._gt_dev())]
        k = {"axial": 0, "coronal": 1, "sagital": 2}[self.plano]
        return [int(i) for i in np.flatnonzero(self._flags[k])]

    def indices_a_usar(self, num_cortes=None):
        return _M.ventana_central(self.indices_cortes_con_lesion(), num_cortes)

    # ---- batched extraction (utils/Paciente.py:281-308) ----
    def cortes_con_lesion_gris(self, num_cortes=None, layout="G", en_dispositivo=False):
        """{modality: (indices, uint8 stack)}: verificar_grises(aplicar_mejora(slice)) for the selected slices in ONE
        kernel launch per modality - what guardar_cortes consumes.  layout "G" (slice orientation), "P" (PNG
        orientation) or "PNG_RGBA" (the pixels plt.imsave would write).  en_dispositivo: keep the stack on the GPU."""
        indices = self.indices_a_usar(num_cortes)
        res = {}
        for m in self.modalidad:
            st = ops.enhance_slices(self._vol_dev(m), self.mejora, self.plano, [0] * len(indices), indices, layout=layout)
            res[m] = (indices, st if en_dispositivo else st.cpu().numpy())
        return res

    def cortes_con_lesion_img(self, num_cortes=None):
        """{modality: [(index, slice), ...]} like the reference: 3-channel uint8 slices with an enhancement, the raw
        float64 slice without one."""
        indices = self.indices_a_usar(num_cortes)
        cortes = {}
        for m in self.modalidad:
            if self.mejora is None:
                vol = self.cargar_volumen(m)
                cortes[m] = [(i, vol[self.indice_plano(i)]) for i in indices]
            elif self.mejora == "CLAHE":
                ch = [ops.enhance_slices(self._vol_dev(m), "CLAHE", self.plano, [0] * len(indices), indices, lut_out=c)
                      for c in ("B", "G", "R")]
                st = torch.stack(ch, dim=-1).cpu().numpy()
                cortes[m] = [(i, st[n]) for n, i in enumerate(indices)]
            else:
                st = ops.enhance_slices(self._vol_dev(m), self.mejora, self.plano, [0] * len(indices), indices).cpu().numpy()
                cortes[m] = [(i, np.repeat(st[n][:, :, None], 3, axis=2)) for n, i in enumerate(indices)]
        return cortes

    def cortes_con_lesion_mask(self, num_cortes=None):
        indices = self.indices_a_usar(num_cortes)
        return [(i, self.obtener_corte_mascara(i)) for i in indices]

    def __repr__(self):
        return f"Paciente({self.id})"

    def __str__(self):
        return self.id
