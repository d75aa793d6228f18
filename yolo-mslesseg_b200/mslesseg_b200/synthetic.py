"""Seeded synthetic MSLesSeg-shaped cohorts.

There is no network for the real dataset, so tests and bench.py run on synthetic
182x218x182 volumes that reproduce the properties the hot path is sensitive to:
integer-valued float32 FLAIR intensities inside an ellipsoidal skull-stripped brain
(~26 % non-zero, so ~30 blank slices per plane hit the ptp == 0 branch of
`normalizar_a_uint8`, reference utils/utils.py:396-406), binary lesion masks made of a few
dozen blobs, and per-plane predicted masks that only exist on the lesion-selected slice
indices (reference utils/Paciente.py:261-275) as uint8 {0, 255} images in slice
orientation (reference scripts/generar_predicciones.py:136-140).

All arrays are returned in DEVICE layout: C-contiguous [Z][Y][X] (x fastest), which is
byte-identical to the Fortran-ordered (X, Y, Z) array `nib.load().get_fdata()` yields.
`as_xyz(a)` gives the reference-oriented view.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from functools import lru_cache

import numpy as np

SHAPE_XYZ = (182, 218, 182)
PLANOS = ("axial", "coronal", "sagital")
_CENTRE = (90.0, 109.0, 77.0)
_RADII = (73.0, 90.0, 77.0)


def as_xyz(a_zyx: np.ndarray) -> np.ndarray:
    """[Z][Y][X] C-order buffer -> (X, Y, Z) view as the reference indexes it."""
    return a_zyx.transpose(2, 1, 0)


def seed_for(config_id: int, patient_number: int) -> int:
    return 20260000 + 1000 * config_id + patient_number


@lru_cache(maxsize=4)
def brain_mask(shape_xyz=SHAPE_XYZ) -> np.ndarray:
    X, Y, Z = shape_xyz
    sx, sy, sz = X / 182.0, Y / 218.0, Z / 182.0
    z = (np.arange(Z, dtype=np.float32) - _CENTRE[2] * sz) / (_RADII[2] * sz)
    y = (np.arange(Y, dtype=np.float32) - _CENTRE[1] * sy) / (_RADII[1] * sy)
    x = (np.arange(X, dtype=np.float32) - _CENTRE[0] * sx) / (_RADII[0] * sx)
    r2 = z[:, None, None] ** 2 + y[None, :, None] ** 2 + x[None, None, :] ** 2
    m = r2 <= 1.0
    m.setflags(write=False)
    return m


def _slice_any(gt_zyx: np.ndarray, plano: str) -> np.ndarray:
    if plano == "axial":
        return gt_zyx.any(axis=(1, 2))
    if plano == "coronal":
        return gt_zyx.any(axis=(0, 2))
    return gt_zyx.any(axis=(0, 1))


def central_window(indices, num_cortes):
    """List arithmetic of reference utils/Paciente.py:261-275."""
    indices = list(indices)
    if num_cortes is None or len(indices) <= num_cortes:
        return indices
    start = max(0, len(indices) // 2 - num_cortes // 2)
    return indices[start:start + num_cortes]


def take_slices(vol_zyx: np.ndarray, plano: str, indices) -> np.ndarray:
    """Stack of slices in the reference's slice orientation (rows, cols):
    axial (X, Y), coronal (X, Z), sagital (Y, Z)."""
    v = as_xyz(vol_zyx)
    if plano == "axial":
        return np.ascontiguousarray(np.stack([v[:, :, i] for i in indices])) if len(indices) else \
            np.zeros((0,) + (v.shape[0], v.shape[1]), vol_zyx.dtype)
    if plano == "coronal":
        return np.ascontiguousarray(np.stack([v[:, i, :] for i in indices])) if len(indices) else \
            np.zeros((0,) + (v.shape[0], v.shape[2]), vol_zyx.dtype)
    return np.ascontiguousarray(np.stack([v[i, :, :] for i in indices])) if len(indices) else \
        np.zeros((0,) + (v.shape[1], v.shape[2]), vol_zyx.dtype)


@dataclass
class SyntheticPatient:
    id: str
    seed: int
    flair: np.ndarray                      # float32 [Z][Y][X]
    gt: np.ndarray                         # uint8 {0,1} [Z][Y][X]
    pred_indices: dict = field(default_factory=dict)   # plano -> list[int]
    pred_slices: dict = field(default_factory=dict)    # plano -> uint8 {0,255} [n][rows][cols]


def make_patient(patient_number: int, config_id: int = 1, num_cortes=None, shape_xyz=SHAPE_XYZ,
                 with_predictions: bool = True) -> SyntheticPatient:
    seed = seed_for(config_id, patient_number)
    rng = np.random.default_rng(seed)
    X, Y, Z = shape_xyz
    brain = brain_mask(shape_xyz)
    bidx = np.flatnonzero(brain)
    nb = bidx.size

    scale = rng.uniform(120.0, 700.0)
    vals = scale * (1.0 + 0.25 * rng.standard_normal(nb, dtype=np.float32))
    vals = np.round(np.clip(vals, 1.0, 2.4 * scale)).astype(np.float32)

    gt = np.zeros((Z, Y, X), dtype=np.uint8)
    K = int(rng.integers(3, 60))
    centres = bidx[rng.integers(0, nb, K)]
    radii = rng.integers(1, 7, K)
    cz, rem = np.divmod(centres, Y * X)
    cy, cx = np.divmod(rem, X)
    for z0, y0, x0, r in zip(cz, cy, cx, radii):
        r = int(r)
        zs, ze = max(0, z0 - r), min(Z, z0 + r + 1)
        ys, ye = max(0, y0 - r), min(Y, y0 + r + 1)
        xs, xe = max(0, x0 - r), min(X, x0 + r + 1)
        zz = (np.arange(zs, ze) - z0)[:, None, None]
        yy = (np.arange(ys, ye) - y0)[None, :, None]
        xx = (np.arange(xs, xe) - x0)[None, None, :]
        gt[zs:ze, ys:ye, xs:xe] |= ((zz * zz + yy * yy + xx * xx) <= r * r).astype(np.uint8)
    gt &= brain.astype(np.uint8)

    flair = np.zeros((Z, Y, X), dtype=np.float32)
    flair.ravel()[bidx] = vals
    les = gt.astype(bool)
    flair[les] = np.round(flair[les] * np.float32(1.4))

    pat = SyntheticPatient(id=f"P{patient_number}", seed=seed, flair=flair, gt=gt)
    if with_predictions:
        children = rng.spawn(3)
        gidx = np.flatnonzero(gt)
        for plano, crng in zip(PLANOS, children):
            pred = np.zeros((Z, Y, X), dtype=np.uint8)
            keep = crng.random(gidx.size, dtype=np.float32) > 0.25
            pred.ravel()[gidx[keep]] = 255
            add = crng.random(nb, dtype=np.float32) < 0.001
            pred.ravel()[bidx[add]] = 255
            idx = central_window(np.flatnonzero(_slice_any(gt, plano)).tolist(), num_cortes)
            pat.pred_indices[plano] = [int(i) for i in idx]
            pat.pred_slices[plano] = take_slices(pred, plano, idx)
    return pat


def make_cohort(patient_numbers, config_id: int = 1, num_cortes=None, shape_xyz=SHAPE_XYZ,
                with_predictions: bool = True):
    return [make_patient(n, config_id, num_cortes, shape_xyz, with_predictions) for n in patient_numbers]
