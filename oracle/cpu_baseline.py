"""CPU baseline runner.  BENCH INFRASTRUCTURE - NOT PRODUCT CODE.

Times the reference's CPU path - the REAL reference functions through oracle/ref_real.py when the package is staged
under oracle/_ref (oracle/build_ref.py; kind "reference"), else the port oracle/ref_path.py (the same cv2 / NumPy /
scikit-learn calls; kind "port") - on the host cores of the box bench.py runs on, over a bounded sample of the bench
workload: whole synthetic patients, every enhancement x plane over ALL slices, recon of the three
planes, consensus, metrics of the three planes + consensus.

Parallelism: the reference itself is a single sequential process; to use "all the host threads it
can" the per-patient chain is cut into 16 independent tasks (12 enhance (mejora, plano) tasks, 3
recon+eval tasks, 1 recon x3 + consensus + eval task) spread over a fork()ed multiprocessing pool,
one worker per core.  The consensus task recomputes the three reconstructions (~4 % extra work).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for _p in (str(ROOT), str(ROOT / "yolo-mslesseg_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from oracle import oracle as O            # noqa: E402
from oracle import ref_real as _RR         # noqa: E402
if _RR.available():
    R, KIND = _RR, "reference"
else:                                     # pragma: no cover - only without oracle/_ref
    from oracle import ref_path as R      # noqa: E402
    KIND = "port"
from mslesseg_b200 import synthetic as S  # noqa: E402

_PATIENTS = []          # filled in the parent before the pool forks


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _task(args):
    p, kind, a, b = args
    pat = _PATIENTS[p]
    if kind == "enh":
        vol = np.asfortranarray(pat.flair.transpose(2, 1, 0).astype(np.float64))
        return len(R.enhance_plane(vol, b, a))
    gt = np.asfortranarray(pat.gt.transpose(2, 1, 0).astype(np.float64))
    if kind == "rec_eval":
        vol = R.reconstruir(pat.pred_slices[a], pat.pred_indices[a], S.SHAPE_XYZ, a).astype(np.float64)
        return R.generar_diccionario_metricas(gt, vol)["DSC"]
    vols = [R.reconstruir(pat.pred_slices[pl], pat.pred_indices[pl], S.SHAPE_XYZ, pl).astype(np.float64) for pl in O.PLANOS]
    cons = R.combinar_volumenes(vols[0], vols[1], vols[2], 2).astype(np.float64)
    return R.generar_diccionario_metricas(gt, cons)["DSC"]


def _tasks(n_patients: int):
    t = []
    for p in range(n_patients):               # most expensive first
        t.append((p, "cons_eval", None, None))
        t += [(p, "rec_eval", pl, None) for pl in O.PLANOS]
    for p in range(n_patients):
        t += [(p, "enh", "CLAHE", pl) for pl in O.PLANOS]
    for p in range(n_patients):
        t += [(p, "enh", m, pl) for m in ("HE", "LT", "GC") for pl in O.PLANOS]
    return t


class CpuBaseline:
    """with CpuBaseline(n_patients) as cb: seconds = cb.step()"""

    def __init__(self, n_patients: int, cores: int | None = None, config_id: int = 4, num_cortes: int = 40):
        self.cores = cores or host_cores()
        self.n_patients = n_patients
        global _PATIENTS
        _PATIENTS = [S.make_patient(n + 1, config_id=config_id, num_cortes=num_cortes) for n in range(n_patients)]
        if KIND == "reference":
            _RR._ns()                      # import the reference once in the parent; the forked workers inherit it
        self.pool = mp.get_context("fork").Pool(self.cores)
        self.voxels_per_step = n_patients * int(np.prod(S.SHAPE_XYZ))

    def step(self) -> float:
        t0 = time.perf_counter()
        for _ in self.pool.imap_unordered(_task, _tasks(self.n_patients), chunksize=1):
            pass
        return time.perf_counter() - t0

    def describe(self) -> str:
        return (f"{self.n_patients} synthetic patients x (4 enhancements x 3 planes over all 582 slices + recon x3 + "
                f"consensus + 4 metric sets), {self.cores} worker processes, backend {R.BACKEND}, CPU {cpu_model()}")

    def close(self):
        self.pool.close()
        self.pool.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def default_sample_patients(cores: int) -> int:
    return max(2, cores // 4)
