"""Multi-GPU plumbing of the voxel path.

The unit of work is a patient.  Patients are sharded across ranks (one process per GPU) with no
data-path collective; the only exchange is one all-reduce (SUM) of the int64 confusion-count table
[n_patients, 4 planes, (tp, fp, fn, tn)] - every rank fills the rows of its own patients, the others stay
zero - plus, when num_cortes is a percentile, an int32 all-reduce of the lesion-slice counts.  The reference
never pools counts across patients (fold / global figures are means of per-patient ROUNDED metrics,
scripts/eval.py:154-155, scripts/promediar_folds.py:131-132), so metrics are computed per row of the reduced
table.  Backend: NCCL over NVLink for CUDA tensors, gloo for the CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import metrics as M

PLANOS4 = ("axial", "coronal", "sagital", "consenso")


def patient_number(pid: str) -> int:
    return int(pid[1:]) if pid[1:].isdigit() else 1_000_000


def sort_patients(ids: Sequence[str]) -> List[str]:
    """Same order as listar_pacientes (reference utils/utils.py:286-296)."""
    return sorted(ids, key=patient_number)


def shard_patients(ids: Sequence[str], world_size: int, rank: int, k_folds: Optional[int] = None,
                   n_ids: int = 53) -> List[str]:
    """Patients of `rank`.  Fold-major when there are at least as many folds as ranks (a rank owns whole
    folds, so per-fold statistics need no further exchange), otherwise round-robin over the sorted list."""
    ids = sort_patients(ids)
    if k_folds is not None and k_folds >= world_size:
        return [p for p in ids if (M.calcular_fold(p, k_folds, n_ids) - 1) % world_size == rank]
    return ids[rank::world_size]


def _all_reduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_reduce_count_table(ids_all: Sequence[str], ids_local: Sequence[str], counts_local: torch.Tensor,
                           group=None) -> torch.Tensor:
    """counts_local: int64 [len(ids_local), 4, 4] on any device -> int64 [len(ids_all), 4, 4] holding every
    patient's counts on every rank (rows ordered like sort_patients(ids_all))."""
    order = {p: i for i, p in enumerate(sort_patients(ids_all))}
    table = torch.zeros((len(order), 4, 4), dtype=torch.int64, device=counts_local.device)
    if len(ids_local):
        rows = torch.as_tensor([order[p] for p in ids_local], dtype=torch.long, device=counts_local.device)
        table[rows] = counts_local.to(torch.int64)
    return _all_reduce_sum(table, group)


def all_reduce_lesion_counts(ids_all: Sequence[str], ids_local: Sequence[str], n_lesion_local: Sequence[int],
                             device="cpu", group=None) -> List[int]:
    """Number of lesion slices of every patient (for num_cortes = "Pxx": scripts/extraer_dataset.py:110-135)."""
    order = {p: i for i, p in enumerate(sort_patients(ids_all))}
    t = torch.zeros(len(order), dtype=torch.int32, device=device)
    for p, n in zip(ids_local, n_lesion_local):
        t[order[p]] = int(n)
    return [int(x) for x in _all_reduce_sum(t, group).cpu()]


def metrics_from_table(ids_all: Sequence[str], table: torch.Tensor) -> Dict[str, Dict[str, dict]]:
    """{patient: {plano: {"DSC","AUC","Precision","Recall"}}} from the reduced table (host float64)."""
    t = table.cpu().numpy()
    out = {}
    for i, p in enumerate(sort_patients(ids_all)):
        out[p] = {pl: M.metricas_desde_conteos(*t[i, k]) for k, pl in enumerate(PLANOS4)}
    return out


def fold_and_global_stats(per_patient: Dict[str, dict], k_folds: int, n_ids: int = 53) -> dict:
    """calcular_promedio per fold (scripts/eval.py:144-160, population std) and calcular_resumen_experimento
    over the fold means (scripts/promediar_folds.py:126-134, sample std), for one plane's metric dicts."""
    folds: Dict[int, Dict[str, list]] = {}
    for pid in sort_patients(list(per_patient)):
        f = M.calcular_fold(pid, k_folds, n_ids)
        for k, v in per_patient[pid].items():
            folds.setdefault(f, {}).setdefault(k, []).append(v)
    fold_stats = {f: M.calcular_promedio(d) for f, d in sorted(folds.items())}
    acc: Dict[str, list] = {}
    for f, st in fold_stats.items():
        for k, v in st.items():
            acc.setdefault(k, []).append(v["media"])
    glob = M.calcular_resumen_experimento(acc) if len(fold_stats) > 1 else {}
    return {"folds": fold_stats, "global": glob}
