"""CPU, world_size 2, gloo: patient sharding + the count-table all-reduce (the only collective of the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mslesseg_b200 import dist as D
from mslesseg_b200 import metrics as M


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_counts(pid):
    rng = np.random.default_rng(D.patient_number(pid))
    c = rng.integers(0, 5000, (4, 4)).astype(np.int64)
    c[:, 3] = 7_221_032 - c[:, :3].sum(axis=1)
    return c


def _worker(rank, world, port, ids, k_folds, n_ids, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = D.shard_patients(ids, world, rank, k_folds, n_ids)
    local = torch.from_numpy(np.stack([_fake_counts(p) for p in mine])) if mine else torch.zeros((0, 4, 4), dtype=torch.int64)
    table = D.all_reduce_count_table(ids, mine, local)
    lesion = D.all_reduce_lesion_counts(ids, mine, [D.patient_number(p) % 37 for p in mine])
    torch.save({"mine": mine, "table": table, "lesion": lesion}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("k_folds,world", [(5, 2), (None, 2)])
def test_sharded_count_table(tmp_path, k_folds, world):
    ids = [f"P{n}" for n in range(1, 24)]
    port = _free_port()
    mp.spawn(_worker, args=(world, port, ids, k_folds, 23, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    # every patient on exactly one rank
    assert sorted(sum((r["mine"] for r in res), []), key=D.patient_number) == D.sort_patients(ids)
    want = np.stack([_fake_counts(p) for p in D.sort_patients(ids)])
    for r in res:
        assert np.array_equal(r["table"].numpy(), want)
        assert r["lesion"] == [D.patient_number(p) % 37 for p in D.sort_patients(ids)]
    if k_folds:
        for r, rr in enumerate(res):      # fold-major: a rank owns whole folds
            assert {(M.calcular_fold(p, k_folds, 23) - 1) % world for p in rr["mine"]} <= {r}
    # per-patient metrics from the table, then fold / global statistics like the reference computes them
    per = D.metrics_from_table(ids, res[0]["table"])
    assert set(per["P1"]) == set(D.PLANOS4)
    stats = D.fold_and_global_stats({p: d["consenso"] for p, d in per.items()}, 5, 23)
    assert sorted(stats["folds"]) == [1, 2, 3, 4, 5] and set(stats["global"]) == {"DSC", "AUC", "Precision", "Recall"}
    dsc = [per[p]["consenso"]["DSC"] for p in D.sort_patients(ids) if M.calcular_fold(p, 5, 23) == 1]
    assert stats["folds"][1]["DSC"] == {"media": float(np.round(np.mean(dsc), 3)), "std": float(np.round(np.std(dsc), 3))}


def test_single_process_is_identity():
    ids = ["P3", "P1", "P2"]
    local = torch.from_numpy(np.stack([_fake_counts(p) for p in ids]))
    table = D.all_reduce_count_table(ids, ids, local)
    assert np.array_equal(table.numpy(), np.stack([_fake_counts(p) for p in ["P1", "P2", "P3"]]))
    assert D.shard_patients(ids, 1, 0) == ["P1", "P2", "P3"]
    assert D.shard_patients([f"P{n}" for n in range(1, 54)], 2, 1, k_folds=5) == [f"P{n}" for n in range(12, 23)] + [f"P{n}" for n in range(34, 44)]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_cohort75_sharding_partitions_the_cohort(world):
    """bench.py --config cohort75 / stress: every patient on exactly one rank; fold-major while there are at least as many
    folds as ranks (5 folds: 1, 2, 4 ranks), round-robin on the sorted list beyond that (8 ranks)."""
    ids = [f"P{n}" for n in range(1, 76)]
    shards = [D.shard_patients(ids, world, r, k_folds=5, n_ids=75) for r in range(world)]
    assert sorted(sum(shards, []), key=D.patient_number) == D.sort_patients(ids)
    if world <= 5:
        for r, sh in enumerate(shards):
            assert {(M.calcular_fold(p, 5, 75) - 1) % world for p in sh} <= {r}
    else:
        assert shards[3] == D.sort_patients(ids)[3::world]
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 15
    # the generalised fold rule keeps the reference's assignment for the real cohort (P1..P53, 5 folds)
    assert [M.calcular_fold(f"P{n}", 5) for n in (1, 11, 12, 22, 23, 33, 34, 43, 44, 53)] == [1, 1, 2, 2, 3, 3, 4, 4, 5, 5]
