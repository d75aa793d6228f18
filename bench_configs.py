"""The BASELINE.json configurations besides the headline one (bench.py --config ...):

  cohort22  configs[2]: 22-volume test cohort, per-plane reconstruction + 3-plane consensus + DSC / IoU / precision / recall (1 GPU)
  cohort75  configs[3]: 75 volumes, 5 folds, the whole chain enhance->slice->recon->consensus->eval->promediar_folds, patients sharded
            with dist.shard_patients over the ranks, ONE NCCL all-reduce of the [75, 4, 4] count table per pass, per-patient
            metrics / fold means / global summary asserted against the oracle on EVERY rank
  stress    configs[4]: 1,024 volumes, CLAHE tri-planar + consensus / eval, 1024 / world volumes per rank
  dropin    the path the import swap takes: lesion flags -> P50 central window -> one enhancement, one plane, slice lists
            (enhance_slices_kernel) with its own roofline line

Patients: four CPU-generated base patients (mslesseg_b200.synthetic, the same on every rank); the patient at position j of a
rank's shard uses base j % 4, its FLAIR re-scaled by a factor that depends on its id (integer-valued, own per-slice statistics).
Every rank can therefore recompute the expected count table of the WHOLE cohort from the oracle counts of the four bases.
"""
from __future__ import annotations

import json
import os
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
N_VOX = 182 * 218 * 182
PLANOS = ("axial", "coronal", "sagital")
MEJORAS = ("HE", "CLAHE", "GC", "LT")
CHUNK = 32

CONFIGS = {
    "cohort22": dict(n=22, k_folds=None, mejoras=(), first_id=54, what="configs[2]: 22-volume test cohort, recon x3 -> consensus -> counts -> DSC / IoU / precision / recall"),
    "cohort75": dict(n=75, k_folds=5, mejoras=MEJORAS, first_id=1, what="configs[3]: 75 volumes, 5 folds, enhance->slice (4 enhancements x 3 planes, all slices) -> recon -> consensus -> eval -> promediar_folds"),
    "stress": dict(n=1024, k_folds=5, mejoras=("CLAHE",), first_id=1, what="configs[4]: 1,024 volumes, CLAHE x 3 planes (all slices) + recon -> consensus -> eval"),
}


def _peak():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def _scale_of(pid: str) -> float:
    return 0.6 + 0.8 * ((int(pid[1:]) * 0.6180339887) % 1.0)


def run_cohort(args, name, emit, ClockSampler, METRIC, UNIT):
    import torch
    import torch.distributed as dist
    from mslesseg_b200 import _lib, ops, dist as D, metrics as M, synthetic as S
    from oracle import oracle as O
    cfg = CONFIGS[name]
    _lib.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    n_all, k_folds = cfg["n"], cfg["k_folds"]
    ids_all = [f"P{cfg['first_id'] + i}" for i in range(n_all)]
    n_ids = cfg["first_id"] + n_all - 1
    shards = [D.shard_patients(ids_all, world, r, k_folds=k_folds, n_ids=n_ids) for r in range(world)]
    mine = shards[rank]
    L = len(mine)
    mejoras = tuple(cfg["mejoras"])
    # ---- data
    base = [S.make_patient(1 + b, config_id=4, num_cortes=args.num_cortes) for b in range(4)]
    bflair = torch.from_numpy(np.stack([p.flair for p in base])).to(device)
    bgt = torch.from_numpy(np.stack([p.gt for p in base])).to(device)
    gt = torch.empty((L, 182, 218, 182), dtype=torch.uint8, device=device)
    flair = torch.empty((L, 182, 218, 182), dtype=torch.float32, device=device) if mejoras else None
    for j, pid in enumerate(mine):
        gt[j] = bgt[j % 4]
        if mejoras:
            flair[j] = torch.round(bflair[j % 4] * _scale_of(pid))
    # predicted slices of a full chunk (position j -> base j % 4: every chunk that starts at a multiple of 4 shares them)
    def chunk_preds(n):
        out = {}
        for pl in PLANOS:
            sl = np.concatenate([base[j % 4].pred_slices[pl] for j in range(n)])
            vs = np.concatenate([np.full(len(base[j % 4].pred_indices[pl]), j, np.int32) for j in range(n)])
            ix = np.concatenate([np.asarray(base[j % 4].pred_indices[pl], np.int32) for j in range(n)])
            out[pl] = (torch.from_numpy(sl).to(device), torch.from_numpy(vs).to(device), torch.from_numpy(ix).to(device))
        return out
    chunks = [(c0, min(CHUNK, L - c0)) for c0 in range(0, L, CHUNK)]
    preds_by_n = {n: chunk_preds(n) for n in sorted({n for _, n in chunks})}
    nmax = max((n for _, n in chunks), default=0)
    outs = {}
    for m in mejoras:
        for pl in PLANOS:
            n_p, rows, cols = ops.plane_dims(pl, 182, 218, 182)
            outs[(m, pl)] = torch.empty((nmax, n_p, cols, rows), dtype=torch.uint8, device=device)
    ws = torch.empty(ops.enhance_volumes_workspace_bytes(max(nmax, 1), 182, 218, 182), dtype=torch.uint8, device=device) if mejoras else None
    rvol = {pl: torch.empty((max(nmax, 1), 182, 218, 182), dtype=torch.uint8, device=device) for pl in PLANOS}
    order = {p: i for i, p in enumerate(D.sort_patients(ids_all))}
    rows_mine = torch.as_tensor([order[p] for p in mine], dtype=torch.long, device=device)
    table = torch.zeros((n_all, 4, 4), dtype=torch.int64, device=device)
    h_table = torch.zeros((n_all, 4, 4), dtype=torch.int64).pin_memory()
    local_counts = torch.zeros((max(L, 1), 4, 4), dtype=torch.int64, device=device)
    state = {}

    def one_pass():
        """The whole cohort once on this rank's shard; the count table is all-reduced ONCE, after the last chunk."""
        for c0, n in chunks:
            if mejoras:
                state["flags"] = ops.lesion_slices(gt[c0:c0 + n])
                o = {k: v[:n] for k, v in outs.items()}
                ops.enhance_volumes(flair[c0:c0 + n], mejoras, PLANOS, outs=o, workspace=ws)
            pr = preds_by_n[n]
            for pl in PLANOS:
                sl, vs, ix = pr[pl]
                ops.recon(sl, vs, ix, pl, n, S.SHAPE_XYZ, out=rvol[pl][:n])
            cons, counts = ops.consensus_eval(rvol["axial"][:n], rvol["coronal"][:n], rvol["sagital"][:n], gt[c0:c0 + n], 2)
            local_counts[c0:c0 + n] = counts
            state["cons"] = cons
        table.zero_()
        if L:
            table[rows_mine] = local_counts[:L]
        if world > 1:
            dist.all_reduce(table)                 # NCCL SUM of the int64 count table: once per cohort
        h_table.copy_(table, non_blocking=True)    # the table goes to the host, where the float64 formulas run

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(args.warmup, 3)
    for _ in range(W):
        one_pass()
    barrier()
    l0 = sum(_lib.kernel_launches().values())
    one_pass()
    barrier()
    launches_per_pass = sum(_lib.kernel_launches().values()) - l0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_pass()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = n_all * N_VOX / (ms_per_step * 1e-3) / 1e9

    # ---- host statistics (the reference's float64 formulas; timed separately, outside the device pass)
    t0 = time.perf_counter()
    tab = h_table.numpy().copy()
    per_patient = D.metrics_from_table(ids_all, torch.from_numpy(tab))
    stats = {}
    if k_folds:
        for k, pl in enumerate(D.PLANOS4):
            stats[pl] = D.fold_and_global_stats({p: per_patient[p][pl] for p in ids_all}, k_folds, n_ids)
    iou = {p: {pl: M.iou_desde_conteos(*tab[order[p], k, :3]) for k, pl in enumerate(D.PLANOS4)} for p in ids_all}
    host_ms = (time.perf_counter() - t0) * 1e3

    # ---- verification on EVERY rank against the oracle (counts of the four base patients -> expected table of the cohort)
    exp_counts = []
    for b in range(4):
        gt0 = S.as_xyz(base[b].gt)
        vols = [O.reconstruir(base[b].pred_slices[pl], base[b].pred_indices[pl], S.SHAPE_XYZ, pl) for pl in PLANOS]
        cons = O.combinar_volumenes(vols[0].astype(np.float64), vols[1].astype(np.float64), vols[2].astype(np.float64), 2)
        exp_counts.append([list(O.confusion_counts(gt0, v)) for v in vols + [cons]])
    expected = np.zeros((n_all, 4, 4), np.int64)
    for r in range(world):
        for j, p in enumerate(shards[r]):
            expected[order[p]] = exp_counts[j % 4]
    ok = bool(np.array_equal(tab, expected))
    exp_pp = {p: {pl: O.metricas_desde_conteos(*expected[order[p], k]) for k, pl in enumerate(D.PLANOS4)} for p in ids_all}
    ok = ok and _same(per_patient, exp_pp)
    if k_folds:
        for pl in D.PLANOS4:
            folds = {}
            for p in D.sort_patients(ids_all):
                f = O.calcular_fold(p, k_folds, n_ids)
                for kk, v in exp_pp[p][pl].items():
                    folds.setdefault(f, {}).setdefault(kk, []).append(v)
            fstats = {f: O.calcular_promedio(d) for f, d in sorted(folds.items())}
            acc = {}
            for f, st in fstats.items():
                for kk, v in st.items():
                    acc.setdefault(kk, []).append(v["media"])
            ok = ok and _same(stats[pl]["folds"], fstats) and _same(stats[pl]["global"], O.calcular_resumen_experimento(acc))
    if mejoras and L:       # enhancement outputs of the last chunk against the oracle (a few slices of its first and last patient)
        import warnings
        c0, n = chunks[-1]
        for v in sorted({0, n - 1}):
            vxyz = S.as_xyz(flair[c0 + v].cpu().numpy()).astype(np.float64)
            for pl in PLANOS:
                n_p = vxyz.shape[O.plane_axis(pl)]
                for i in (n_p // 2, 1):
                    for m in mejoras:
                        with warnings.catch_warnings():
                            warnings.simplefilter("ignore")
                            want = O.png_orient(O.enhance_slice(O.slice_of(vxyz, pl, i), m))
                        ok = ok and bool(np.array_equal(outs[(m, pl)][v, i].cpu().numpy(), want))
    if world > 1:
        f = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        ok_all = bool(f.item())
    else:
        ok_all = ok

    # ---- roofline of the dominant kernel (one extra profiled pass)
    _lib.profile_enable(True)
    one_pass()
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    peak, peak_src = _peak()
    pred_bytes = sum(int(preds_by_n[n][pl][0].numel()) for _, n in chunks for pl in PLANOS)
    alg = {"enhance_dense": (1 + len(mejoras)) * 3 * L * N_VOX, "plane_stats_f32": 4 * L * N_VOX, "norm_scatter": (4 + 3) * L * N_VOX,
           "lesion_flags": L * N_VOX, "recon_gather": pred_bytes + 3 * L * N_VOX, "consensus_eval": 5 * L * N_VOX}
    kernels = {}
    for kname, (kms, cnt) in prof.items():
        kernels[kname] = {"ms_per_pass": kms, "launches_per_pass": cnt}
        if kname in alg and kms > 0:
            kernels[kname]["gb_s"] = alg[kname] / kms / 1e6
    roof = None
    if kernels:
        dom = max(kernels, key=lambda k: kernels[k]["ms_per_pass"])
        dk = kernels[dom]
        roof = {"bound": "hbm", "kernel": dom, "achieved": dk.get("gb_s", 0.0), "peak": peak, "unit": "GB/s", "frac": dk.get("gb_s", 0.0) / peak,
                "traffic": None, "peak_source": peak_src, "avg_launch_ms": dk["ms_per_pass"] / max(1, dk["launches_per_pass"]),
                "algorithmic_bytes_per_launch": alg.get(dom, 0) / max(1, dk["launches_per_pass"]),
                "share_of_pass": dk["ms_per_pass"] / sum(k["ms_per_pass"] for k in kernels.values()),
                "note": "this rank's shard; algorithmic bytes as in DESIGN.md section 3"}
    bpv = (17 if len(mejoras) == 4 else (1 + 4 + 3 * len(mejoras)) if mejoras else 0) + 5
    if rank == 0:
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u8 (int64 counts)", "data": "synthetic",
            "config": {"workload": cfg["what"], "name": name, "patients": n_all, "patients_per_rank": [len(s) for s in shards], "k_folds": k_folds,
                       "sharding": "fold-major (dist.shard_patients)" if k_folds and k_folds >= world else "round-robin on the sorted list",
                       "chunk": CHUNK, "pred_slices_per_plane": args.num_cortes, "collective": "one all_reduce(SUM) of the int64 [n, 4, 4] table per pass" if world > 1 else "none",
                       "l2": "inputs per pass exceed the 126 MB L2; no flush needed"},
            "e2e": None, "gpu_launches": int(launches_per_pass * args.steps), "clocks": clocks, "roofline": roof, "cpu_baseline": None,
            "step_level": {"algorithmic_bytes_per_voxel": bpv, "gb_s": bpv * n_all * N_VOX / ms_per_step / 1e6 / world,
                           "frac_per_gpu": bpv * n_all * N_VOX / ms_per_step / 1e6 / world / peak},
            "kernels": kernels, "host_stats_ms": host_ms, "verified": ok_all,
            "verified_what": "count table, per-patient metrics" + (", fold means / std, global summary" if k_folds else "") + (", enhancement slices" if mejoras else "") + " against the oracle on every rank",
            "sample": {"global_consenso": stats.get("consenso", {}).get("global") if k_folds else None,
                       "first_patient": {"metrics": per_patient[ids_all[0]]["consenso"], "IoU": iou[ids_all[0]]["consenso"]}},
        })
    if world > 1:
        dist.destroy_process_group()
    return 0


def _same(a, b):
    import math
    if isinstance(a, float) and isinstance(b, float):
        return a == b or (math.isnan(a) and math.isnan(b))
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    return a == b


def run_dropin(args, emit, ClockSampler, METRIC, UNIT):
    """The slice-list path of the import swap (compat.Paciente.cortes_con_lesion_gris -> ops.enhance_slices): lesion flags,
    the cohort's P50 number of slices, the central window of every patient, ONE enhancement on ONE plane."""
    import torch
    import warnings
    from mslesseg_b200 import _lib, ops, metrics as M, synthetic as S
    from oracle import oracle as O
    _lib.load()
    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)
    B = args.batch
    mejora, plano = "CLAHE", "axial"
    base = [S.make_patient(1 + b, config_id=4, num_cortes=args.num_cortes) for b in range(4)]
    bflair = torch.from_numpy(np.stack([p.flair for p in base])).to(device)
    bgt = torch.from_numpy(np.stack([p.gt for p in base])).to(device)
    flair = torch.stack([torch.round(bflair[b % 4] * (0.6 + 0.8 * ((b * 0.6180339887) % 1.0))) if b >= 4 else bflair[b] for b in range(B)])
    gt = torch.stack([bgt[b % 4] for b in range(B)])
    k = PLANOS.index(plano)
    n_p, rows, cols = ops.plane_dims(plano, 182, 218, 182)
    state = {}

    def select():
        flags = ops.lesion_slices(gt)[k].cpu().numpy()                  # E0 on the GPU, the list arithmetic on the host
        idx = [np.flatnonzero(f).tolist() for f in flags]
        ncortes = M.num_cortes_percentil([len(i) for i in idx], 50)
        use = [M.ventana_central(i, ncortes) for i in idx]
        vs = np.concatenate([np.full(len(u), b, np.int32) for b, u in enumerate(use)])
        ix = np.concatenate([np.asarray(u, np.int32) for u in use])
        return ncortes, use, vs, ix          # host lists, as compat.Paciente passes them (range-checked by ops, uploaded per call)

    ncortes, use, vs, ix = select()
    ns = int(ix.size)
    out = torch.empty((ns, cols, rows), dtype=torch.uint8, device=device)
    from mslesseg_b200 import _lib as _L

    def launches():
        import ctypes
        lib = _L.load()
        n = lib.msl_kernel_kinds()
        per = (ctypes.c_ulonglong * n)()
        lib.msl_kernel_launches(per)
        return {lib.msl_kernel_name(i).decode(): int(per[i]) for i in range(n) if per[i]}

    def step():
        state["out"] = ops.enhance_slices(flair, mejora, plano, vs, ix, layout="P", out=out)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    l0 = launches()
    step()
    l1 = launches()
    per_step = {k_: l1[k_] - l0.get(k_, 0) for k_ in l1 if l1[k_] - l0.get(k_, 0)}
    torch.cuda.synchronize()
    _L.profile_enable(True)
    step()
    torch.cuda.synchronize()
    kernel_ms = {k_: round(v_[0], 4) for k_, v_ in _L.profile_collect().items()}
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the selected slices of 32 patients (~50 MB) fit the L2: a 256 MB scratch write between iterations flushes it
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    tot = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    clocks = sampler.stop()
    ms = tot / args.steps
    t0 = time.perf_counter()
    select()
    torch.cuda.synchronize()
    select_ms = (time.perf_counter() - t0) * 1e3
    # verification: every selected slice of the first and last patient against the oracle
    ok = True
    got = state["out"].cpu().numpy()
    pos = 0
    for b, u in enumerate(use):
        if b in (0, B - 1):
            vxyz = S.as_xyz(flair[b].cpu().numpy()).astype(np.float64)
            ok = ok and u == O.indices_a_usar(S.as_xyz(gt[b].cpu().numpy()), plano, ncortes)
            for j, i in enumerate(u):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    ok = ok and bool(np.array_equal(got[pos + j], O.png_orient(O.enhance_slice(O.slice_of(vxyz, plano, i), mejora))))
        pos += len(u)
    peak, peak_src = _peak()
    alg = ns * rows * cols * (4 + 1)
    emit({
        "metric": METRIC, "value": ns * rows * cols / (ms * 1e-3) / 1e9, "unit": "Gvoxel/s (selected slice pixels)", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u8",
        "data": "synthetic",
        "config": {"workload": f"drop-in slice-list path: {B} patients, lesion flags -> P50 = {ncortes} slices per patient (central window) -> {mejora} on the {plano} slices ({ns} slices) through ops.enhance_slices (E1 per slice into a staged stack, then the dense kernel over the stack)",
                   "kernels_per_step": per_step, "kernel_ms_per_step": kernel_ms,
                   "name": "dropin", "l2": "256 MB scratch write between timed iterations (the slices of a step fit the L2)"},
        "e2e": None, "gpu_launches": int(args.steps) * sum(per_step.values()), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": " + ".join(sorted(per_step)), "achieved": alg / ms / 1e6, "peak": peak, "unit": "GB/s",
                     "frac": alg / ms / 1e6 / peak, "traffic": None, "peak_source": peak_src, "avg_launch_ms": ms,
                     "algorithmic_bytes_per_launch": alg, "note": "4 B float32 read + 1 B uint8 written per selected pixel; avg_launch_ms = the whole call (both kernels + the upload of the index lists)"},
        "cpu_baseline": None, "selection_ms": select_ms, "verified": bool(ok),
    })
    return 0
