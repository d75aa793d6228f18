#!/usr/bin/env python
"""bench.py - throughput of the YOLO-MSLesSeg voxel path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle/_ref), host cores

One STEP = one pass of the hot path over one batch of synthetic patients resident in HBM:
  enhance->slice : lesion-slice flags (E0) + HE / CLAHE / GC / LT x axial / coronal / sagital over ALL
                   slices of every volume (E1-E7, msl_enhance_volumes: 12 PNG-oriented uint8 stacks)
  recon->consensus->eval : stack the predicted masks of the three planes into volumes (R1-R2), tri-planar
                   vote fused with the 4x4 confusion counts (R3-R4); at N > 1 the int64 count table is
                   all-reduced over NCCL; the float64 metric formulas run on the host outside the step.
This is configs[1] of BASELINE.json ("single synthetic patient, all four enhancements + tri-planar
slicing") batched over `--batch` distinct patients per GPU so that the inputs (batch x 28.9 MB) exceed
the 126 MB L2 - no L2 flush is needed between iterations - followed by the output side of configs[2].

`value` = patients x 7,221,032 voxels / step time (data resident in HBM, CUDA events, max over ranks).
`e2e`   = same metric through the public API with pinned HOST buffers: H2D of the volumes / masks /
          predicted slices and D2H of every result (12 stacks, 3 recon volumes, consensus, counts)
          inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for _p in (str(ROOT), str(ROOT / "yolo-mslesseg_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

N_VOX = 182 * 218 * 182
METRIC = "Gvoxel/s enhance->slice (HE/CLAHE/GC/LT x 3 planes) + recon->consensus->eval"
UNIT = "Gvoxel/s"
PLANOS = ("axial", "coronal", "sagital")
MEJORAS = ("HE", "CLAHE", "GC", "LT")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="default", choices=["default", "cohort22", "cohort75", "stress", "dropin"],
                    help="default = the headline workload (configs[1] batched + configs[2] output side); the others are BASELINE.json's remaining configs (bench_configs.py)")
    ap.add_argument("--batch", type=int, default=32, help="synthetic patients per GPU per step")
    ap.add_argument("--num-cortes", type=int, default=40, help="predicted slices kept per plane (indices_a_usar)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-full-copies", action="store_true", help="end-to-end arm: copy every result byte to the host instead of the non-zero boxes")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e", default="files", choices=["files", "arrays"], help="end-to-end arm: host buffers hold the stages' FILES (.nii.gz / PNG, codec on the GPU; default) or raw arrays (round-1 arm)")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--overlap", action="store_true", help="run the output side (recon -> consensus -> eval) on a second stream; measured slower than one stream since the kernels got faster: 3.24 vs 3.19 ms per step")
    ap.add_argument("--no-overlap", action="store_true", help="(default; kept for older command lines) one stream")
    ap.add_argument("--no-graph", action="store_true", help="launch every step kernel by kernel instead of replaying a CUDA graph of one step (the NCCL all-reduce stays outside the graph)")
    return ap.parse_args()


def workload_config(args, extra=None):
    cfg = {
        "workload": (f"configs[1] batched: {args.batch} synthetic 182x218x182 patients per GPU per step, "
                     "HE+CLAHE+GC+LT x axial+coronal+sagital over all 582 slices, then recon x3 -> consensus(umbral 2) "
                     "-> 4x4 confusion counts (configs[2] output side)"),
        "patients_per_gpu": args.batch,
        "volume_shape_xyz": [182, 218, 182],
        "pred_slices_per_plane": args.num_cortes,
        "l2": "inputs per step (batch x 28.9 MB float32) exceed the 126 MB L2; no flush needed",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cpu_baseline as CB
    cores = CB.host_cores()
    npat = CB.default_sample_patients(cores)
    with CB.CpuBaseline(npat, cores, num_cortes=args.num_cortes) as cb:
        for _ in range(args.warmup):
            cb.step()
        times = [cb.step() for _ in range(args.steps)]
        desc = cb.describe()
    total = sum(times)
    value = cb.voxels_per_step * args.steps / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+u8 (f64 metrics)", "data": "synthetic",
        "config": workload_config(args, {"note": "each step is a bounded sample of the workload: "
                                         f"{npat} patients through the reference's CPU path"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": CB.KIND, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------ our arm
def build_inputs(args, torch, device, rank):
    """Distinct synthetic patients resident on `device`.  A handful of CPU-generated base patients
    (mslesseg_b200.synthetic) are re-scaled per volume on the device so that every
    volume has its own intensities / per-slice statistics while staying integer-valued."""
    from mslesseg_b200 import synthetic as S
    nbase = min(4, args.batch)
    base = [S.make_patient(1 + n + 100 * rank, config_id=4, num_cortes=args.num_cortes) for n in range(nbase)]
    bflair = torch.from_numpy(np.stack([p.flair for p in base])).to(device)
    bgt = torch.from_numpy(np.stack([p.gt for p in base])).to(device)
    B = args.batch
    flair = torch.empty((B, 182, 218, 182), dtype=torch.float32, device=device)
    gt = torch.empty((B, 182, 218, 182), dtype=torch.uint8, device=device)
    for b in range(B):
        s = 0.6 + 0.8 * ((b * 0.6180339887) % 1.0)
        flair[b] = torch.round(bflair[b % nbase] * s) if b >= nbase else bflair[b]
        gt[b] = bgt[b % nbase]
    preds = {}
    for pl in PLANOS:
        sl = np.concatenate([base[b % nbase].pred_slices[pl] for b in range(B)])
        vs = np.concatenate([np.full(len(base[b % nbase].pred_indices[pl]), b, np.int32) for b in range(B)])
        ix = np.concatenate([np.asarray(base[b % nbase].pred_indices[pl], np.int32) for b in range(B)])
        preds[pl] = (torch.from_numpy(sl).to(device), torch.from_numpy(vs).to(device), torch.from_numpy(ix).to(device))
    return base, flair, gt, preds


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mslesseg_b200 import _lib, ops, metrics as M, synthetic as S
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    _lib.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    B = args.batch
    base, flair, gt, preds = build_inputs(args, torch, device, rank)

    # pre-allocated outputs (the step allocates nothing)
    outs = {}
    for m in MEJORAS:
        for pl in PLANOS:
            n_p, rows, cols = ops.plane_dims(pl, 182, 218, 182)
            outs[(m, pl)] = torch.empty((B, n_p, cols, rows), dtype=torch.uint8, device=device)
    ws = torch.empty(ops.enhance_volumes_workspace_bytes(B, 182, 218, 182), dtype=torch.uint8, device=device)
    rvol = {pl: torch.empty((B, 182, 218, 182), dtype=torch.uint8, device=device) for pl in PLANOS}
    # Two tables: the all-reduce of step k runs asynchronously (NCCL's own stream) while step k + 1 computes
    tables = [torch.zeros((world * B, 4, 4), dtype=torch.int64, device=device) for _ in range(2)]
    works = [None, None]
    state = {"k": 0}

    side = torch.cuda.Stream(device=device)

    def output_side():
        for pl in PLANOS:
            sl, vs, ix = preds[pl]
            ops.recon(sl, vs, ix, pl, B, S.SHAPE_XYZ, out=rvol[pl])
        cons, counts = ops.consensus_eval(rvol["axial"], rvol["coronal"], rvol["sagital"], gt, 2)
        state["cons"], state["counts"] = cons, counts

    def exchange():
        """Once per cohort pass (= per step), off the critical path: the int64 count table is summed over NCCL
        asynchronously; the next step does not wait for it (it only waits before it reuses the same table, two steps later)."""
        if world > 1:
            k = state["k"] & 1
            if works[k] is not None:
                works[k].wait()
            t = tables[k]
            t.zero_()
            t[rank * B:(rank + 1) * B] = state["counts"]
            works[k] = dist.all_reduce(t, async_op=True)
            state["k"] += 1
            state["last_table"] = t

    def step():
        # The two halves of the path are independent (different inputs, different outputs); --overlap runs the output
        # side on a second stream.  One stream is the default: it measured faster once the kernels were tuned.
        cur = torch.cuda.current_stream()
        if not args.overlap:
            output_side()
        else:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                output_side()
        state["flags"] = ops.lesion_slices(gt)
        ops.enhance_volumes(flair, MEJORAS, PLANOS, outs=outs, workspace=ws)
        if args.overlap:
            cur.wait_stream(side)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
        exchange()
    barrier()
    launches0 = sum(_lib.kernel_launches().values())
    step()
    exchange()
    launches_per_step = sum(_lib.kernel_launches().values()) - launches0
    run_step = step
    use_graph = False
    if not args.no_graph and not args.overlap:
        # the ~25 launches of a step are captured once; replay removes the launch gaps between dependent kernels
        # (3.147 -> 3.111 ms per step).  Any capture problem falls back to plain launches.
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            run_step, use_graph = graph.replay, True
        except Exception as exc:      # noqa: BLE001
            print(f"[bench] CUDA graph capture failed ({exc!r}); launching kernel by kernel", file=sys.stderr)
            torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        run_step()
        exchange()
    for w_ in works:                      # the last exchanges complete inside the timed region
        if w_ is not None:
            w_.wait()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = launches_per_step * args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B * N_VOX / (ms_per_step * 1e-3) / 1e9

    # ---- per-stage device times (untimed extra passes, CUDA events on the current stream)
    def time_fn(fn, reps=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def stage_out():
        for pl in PLANOS:
            sl, vs, ix = preds[pl]
            ops.recon(sl, vs, ix, pl, B, S.SHAPE_XYZ, out=rvol[pl])
        ops.consensus_eval(rvol["axial"], rvol["coronal"], rvol["sagital"], gt, 2)

    t_in = time_fn(lambda: (ops.lesion_slices(gt), ops.enhance_volumes(flair, MEJORAS, PLANOS, outs=outs, workspace=ws)))
    t_out = time_fn(stage_out)
    pred_bytes = sum(int(preds[pl][0].numel()) for pl in PLANOS)
    stages = {
        "enhance_slice": {"ms": t_in, "gvoxel_s": B * N_VOX / t_in / 1e6, "algorithmic_bytes_per_voxel": 17,
                          "gb_s": 17 * B * N_VOX / t_in / 1e6,
                          "note": "1 B mask read (E0) + 4 B float32 read + 12 x 1 B uint8 written per input voxel"},
        "recon_consensus_eval": {"ms": t_out, "gvoxel_s": B * N_VOX / t_out / 1e6,
                                 "algorithmic_bytes": pred_bytes + 5 * B * N_VOX,
                                 "gb_s": (pred_bytes + 5 * B * N_VOX) / t_out / 1e6,
                                 "note": "predicted slices present + GT read, 3 recon volumes + consensus written"},
    }

    # ---- roofline of the dominant kernel: per-launch CUDA events recorded inside the library
    _lib.profile_enable(True)
    for _ in range(2):                 # one stream here: concurrent kernels would stretch each other's event times
        output_side()
        state["flags"] = ops.lesion_slices(gt)
        ops.enhance_volumes(flair, MEJORAS, PLANOS, outs=outs, workspace=ws)
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    npx = {"axial": 182 * 218, "coronal": 182 * 182, "sagital": 218 * 182}
    chunk = int(os.environ.get("MSL_VOLUME_CHUNK", "32"))
    nchunks = -(-B // chunk)
    alg_bytes = {   # algorithmic bytes over ONE step, per kernel kind
        "enhance_dense": (1 + 4) * 3 * B * N_VOX,      # per plane: staged uint8 slice in, HE + CLAHE + GC + LT out
        "plane_stats_f32": 4 * B * N_VOX, "norm_scatter": (4 + 3) * B * N_VOX, "lesion_flags": B * N_VOX,
        "recon_gather": pred_bytes + 3 * B * N_VOX, "consensus_eval": 5 * B * N_VOX,
    }
    kernels = {}
    for name, (kms, n) in prof.items():
        per_step_ms = kms / 2
        kernels[name] = {"ms_per_step": per_step_ms, "launches_per_step": n // 2}
        if name in alg_bytes:
            kernels[name]["gb_s"] = alg_bytes[name] / per_step_ms / 1e6
    if "recon_gather" in kernels:   # the volumes' zero fill and the slice map belong to the same algorithmic bytes
        t_recon = sum(kernels[n]["ms_per_step"] for n in ("recon_fill", "recon_slot_map", "recon_gather") if n in kernels)
        kernels["recon_gather"]["gb_s"] = alg_bytes["recon_gather"] / t_recon / 1e6
        kernels["recon_gather"]["gb_s_note"] = "over recon_fill + recon_slot_map + recon_gather"
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    dk = kernels[dom]
    achieved = dk.get("gb_s", 0.0)
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/r2_traffic.json):
    # measured dram__bytes_read.sum + dram__bytes_write.sum per pixel of that capture, scaled to this launch size.
    traffic, traffic_src = None, None
    tfile = ROOT / "profiles" / "r2_traffic.json"
    if tfile.exists():
        tj = json.loads(tfile.read_text()).get(dom)
        if tj:
            units_per_launch = alg_bytes.get(dom, 0) / max(1, dk["launches_per_step"]) / tj["algorithmic_bytes_per_unit"]
            traffic = tj["dram_bytes_per_unit"] * units_per_launch
            traffic_src = tj["source"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "avg_launch_ms": dk["ms_per_step"] / max(1, dk["launches_per_step"]),
                "algorithmic_bytes_per_launch": alg_bytes.get(dom, 0) / max(1, dk["launches_per_step"]),
                "share_of_step": dk["ms_per_step"] / sum(k["ms_per_step"] for k in kernels.values()),
                "step_level": {"algorithmic_bytes_per_voxel": 22, "gb_s": 22 * B * N_VOX / ms_per_step / 1e6,
                               "frac": 22 * B * N_VOX / ms_per_step / 1e6 / peak,
                               "note": "17 B/voxel input side + 5 B/voxel output side (recon volumes counted once)"}}

    # ---- verification of what the timed steps produced (oracle = checker only)
    verified = None
    if not args.no_verify:
        verified = verify(torch, ops, M, S, base, flair, outs, rvol, state)
        if world > 1:
            # every rank checks the reduced table row by row: the oracle counts of each rank's base patients are gathered
            # (tiny Python objects) and compared with what NCCL delivered; the flags are then AND-ed over the ranks
            from oracle import oracle as O
            mine = []
            for p in base:
                g0 = S.as_xyz(p.gt)
                vols = [O.reconstruir(p.pred_slices[pl], p.pred_indices[pl], S.SHAPE_XYZ, pl) for pl in PLANOS]
                cons = O.combinar_volumenes(vols[0].astype(np.float64), vols[1].astype(np.float64), vols[2].astype(np.float64), 2)
                mine.append([list(O.confusion_counts(g0, v)) for v in vols + [cons]])
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            expected = np.zeros((world * B, 4, 4), np.int64)
            for r in range(world):
                for b in range(B):
                    expected[r * B + b] = gathered[r][b % len(gathered[r])]
            verified = bool(verified) and bool(np.array_equal(state["last_table"].cpu().numpy(), expected))
            f = torch.tensor([1 if verified else 0], dtype=torch.int32, device=device)
            dist.all_reduce(f, op=dist.ReduceOp.MIN)
            verified = bool(f.item())

    # ---- end-to-end with host buffers
    e2e = None
    if not args.no_e2e:
        if args.e2e == "files":
            e2e = run_e2e_files(args, torch, dist, ops, S, device, world, flair, gt, preds)
        else:
            e2e = run_e2e(args, torch, dist, ops, S, device, world, flair, gt, preds)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline as CB
        cores = CB.host_cores()
        with CB.CpuBaseline(CB.default_sample_patients(cores), cores, num_cortes=args.num_cortes) as cb:
            cb.step()
            sec = cb.step()
            cpu = {"value": cb.voxels_per_step / sec / 1e9, "unit": UNIT, "cores": cores, "kind": CB.KIND, "sample": cb.describe()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+u8 (int64 counts)", "data": "synthetic", "config": workload_config(args, {"volume_chunk": chunk, "chunks": nchunks, "streams": 2 if args.overlap else 1, "cuda_graph": use_graph}),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "stages": stages, "kernels": kernels, "verified": verified,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def verify(torch, ops, M, S, base, flair, outs, rvol, state):
    """Spot-check the outputs of the timed steps against the oracle (volume 0 and a re-scaled one)."""
    import warnings
    from oracle import oracle as O
    ok = True
    for b in (0, flair.shape[0] - 1):
        vxyz = S.as_xyz(flair[b].cpu().numpy()).astype(np.float64)
        for pl in PLANOS:
            n_p = vxyz.shape[O.plane_axis(pl)]
            for i in (n_p // 2, n_p // 3, 1):
                for m in MEJORAS:
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        want = O.png_orient(O.enhance_slice(O.slice_of(vxyz, pl, i), m))
                    ok &= bool(np.array_equal(outs[(m, pl)][b, i].cpu().numpy(), want))
    p0 = base[0]
    gt0 = S.as_xyz(p0.gt)
    vols = [O.reconstruir(p0.pred_slices[pl], p0.pred_indices[pl], S.SHAPE_XYZ, pl) for pl in PLANOS]
    for k, pl in enumerate(PLANOS):
        ok &= bool(np.array_equal(S.as_xyz(rvol[pl][0].cpu().numpy()), vols[k].astype(np.uint8)))
    cons = O.combinar_volumenes(vols[0].astype(np.float64), vols[1].astype(np.float64), vols[2].astype(np.float64), 2)
    ok &= bool(np.array_equal(S.as_xyz(state["cons"][0].cpu().numpy()), cons))
    c = state["counts"][0].cpu().numpy()
    for k, v in enumerate(vols + [cons]):
        ok &= c[k].tolist() == list(O.confusion_counts(gt0, v))
    ok &= M.metricas_desde_conteos(*c[3]) == O.metricas_desde_conteos(*O.confusion_counts(gt0, cons))
    for k, pl in enumerate(PLANOS):
        ok &= np.flatnonzero(state["flags"][k][0].cpu().numpy()).tolist() == O.indices_cortes_con_lesion(gt0, pl)
    return bool(ok)


def run_e2e(args, torch, dist, ops, S, device, world, flair, gt, preds):
    """Same step through the public API with HOST buffers: pinned inputs -> H2D -> kernels -> D2H of every
    result, chunked over three streams so that copies in both directions overlap the kernels.  Results travel as their
    non-zero boxes (ops.nonzero_flags + ops.HostResult: skull-stripped volumes are two thirds background, masks ~99 %
    zeros); --e2e-full-copies copies every byte instead."""
    B = args.batch
    CH = 4
    nstream = 3
    boxed = not args.e2e_full_copies
    streams = [torch.cuda.Stream(device=device) for _ in range(nstream)]
    h_flair = flair.cpu().pin_memory()
    h_gt = gt.cpu().pin_memory()
    # per-volume predicted slices (host, pinned)
    h_pred = {}
    for pl in PLANOS:
        sl, vs, ix = (t.cpu() for t in preds[pl])
        h_pred[pl] = (sl.pin_memory(), vs.numpy(), ix.numpy())
    dims = {pl: ops.plane_dims(pl, 182, 218, 182) for pl in PLANOS}
    keys = [(m, pl) for m in MEJORAS for pl in PLANOS] + [("recon", pl) for pl in PLANOS] + [("consenso", "")]
    shape_of = {k: ((B, dims[k[1]][0], dims[k[1]][2], dims[k[1]][1]) if k[0] in MEJORAS else (B, 182, 218, 182)) for k in keys}
    h_res = {k: ops.HostResult(shape_of[k]) for k in keys}
    h_counts = torch.empty((B, 4, 4), dtype=torch.int64).pin_memory()
    nflag = sum(shape_of[k][1] + shape_of[k][2] for k in keys)
    flag_off, off = {}, 0
    for k in keys:                               # per key: [CH][A] then [CH][B] inside one flat flags buffer
        flag_off[k] = ((off, shape_of[k][1]), (off + CH * shape_of[k][1], shape_of[k][2]))
        off += CH * (shape_of[k][1] + shape_of[k][2])
    bufs = []
    for _ in range(nstream):
        d = {"flair": torch.empty((CH, 182, 218, 182), dtype=torch.float32, device=device),
             "gt": torch.empty((CH, 182, 218, 182), dtype=torch.uint8, device=device),
             "outs": {(m, pl): torch.empty((CH, dims[pl][0], dims[pl][2], dims[pl][1]), dtype=torch.uint8, device=device) for m in MEJORAS for pl in PLANOS},
             "ws": torch.empty(ops.enhance_volumes_workspace_bytes(CH, 182, 218, 182), dtype=torch.uint8, device=device),
             "rvol": {pl: torch.empty((CH, 182, 218, 182), dtype=torch.uint8, device=device) for pl in PLANOS},
             "flags_dev": torch.empty(CH * nflag, dtype=torch.uint8, device=device),
             "flags": torch.empty(CH * nflag, dtype=torch.uint8).pin_memory(),
             "event": torch.cuda.Event()}
        bufs.append(d)
    # slice ranges of each chunk in the concatenated prediction stacks
    ranges = {}
    for pl in PLANOS:
        vs = h_pred[pl][1]
        ranges[pl] = [(int(np.searchsorted(vs, c0)), int(np.searchsorted(vs, min(c0 + CH, B)))) for c0 in range(0, B, CH)]
    h2d = d2h = 0

    def compute(ci, c0, count):
        """H2D of the chunk's inputs and all kernels; in boxed mode also the non-zero flags of every result."""
        nonlocal h2d, d2h
        n = min(CH, B - c0)
        st, d = streams[ci % nstream], bufs[ci % nstream]
        with torch.cuda.stream(st):
            d["flair"][:n].copy_(h_flair[c0:c0 + n], non_blocking=True)
            d["gt"][:n].copy_(h_gt[c0:c0 + n], non_blocking=True)
            fl, g = d["flair"][:n], d["gt"][:n]
            flags = ops.lesion_slices(g)
            o = {k: t[:n] for k, t in d["outs"].items()}
            ops.enhance_volumes(fl, MEJORAS, PLANOS, outs=o, workspace=d["ws"])
            nb = fl.numel() * 4 + g.numel()
            for pl in PLANOS:
                a, b = ranges[pl][ci]
                sl = h_pred[pl][0][a:b].to(device, non_blocking=True)
                vs = torch.from_numpy(h_pred[pl][1][a:b] - c0).to(device, non_blocking=True)
                ix = torch.from_numpy(h_pred[pl][2][a:b]).to(device, non_blocking=True)
                nb += sl.numel() + 8 * (b - a)
                ops.recon(sl, vs, ix, pl, n, S.SHAPE_XYZ, out=d["rvol"][pl][:n])
            cons, counts = ops.consensus_eval(d["rvol"]["axial"][:n], d["rvol"]["coronal"][:n], d["rvol"]["sagital"][:n], g, 2)
            res = dict(o)
            for pl in PLANOS:
                res[("recon", pl)] = d["rvol"][pl][:n]
            res[("consenso", "")] = cons
            h_counts[c0:c0 + n].copy_(counts, non_blocking=True)
            nd = counts.numel() * 8 + sum(f.numel() for f in flags)
            if boxed:
                for k in keys:          # all flags land in one device buffer and go to the host in one copy
                    (oa, A_), (ob, B_) = flag_off[k]
                    ops.nonzero_flags(res[k], out=(d["flags_dev"][oa:oa + n * A_].view(n, A_), d["flags_dev"][ob:ob + n * B_].view(n, B_)))
                d["flags"].copy_(d["flags_dev"], non_blocking=True)
                nd += CH * nflag
            else:
                for k in keys:
                    h_res[k].host[c0:c0 + n].copy_(res[k], non_blocking=True); nd += res[k].numel()
            d["event"].record(st)
            d["keep"] = (res, counts, flags)
            if count:
                h2d += nb; d2h += nd

    def deliver(ci, c0, count):
        """Boxed mode: once the chunk's flags are on the host, enqueue the box copies of its results."""
        nonlocal d2h
        if not boxed:
            return
        n = min(CH, B - c0)
        st, d = streams[ci % nstream], bufs[ci % nstream]
        d["event"].synchronize()
        res = d["keep"][0]
        with torch.cuda.stream(st):
            for k in keys:
                (oa, A_), (ob, B_) = flag_off[k]
                fl = d["flags"].numpy()
                moved = h_res[k].update(c0, res[k], fl[oa:oa + n * A_].reshape(n, A_), fl[ob:ob + n * B_].reshape(n, B_))
                if count:
                    d2h += moved

    d2h_events = None        # profile pass: CUDA events around every chunk's device-to-host copies

    def e2e_step(count=False):
        chunks = list(enumerate(range(0, B, CH)))
        for i, (ci, c0) in enumerate(chunks):
            compute(ci, c0, count)
            if i > 0:                       # the previous chunk's copies go out while this chunk computes
                deliver(*chunks[i - 1], count)
        deliver(*chunks[-1], count)

    def sync_all():
        for st in streams:
            st.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e2e_step(); sync_all(); e2e_step(count=True)
    sync_all()
    K = max(3, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    sync_all()
    sec = (time.perf_counter() - t0) / K
    # the host copies equal the device results (last chunks still resident on the device)
    ok = True
    nchunks = len(range(0, B, CH))
    for ci in range(max(0, nchunks - nstream), nchunks):
        c0 = ci * CH
        n = min(CH, B - c0)
        res = bufs[ci % nstream]["keep"][0]
        for k in keys:
            ok = ok and bool(torch.equal(h_res[k].host[c0:c0 + n], res[k].cpu()))
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    return {"value": world * B * N_VOX / sec / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": sec * 1e3, "steps": K, "host_results_equal_device": ok,
            "note": ("pinned host buffers, 4-volume chunks over 3 streams; results are copied as their non-zero boxes "
                     "(full rows; the host arrays stay complete, background is zero)" if boxed else
                     "pinned host buffers, 4-volume chunks over 3 streams; every result byte copied") +
                    "; wall clock around synchronised steps"}


class _HostFiles:
    """n files back to back in ONE pinned host buffer (what a reader thread would fill from disk)."""

    def __init__(self, torch, blobs):
        self.off = np.zeros(len(blobs) + 1, np.int64)
        np.cumsum([len(b) for b in blobs], out=self.off[1:])
        self.buf = torch.empty(int(self.off[-1]) + 16, dtype=torch.uint8).pin_memory()
        hv = self.buf.numpy()
        hv[int(self.off[-1]):] = 0
        for i, b in enumerate(blobs):
            hv[self.off[i]:self.off[i + 1]] = np.frombuffer(b, np.uint8)
        self.np = hv


def run_e2e_files(args, torch, dist, ops, S, device, world, flair, gt, preds):
    """The step end to end on FILES: the host buffers hold what the reference's stages read and write - .nii.gz volumes
    (FLAIR, ground truth), predicted-mask PNGs in, PNG slices (12 stacks), reconstructed / consensus .nii.gz and the count
    table out.  H2D carries the compressed file bytes, inflate / PNG unfiltering / datatype conversion run on the GPU;
    the results are deflated into PNG files and gzip members on the GPU and only those bytes travel back.
    Chunks of 4 patients over 3 streams; per chunk the host walks the container indexes (numpy) and, once the chunk's
    sizes have arrived, enqueues the copy of its packed result bytes."""
    from mslesseg_b200 import codec
    B, CH, nstream = args.batch, int(os.environ.get("MSL_E2E_CH", "4")), int(os.environ.get("MSL_E2E_STREAMS", "4"))
    X, Y, Z = S.SHAPE_XYZ
    N = X * Y * Z
    aff = np.diag([1.0, 1.0, 1.0, 1.0])
    dims = {pl: ops.plane_dims(pl, X, Y, Z) for pl in PLANOS}          # (n_p, rows, cols)
    # ---------------- the cohort's files (outside the timed region)
    flair_blobs, gt_blobs = [], []
    for c0 in range(0, B, CH):
        ps = codec.nifti_gz_device(flair[c0:c0 + CH], aff)
        flair_blobs += [codec.nifti_gz_bytes(ps, f) for f in range(min(CH, B - c0))]
        ps = codec.nifti_gz_device(gt[c0:c0 + CH], aff, como_float32=True)       # the dataset's MASK files are float32
        gt_blobs += [codec.nifti_gz_bytes(ps, f) for f in range(min(CH, B - c0))]
    import cv2
    pred_files, pred_vs, pred_ix = {}, {}, {}
    for pl in PLANOS:
        sl, vs, ix = (t.cpu().numpy() for t in preds[pl])
        pred_files[pl] = [cv2.imencode(".png", q, [cv2.IMWRITE_PNG_COMPRESSION, 3])[1].tobytes() for q in sl]   # guardar_prediccion (generar_predicciones.py:153)
        pred_vs[pl], pred_ix[pl] = vs, ix
    HF, HG = _HostFiles(torch, flair_blobs), _HostFiles(torch, gt_blobs)
    HP = {pl: _HostFiles(torch, pred_files[pl]) for pl in PLANOS}
    ranges = {pl: [(int(np.searchsorted(pred_vs[pl], c0)), int(np.searchsorted(pred_vs[pl], min(c0 + CH, B)))) for c0 in range(0, B, CH)]
              for pl in PLANOS}
    RAWF = 352 + 4 * N                                                   # bytes of an inflated float32 .nii
    RAWP = (RAWF + 15) & ~15
    lib_bound = lambda n, cid, raw: int(_lib_mod.load().msl_deflate_bound(n, cid, raw))
    from mslesseg_b200 import _lib as _lib_mod
    max_in = max(int(HF.off[min(c0 + CH, B)] - HF.off[c0]) for c0 in range(0, B, CH)) + 64
    max_gt = max(int(HG.off[min(c0 + CH, B)] - HG.off[c0]) for c0 in range(0, B, CH)) + 64
    max_pr = {pl: max(int(HP[pl].off[b] - HP[pl].off[a]) for a, b in ranges[pl]) + 64 for pl in PLANOS}
    max_ps = {pl: max(b - a for a, b in ranges[pl]) for pl in PLANOS}
    gz_spv = -(-RAWF // codec.CHUNK)
    gz_spv8 = -(-(352 + N) // codec.CHUNK)
    streams = [torch.cuda.Stream(device=device) for _ in range(nstream)]
    slots = []
    for _ in range(nstream):
        d = {"in_f": torch.empty(max_in, dtype=torch.uint8, device=device), "in_g": torch.empty(max_gt, dtype=torch.uint8, device=device),
             "raw_f": torch.empty(CH * RAWP + 64, dtype=torch.uint8, device=device), "raw_g": torch.empty(CH * RAWP + 64, dtype=torch.uint8, device=device),
             "flair": torch.empty((CH, Z, Y, X), dtype=torch.float32, device=device), "gt": torch.empty((CH, Z, Y, X), dtype=torch.uint8, device=device),
             "inexact": torch.zeros(2, dtype=torch.int64, device=device),
             "ws": torch.empty(ops.enhance_volumes_workspace_bytes(CH, X, Y, Z), dtype=torch.uint8, device=device),
             "rvol": {pl: torch.empty((CH, Z, Y, X), dtype=torch.uint8, device=device) for pl in PLANOS},
             "event": torch.cuda.Event(), "done": torch.cuda.Event()}
        d["outs4"] = {pl: torch.empty((4, CH, dims[pl][0], dims[pl][2], dims[pl][1]), dtype=torch.uint8, device=device) for pl in PLANOS}
        # the PNG files of the three planes share one buffer (and one inflate launch per chunk)
        d["in_p"] = torch.empty(sum((max_pr[pl] + 15) & ~15 for pl in PLANOS) + 64, dtype=torch.uint8, device=device)
        d["raw_p"] = torch.empty(sum(max_ps[pl] * ((dims[pl][1] * (dims[pl][2] + 1) + 15) & ~15) for pl in PLANOS) + 64, dtype=torch.uint8, device=device)
        d["pred"] = {pl: torch.empty((max_ps[pl], dims[pl][1], dims[pl][2]), dtype=torch.uint8, device=device) for pl in PLANOS}
        npng = {pl: 4 * CH * dims[pl][0] for pl in PLANOS}
        rawpng = {pl: dims[pl][2] * (dims[pl][1] + 1) for pl in PLANOS}
        d["png_out"] = {pl: torch.empty(lib_bound(npng[pl], _lib_mod.Z_PNG, rawpng[pl]), dtype=torch.uint8, device=device) for pl in PLANOS}
        d["png_ws"] = torch.empty(max(int(_lib_mod.load().msl_deflate_workspace_bytes(npng[pl], _lib_mod.Z_PNG, rawpng[pl])) for pl in PLANOS),
                                  dtype=torch.uint8, device=device)
        d["gz_out"] = torch.empty(lib_bound(3 * CH * gz_spv, _lib_mod.Z_GZIP, codec.CHUNK), dtype=torch.uint8, device=device)
        d["gz8_out"] = torch.empty(lib_bound(CH * gz_spv8, _lib_mod.Z_GZIP, codec.CHUNK), dtype=torch.uint8, device=device)
        d["gz_ws"] = torch.empty(int(_lib_mod.load().msl_deflate_workspace_bytes(3 * CH * gz_spv, _lib_mod.Z_GZIP, codec.CHUNK)), dtype=torch.uint8, device=device)
        # host side of the slot: sizes first (small), then the packed bytes
        d["status"] = torch.zeros((2 * CH * (gz_spv + 2) + sum(max_ps.values()) + 8, 4), dtype=torch.int32, device=device)
        d["h_tot"] = torch.zeros(8, dtype=torch.int64).pin_memory()
        d["d_tot"] = torch.zeros(8, dtype=torch.int64, device=device)
        slots.append(d)
    nchunks = len(range(0, B, CH))
    # host results: per chunk the packed files of every kind + their offsets
    h_png = {pl: [torch.empty(slots[0]["png_out"][pl].numel(), dtype=torch.uint8).pin_memory() for _ in range(nchunks)] for pl in PLANOS} if B <= 32 else None
    if h_png is None:        # large batches: results of chunk c overwrite those of chunk c - nstream (they would be on disk by then)
        h_png = {pl: [torch.empty(slots[0]["png_out"][pl].numel(), dtype=torch.uint8).pin_memory() for _ in range(nstream)] for pl in PLANOS}
    nhost = len(h_png[PLANOS[0]])
    h_gz = [torch.empty(slots[0]["gz_out"].numel(), dtype=torch.uint8).pin_memory() for _ in range(nhost)]
    h_gz8 = [torch.empty(slots[0]["gz8_out"].numel(), dtype=torch.uint8).pin_memory() for _ in range(nhost)]
    pin = lambda n_, dt: torch.zeros(n_, dtype=dt).pin_memory()
    h_off = {pl: [pin(4 * CH * dims[pl][0] + 1, torch.int64) for _ in range(nhost)] for pl in PLANOS}
    h_off["gz"] = [[pin(CH * gz_spv + 1, torch.int64) for _ in range(3)] for _ in range(nhost)]
    h_off["gz8"] = [pin(CH * gz_spv8 + 1, torch.int64) for _ in range(nhost)]
    h_counts = torch.empty((B, 4, 4), dtype=torch.int64).pin_memory()
    h_status = [torch.zeros(tuple(slots[0]["status"].shape), dtype=torch.int32).pin_memory() for _ in range(nhost)]
    h_inexact = [pin(2, torch.int64) for _ in range(nhost)]
    n_status = [0] * nhost
    hdr_f32 = torch.from_numpy(np.frombuffer(codec.nifti_header_bytes((X, Y, Z), np.float32, aff), np.uint8).copy()).to(device)
    hdr_u8 = torch.from_numpy(np.frombuffer(codec.nifti_header_bytes((X, Y, Z), np.uint8, aff), np.uint8).copy()).to(device)
    h2d = d2h = 0
    # MSL_E2E_MAPPED=1: the pack kernels store straight into the pinned (mapped) host buffers instead of packing on the device
    # and copying each stack once its size is on the host.  Measured slower (44.5 vs 37.1 ms per step): the pack kernels then
    # run at the speed of the host link (35 GB/s) and keep their SM slots for that long.
    mapped = os.environ.get("MSL_E2E_MAPPED") is not None
    waits = [0.0]            # seconds the host spent blocked on the chunk events (the rest of a step is enqueue work)

    def member_offsets(H, c0, n, pitch):
        """src / dst offsets (int64, n_members + 1) of the gzip members of files c0 .. c0+n inside the chunk's buffers."""
        base = int(H.off[c0])
        so, do = [], []
        for v in range(n):
            a, b = int(H.off[c0 + v]), int(H.off[c0 + v + 1])
            tab = codec.gzip_member_table(H.np[a:b])
            if tab is None:
                raise RuntimeError("input .nii.gz without the member index")
            so.append(tab[:, 0] + (a - base))
            d0 = np.zeros(len(tab), np.int64)
            np.cumsum(tab[:-1, 2], out=d0[1:])
            do.append(d0 + v * pitch)
        so.append(np.asarray([int(H.off[c0 + n]) - base], np.int64))
        do.append(np.asarray([n * pitch], np.int64))
        return np.concatenate(so), np.concatenate(do)

    def compute(ci, c0, count):
        nonlocal h2d, d2h
        n = min(CH, B - c0)
        st, d = streams[ci % nstream], slots[ci % nstream]
        hi = ci % nhost
        with torch.cuda.stream(st):
            nb = 0
            srow = 0                # rows of the slot's status tensor used so far
            # ---- inputs: file bytes up, inflate, convert
            for H, key_in, key_raw, out_t, k in ((HF, "in_f", "raw_f", d["flair"], 0), (HG, "in_g", "raw_g", d["gt"], 1)):
                a, b = int(H.off[c0]), int(H.off[c0 + n])
                d[key_in][:b - a].copy_(H.buf[a:b], non_blocking=True)
                so, do = member_offsets(H, c0, n, RAWP)
                so_d, do_d = torch.from_numpy(so).to(device, non_blocking=True), torch.from_numpy(do).to(device, non_blocking=True)
                codec.inflate_device(d[key_in], so_d, d[key_raw], do_d, "gzip", status=d["status"][srow:srow + len(so) - 1])
                srow += len(so) - 1
                for v in range(n):
                    codec.nifti_convert_device(d[key_raw][v * RAWP + 352:], 16, out_t[v], d["inexact"][k:k + 1])
                nb += (b - a) + so.nbytes + do.nbytes
            fl, g = d["flair"][:n], d["gt"][:n]
            flags = ops.lesion_slices(g)
            o = {(m, pl): d["outs4"][pl][k, :n] for k, m in enumerate(MEJORAS) for pl in PLANOS}
            if n == CH:
                ops.enhance_volumes(fl, MEJORAS, PLANOS, outs=o, workspace=d["ws"])
            else:
                o = ops.enhance_volumes(fl, MEJORAS, PLANOS, workspace=d["ws"])
            # ---- predicted-mask PNGs: file bytes up, ONE inflate launch for the three planes, then unfilter + stack per plane
            so_all, do_all, per_plane = [], [], {}
            in_base = raw_base = nstr = 0
            for pl in PLANOS:
                a, b = ranges[pl][ci]
                fa, fb = int(HP[pl].off[a]), int(HP[pl].off[b])
                d["in_p"][in_base:in_base + fb - fa].copy_(HP[pl].buf[fa:fb], non_blocking=True)
                w, h, bpp, istart, ilen = codec.png_table(HP[pl].np, HP[pl].off[a:b + 1])
                rp = (h * (w * bpp + 1) + 15) & ~15
                so_all.append(istart - fa + in_base)      # stream i ends where i+1 starts: the chunk tails (CRC, IEND, next header) are ignored by the zlib reader
                do_all.append(raw_base + np.arange(b - a, dtype=np.int64) * rp)
                per_plane[pl] = (a, b, h, w, bpp, nstr)
                nstr += b - a
                in_base += (fb - fa + 15) & ~15
                raw_base += (b - a) * rp
                nb += (fb - fa) + 2 * 8 * (b - a) + 8 * (b - a)
            so = np.concatenate(so_all + [np.asarray([in_base], np.int64)]).astype(np.int64)
            do = np.concatenate(do_all + [np.asarray([raw_base], np.int64)]).astype(np.int64)
            so_d, do_d = torch.from_numpy(so).to(device, non_blocking=True), torch.from_numpy(do).to(device, non_blocking=True)
            codec.inflate_device(d["in_p"], so_d, d["raw_p"], do_d, "zlib", status=d["status"][srow:srow + nstr])
            srow += nstr
            for pl in PLANOS:
                a, b, h, w, bpp, s0 = per_plane[pl]
                pred = d["pred"][pl][:b - a]
                codec.png_unfilter_device(d["raw_p"], do_d[s0:s0 + (b - a) + 1], h, w, bpp, pred)
                vs = torch.from_numpy(pred_vs[pl][a:b] - c0).to(device, non_blocking=True)
                ix = torch.from_numpy(pred_ix[pl][a:b]).to(device, non_blocking=True)
                ops.recon(pred, vs, ix, pl, n, S.SHAPE_XYZ, out=d["rvol"][pl][:n])
            cons, counts = ops.consensus_eval(d["rvol"]["axial"][:n], d["rvol"]["coronal"][:n], d["rvol"]["sagital"][:n], g, 2)
            h_counts[c0:c0 + n].copy_(counts, non_blocking=True)
            # ---- results: PNG files of the 12 stacks, .nii.gz of the three reconstructions (float32) and the consensus (uint8)
            res = {}
            for k, pl in enumerate(PLANOS):
                n_p, rows, cols = dims[pl]
                px = d["outs4"][pl] if n == CH else torch.stack([o[(m, pl)] for m in MEJORAS])
                res[pl] = ops.png_encode(px.reshape(-1, cols, rows), out=h_png[pl][hi] if mapped else d["png_out"][pl], workspace=d["png_ws"])
                d["d_tot"][k:k + 1].copy_(res[pl].off[-1:], non_blocking=True)
            third = d["gz_out"].numel() // 3
            gz_dst = h_gz[hi] if mapped else d["gz_out"]
            res["gz"] = [ops.deflate_files(d["rvol"][pl][:n], prefix=hdr_f32, expand_u8_to_f32=True, chunk_len=codec.CHUNK, container="gzip",
                                           dist2=4, workspace=d["gz_ws"], out=gz_dst[k * third:(k + 1) * third]) for k, pl in enumerate(PLANOS)]
            res["gz8"] = ops.deflate_files(cons, prefix=hdr_u8, chunk_len=codec.CHUNK, container="gzip", dist2=0,
                                           out=h_gz8[hi] if mapped else d["gz8_out"], workspace=d["gz_ws"])
            for k in range(3):
                d["d_tot"][3 + k:4 + k].copy_(res["gz"][k].off[-1:], non_blocking=True)
            d["d_tot"][6:7].copy_(res["gz8"].off[-1:], non_blocking=True)
            d["h_tot"].copy_(d["d_tot"], non_blocking=True)
            d["event"].record(st)
            d["keep"] = (res, o, cons, srow, flags, n, hi)
            if count:
                h2d += nb
                d2h += counts.numel() * 8 + 64

    def deliver(ci, c0, count):
        """The chunk's sizes are on the host: enqueue the copies of exactly the bytes its files occupy (+ offsets, status)."""
        nonlocal d2h
        st, d = streams[ci % nstream], slots[ci % nstream]
        res, o, cons, srow, flags, n, hi = d["keep"]
        if mapped and not count and d2h_events is None:
            # the packed files were written into the pinned host buffers by the pack kernels themselves: only the small
            # tables are left to copy, and nothing here needs the sizes (no host synchronisation per chunk)
            with torch.cuda.stream(st):
                for k, pl in enumerate(PLANOS):
                    h_off[pl][hi][:res[pl].off.numel()].copy_(res[pl].off, non_blocking=True)
                for k in range(3):
                    h_off["gz"][hi][k][:res["gz"][k].off.numel()].copy_(res["gz"][k].off, non_blocking=True)
                h_off["gz8"][hi][:res["gz8"].off.numel()].copy_(res["gz8"].off, non_blocking=True)
                h_status[hi][:srow].copy_(d["status"][:srow], non_blocking=True)
                h_inexact[hi].copy_(d["inexact"], non_blocking=True)
                n_status[hi] = srow
                d["done"].record(st)
            return
        tw = time.perf_counter()
        d["event"].synchronize()
        if inline:
            waits[0] += time.perf_counter() - tw
        tot = d["h_tot"].numpy()
        with torch.cuda.stream(st):
            moved = 0
            if d2h_events is not None:
                d2h_events.append([torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), 0])
                d2h_events[-1][0].record(st)
            for k, pl in enumerate(PLANOS):
                t = int(tot[k])
                if not mapped:
                    h_png[pl][hi][:t].copy_(res[pl].data[:t], non_blocking=True)
                h_off[pl][hi][:res[pl].off.numel()].copy_(res[pl].off, non_blocking=True)
                moved += t + res[pl].off.numel() * 8
            third = d["gz_out"].numel() // 3
            for k in range(3):
                t = int(tot[3 + k])
                if not mapped:
                    h_gz[hi][k * third:k * third + t].copy_(res["gz"][k].data[:t], non_blocking=True)
                h_off["gz"][hi][k][:res["gz"][k].off.numel()].copy_(res["gz"][k].off, non_blocking=True)
                moved += t + res["gz"][k].off.numel() * 8
            t = int(tot[6])
            if not mapped:
                h_gz8[hi][:t].copy_(res["gz8"].data[:t], non_blocking=True)
            h_off["gz8"][hi][:res["gz8"].off.numel()].copy_(res["gz8"].off, non_blocking=True)
            h_status[hi][:srow].copy_(d["status"][:srow], non_blocking=True)
            h_inexact[hi].copy_(d["inexact"], non_blocking=True)
            n_status[hi] = srow
            moved += t + res["gz8"].off.numel() * 8 + srow * 16 + 16
            d["done"].record(st)
            if d2h_events is not None:
                d2h_events[-1][1].record(st)
                d2h_events[-1][2] = moved
            if count:
                d2h += moved

    d2h_events = None        # profile pass: CUDA events around every chunk's device-to-host copies

    # A chunk's copies can only be enqueued once its sizes are on the host.  A helper thread waits for that (the wait releases
    # the GIL) and enqueues them, so that the main thread keeps enqueueing the next chunks: up to `nstream` chunks are in flight
    # and the latency-bound kernels of one chunk overlap the wide kernels of another.  MSL_E2E_INLINE=1: everything on one thread.
    import queue
    inline = os.environ.get("MSL_E2E_INLINE") is not None
    work_q = queue.Queue()
    pending = [None] * nstream          # per slot: threading.Event set when the slot's copies have been enqueued
    worker_err = []

    def worker():
        torch.cuda.set_device(device)
        while True:
            item = work_q.get()
            if item is None:
                return
            ci_, c0_, count_, ev_ = item
            try:
                deliver(ci_, c0_, count_)
            except BaseException as ex:      # noqa: BLE001 - reported by the main thread
                worker_err.append(ex)
            ev_.set()

    if not inline:
        threading.Thread(target=worker, daemon=True).start()

    def drain():
        for ev_ in pending:
            if ev_ is not None:
                ev_.wait()
        if worker_err:
            raise worker_err[0]

    def e2e_step(count=False):
        chunks = list(enumerate(range(0, B, CH)))
        if inline:
            for i, (ci, c0) in enumerate(chunks):
                compute(ci, c0, count)
                if i > 0:
                    deliver(*chunks[i - 1], count)
            deliver(*chunks[-1], count)
            return
        for ci, c0 in chunks:
            sl = ci % nstream
            if pending[sl] is not None:
                tw = time.perf_counter()
                pending[sl].wait()
                waits[0] += time.perf_counter() - tw
            compute(ci, c0, count)
            pending[sl] = threading.Event()
            work_q.put((ci, c0, count, pending[sl]))

    def sync_all():
        if not inline:
            drain()
        for st in streams:
            st.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e2e_step(); sync_all(); e2e_step(count=True)
    sync_all()
    K = max(3, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    sync_all()
    sec = (time.perf_counter() - t0) / K
    # per-kernel device time of one more (untimed) step, and the host time spent enqueueing it
    from mslesseg_b200 import _lib as _L2
    waits[0] = 0.0
    th0 = time.perf_counter()
    e2e_step()
    host_total_ms = (time.perf_counter() - th0) * 1e3
    host_enqueue_ms = host_total_ms - waits[0] * 1e3          # Python + launch work of one step (unprofiled)
    sync_all()
    _L2.profile_enable(True)
    d2h_events = []
    e2e_step()
    sync_all()
    d2h_ms = sum(a_.elapsed_time(b_) for a_, b_, _ in d2h_events)
    d2h_moved = sum(m_ for _, _, m_ in d2h_events)
    d2h_events = None
    tl = _L2.profile_timeline()
    kern = {k: round(v[0], 3) for k, v in sorted(_L2.profile_collect().items(), key=lambda kv: -kv[1][0])}
    if mapped:                # the result bytes cross the link inside the pack kernels
        d2h_ms = kern.get("deflate_pack", 0.0)
    d2h_gbs = d2h_moved / max(d2h_ms, 1e-9) / 1e6
    # union of the kernel intervals of that step = time during which at least one of this library's kernels was running
    busy, end = 0.0, -1.0
    for _, _, a_, b_ in sorted(tl, key=lambda r: r[2]):
        if b_ > end:
            busy += b_ - max(a_, end)
            end = b_
    span = (max(r[3] for r in tl) - min(r[2] for r in tl)) if tl else 0.0
    if os.environ.get("MSL_BENCH_TIMELINE"):
        with open(os.environ["MSL_BENCH_TIMELINE"], "w") as f:
            json.dump(tl, f)
    # ---- the host files decode (Pillow / gzip on the host = checker) to what the device holds for the last chunks
    import gzip as _gz
    import io as _io
    from PIL import Image
    ok = True
    for ci in range(max(0, nchunks - min(nstream, nhost)), nchunks):
        c0 = ci * CH
        d = slots[ci % nstream]
        res, o, cons, srow, flags, n, hi = d["keep"]
        ok = ok and bool((h_status[hi].numpy()[:n_status[hi], 0] == 0).all()) and int(h_inexact[hi].numpy().sum()) == 0
        for pl in PLANOS:
            n_p, rows, cols = dims[pl]
            off = h_off[pl][hi].numpy()
            buf = h_png[pl][hi].numpy()
            for k, m in enumerate(MEJORAS):
                for (v, i) in ((0, n_p // 2), (n - 1, n_p // 3)):
                    j = (k * n + v) * n_p + i if n != CH else (k * CH + v) * n_p + i
                    im = np.array(Image.open(_io.BytesIO(buf[off[j]:off[j + 1]].tobytes())))
                    ok = ok and bool(np.array_equal(im, o[(m, pl)][v, i].cpu().numpy()))
        third = d["gz_out"].numel() // 3
        for k, pl in enumerate(PLANOS):
            off = h_off["gz"][hi][k].numpy()
            raw = _gz.decompress(h_gz[hi][k * third + int(off[0]):k * third + int(off[gz_spv])].numpy().tobytes())
            ok = ok and len(raw) == RAWF and bool(np.array_equal(np.frombuffer(raw, "<f4", offset=352), d["rvol"][pl][0].reshape(-1).float().cpu().numpy()))
        off = h_off["gz8"][hi].numpy()
        raw = _gz.decompress(h_gz8[hi][int(off[0]):int(off[gz_spv8])].numpy().tobytes())
        ok = ok and bool(np.array_equal(np.frombuffer(raw, np.uint8, offset=352), cons[0].reshape(-1).cpu().numpy()))
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    raw_in = B * (4 * N + 4 * N) + sum(int(preds[pl][0].numel()) for pl in PLANOS)
    raw_out = B * (12 * N + 3 * 4 * N + N)
    return {"value": world * B * N_VOX / sec / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": sec * 1e3, "steps": K, "host_results_equal_device": bool(ok),
            "uncompressed_bytes_per_step": {"inputs": int(raw_in), "results": int(raw_out)},
            "kernel_ms_per_step": kern, "d2h_copy_ms_per_step": round(d2h_ms, 3), "d2h_copy_gb_s": round(d2h_gbs, 1), "kernel_busy_ms_per_step": round(busy, 3), "kernel_span_ms_per_step": round(span, 3), "host_enqueue_ms_per_step": host_enqueue_ms,
            "note": ("host buffers hold FILES: .nii.gz volumes (FLAIR float32, GT float32) and predicted-mask PNGs in; PNG slices of the 12 "
                     "stacks, float32 .nii.gz of the 3 reconstructions, uint8 consensus .nii.gz and the count table out; inflate / "
                     f"deflate on the GPU; {CH}-patient chunks over {nstream} streams, a helper thread enqueues a chunk's copies once its sizes are on the host; wall clock around synchronised steps; "
                     + ("the packed files are written into the pinned host buffers by the pack kernels (mapped memory), d2h_copy_* = those kernels"
                        if mapped else "packed on the device, one copy per stack once its size is on the host"))}


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was
    redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # fd 1 -> stderr for the rest of the run
    if args.impl == "reference":
        return run_reference(args)
    if args.config != "default":
        import bench_configs as BC
        if args.config == "dropin":
            return BC.run_dropin(args, emit, ClockSampler, METRIC, UNIT)
        return BC.run_cohort(args, args.config, emit, ClockSampler, METRIC, UNIT)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
