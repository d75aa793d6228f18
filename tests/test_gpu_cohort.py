"""GPU: cohort-sized runs (BASELINE.json configs[2] / configs[3] shapes) - recon -> consensus -> eval for a 22-volume
synthetic test cohort with "P50" slice selection, checked against the oracle patient by patient, plus the
size-independent properties of the output side on a full batch."""
import numpy as np
import pytest

from oracle import oracle as O
from mslesseg_b200 import synthetic as S

pytestmark = pytest.mark.gpu
PLANOS = O.PLANOS


@pytest.fixture(scope="module")
def env(cuda_device):
    import torch
    from mslesseg_b200 import _lib, ops, metrics, dist
    _lib.load()
    return torch, ops, metrics, dist, cuda_device


def test_22_volume_cohort_p50(env):
    torch, ops, M, D, dev = env
    ids = [f"P{n}" for n in range(54, 76)]                      # the 22 ids of the MSLesSeg test split
    base = [S.make_patient(int(p[1:]), config_id=2, num_cortes=None, with_predictions=False) for p in ids]
    gt = torch.from_numpy(np.stack([p.gt for p in base])).to(dev)
    flags = [f.cpu().numpy() for f in ops.lesion_slices(gt)]
    # num_cortes = "P50": percentile of the per-patient lesion-slice counts (scripts/extraer_dataset.py:110-135)
    num_cortes = {}
    for k, plano in enumerate(PLANOS):
        counts = [int(flags[k][v].sum()) for v in range(len(ids))]
        assert counts == [len(O.indices_cortes_con_lesion(S.as_xyz(p.gt), plano)) for p in base[:3]] + counts[3:]
        num_cortes[plano] = M.num_cortes_percentil(counts, 50)
        assert num_cortes[plano] == O.num_cortes_percentil(counts, 50)
    # predictions exist only on the selected slices of every plane
    pats = []
    for p in base:
        q = S.make_patient(int(p.id[1:]), config_id=2, num_cortes=None)
        for k, plano in enumerate(PLANOS):
            keep = M.ventana_central(q.pred_indices[plano], num_cortes[plano])
            sel = [q.pred_indices[plano].index(i) for i in keep]
            q.pred_indices[plano] = keep
            q.pred_slices[plano] = q.pred_slices[plano][sel]
        pats.append(q)
    vols = {}
    for plano in PLANOS:
        sl = torch.from_numpy(np.concatenate([p.pred_slices[plano] for p in pats])).to(dev)
        vs = np.concatenate([np.full(len(p.pred_indices[plano]), v, np.int32) for v, p in enumerate(pats)])
        ix = np.concatenate([np.asarray(p.pred_indices[plano], np.int32) for p in pats])
        vols[plano] = ops.recon(sl, vs, ix, plano, len(pats), S.SHAPE_XYZ)
    cons, counts = ops.consensus_eval(vols["axial"], vols["coronal"], vols["sagital"], gt, 2)
    counts = counts.cpu().numpy()
    assert np.all(counts.sum(axis=2) == 182 * 218 * 182)        # binary inputs: tp + fp + fn + tn == N for every plane
    # oracle, patient by patient
    per_patient = {}
    for v, p in enumerate(pats):
        gtx = S.as_xyz(p.gt)
        ref_vols = [O.reconstruir(p.pred_slices[pl], p.pred_indices[pl], S.SHAPE_XYZ, pl) for pl in PLANOS]
        ref_cons = O.combinar_volumenes(*(r.astype(np.float64) for r in ref_vols), 2)
        for k, r in enumerate(ref_vols + [ref_cons]):
            assert counts[v, k].tolist() == list(O.confusion_counts(gtx, r)), (p.id, k)
        assert np.array_equal(S.as_xyz(cons[v].cpu().numpy()), ref_cons)
        per_patient[p.id] = O.metricas_desde_conteos(*O.confusion_counts(gtx, ref_cons))
    # sharded table path (single process) + per-patient metrics + the reference's mean / std (ddof=0) over the cohort
    table = D.all_reduce_count_table(ids, ids, torch.from_numpy(counts))
    got = D.metrics_from_table(ids, table)
    assert {p: got[p]["consenso"] for p in ids} == per_patient
    acc = {}
    for p in ids:
        for k_, val in per_patient[p].items():
            acc.setdefault(k_, []).append(val)
    assert M.calcular_promedio(acc) == O.calcular_promedio(acc)


def test_output_side_properties_full_batch(env):
    torch, ops, M, D, dev = env
    g = torch.Generator(device="cpu").manual_seed(3)
    a = (torch.rand((6, 182, 218, 182), generator=g) < 0.02).to(torch.uint8).to(dev)
    b = (torch.rand((6, 182, 218, 182), generator=g) < 0.02).to(torch.uint8).to(dev)
    # idempotence: voting a volume with itself returns it, for majority and unanimity; counts of (a, a) are diagonal
    for umbral in (2, 3):
        cons, counts = ops.consensus_eval(a, a, a, a, umbral)
        assert torch.equal(cons, a)
        c = counts.cpu().numpy()
        assert np.all(c[:, :, 1] == 0) and np.all(c[:, :, 2] == 0)
        assert np.array_equal(c[:, 0, 0], a.flatten(1).sum(1).cpu().numpy())
    # majority of (a, b, 0) == a & b ; unanimity of (a, b, 1) == a & b ; majority of (a, b, 1) == a | b
    zero, one = torch.zeros_like(a), torch.ones_like(a)
    assert torch.equal(ops.consensus_eval(a, b, zero, None, 2)[0], a & b)
    assert torch.equal(ops.consensus_eval(a, b, one, None, 3)[0], a & b)
    assert torch.equal(ops.consensus_eval(a, b, one, None, 2)[0], a | b)
    # symmetry of the confusion table: swapping the roles swaps fp and fn
    c1 = ops.confusion_counts(a, b).cpu().numpy()
    c2 = ops.confusion_counts(b, a).cpu().numpy()
    assert np.array_equal(c1[:, [0, 2, 1, 3]], c2)
    assert np.all(c1.sum(axis=1) == 182 * 218 * 182)
    # recon of every slice of a volume reproduces the volume (encode -> decode round trip) for the three planes
    for plano in PLANOS:
        n_p, rows, cols = ops.plane_dims(plano, 182, 218, 182)
        stack = ops.enhance_slices(a[:1] * 255, None, plano, layout="G")        # uint8 volume -> its slices (x255 like the PNGs)
        back = ops.recon(stack, [0] * n_p, list(range(n_p)), plano, 1, S.SHAPE_XYZ)
        assert torch.equal(back, a[:1]), plano
