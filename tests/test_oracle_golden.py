"""CPU: the oracle (oracle/oracle.py) against the golden vectors frozen from the REAL
reference by oracle/make_golden.py, and against cv2 where it is importable."""
import hashlib
import math
import warnings

import numpy as np
import pytest

from oracle import oracle as O
from mslesseg_b200 import synthetic as S

PLANOS = O.PLANOS
MEJORAS = O.MEJORAS


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def enhance_all(vol_xyz, plano, mejora):
    n = vol_xyz.shape[O.plane_axis(plano)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.stack([O.enhance_slice(O.slice_of(vol_xyz, plano, i), mejora) for i in range(n)])


def test_demo_slices_match_reference(demo_slices):
    keys = [k for k in demo_slices.files if k.endswith("_raw")]
    assert len(keys) == 12
    for k in keys:
        raw = demo_slices[k].astype(np.float64)
        for mej in MEJORAS:
            want = demo_slices[k[:-4] + "_" + mej]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                got = O.enhance_slice(raw, mej)
            assert np.array_equal(got, want), (k, mej)


def test_blank_slice_values():
    # blank slices give HE 0, CLAHE 4, GC 0, LT 0 for the three slice shapes
    for shape in ((182, 218), (182, 182), (218, 182)):
        z = np.zeros(shape, dtype=np.float64)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert np.all(O.enhance_slice(z, "HE") == 0)
            assert np.all(O.enhance_slice(z, "CLAHE") == 4)
            assert np.all(O.enhance_slice(z, "GC") == 0)
            assert np.all(O.enhance_slice(z, "LT") == 0)


@pytest.mark.parametrize("pid", ["P1", "P2"])
def test_synthetic_whole_volume_digests(golden, pid):
    g = golden["synthetic_enhance"][pid]
    pat = S.make_patient(int(pid[1:]), config_id=1, num_cortes=20)
    assert pat.seed == g["seed"]
    assert sha(pat.flair) == g["flair_sha"], "synthetic generator drifted (numpy version?)"
    assert sha(pat.gt) == g["gt_sha"]
    v = S.as_xyz(pat.flair).astype(np.float64)
    gt = S.as_xyz(pat.gt)
    for plano in PLANOS:
        pe = g["planes"][plano]
        lesion = O.indices_cortes_con_lesion(gt, plano)
        assert len(lesion) == pe["n_lesion"]
        assert sha(np.asarray(lesion, dtype=np.int32)) == pe["lesion_sha"]
        assert O.indices_a_usar(gt, plano, 20) == pe["usar20"]
        assert O.indices_a_usar(gt, plano, 7) == pe["usar7"]
        idx7 = O.indices_a_usar(gt, plano, 7)
        gc7 = np.stack([O.enhance_slice(O.slice_of(v, plano, i), "GC") for i in idx7])
        assert sha(gc7) == pe["cortes_img_GC7_sha"]
        m7 = np.stack([O.slice_of(gt, plano, i) for i in idx7]).astype(np.uint8)
        assert sha(m7) == pe["cortes_mask7_sha"]
        for mej in MEJORAS:
            assert sha(enhance_all(v, plano, mej)) == pe[mej], (plano, mej)


def test_noise_volume_digests(golden):
    from oracle.make_golden import noise_volume
    g = golden["noise_enhance"]
    nv = noise_volume(g["seed"], tuple(g["shape_xyz"]))
    assert sha(nv) == g["in_sha"]
    v = S.as_xyz(nv).astype(np.float64)
    for plano in PLANOS:
        for mej in MEJORAS:
            assert sha(enhance_all(v, plano, mej)) == g["planes"][plano][mej], (plano, mej)


def test_demo_lesion_known_answers(golden):
    # SURVEY section 8c known answers, re-derived by make_golden.py through the reference's Paciente
    d = golden["demo"]
    assert [d["P18"][p]["n_lesion"] for p in PLANOS] == [28, 21, 23]
    assert [d["P39"][p]["n_lesion"] for p in PLANOS] == [101, 147, 113]
    assert (d["P39"]["axial"]["usar20"][0], d["P39"]["axial"]["usar20"][-1]) == (74, 93)


@pytest.mark.parametrize("pid", ["P54", "P55", "P56"])
def test_output_side_against_reference(golden, pid):
    g = golden["synthetic_eval"][pid]
    pat = S.make_patient(int(pid[1:]), config_id=2, num_cortes=20)
    assert sha(pat.gt) == g["gt_sha"]
    gt = S.as_xyz(pat.gt).astype(np.float64)
    vols = {}
    for plano in PLANOS:
        pe = g["planes"][plano]
        assert pat.pred_indices[plano] == pe["indices"]
        assert sha(pat.pred_slices[plano]) == pe["slices_sha"]
        vol = O.reconstruir(pat.pred_slices[plano], pat.pred_indices[plano], S.SHAPE_XYZ, plano)
        assert vol.dtype == np.float32
        assert sha(np.ascontiguousarray(vol.transpose(2, 1, 0))) == pe["recon_f32_sha"]
        assert sha(np.ascontiguousarray(vol.transpose(2, 1, 0)).astype(np.uint8)) == pe["recon_u8_sha"]
        vols[plano] = vol.astype(np.float64)
        counts = O.confusion_counts(gt, vols[plano])
        assert list(counts) == pe["counts"]
        assert O.metricas_desde_conteos(*counts) == pe["metricas"]
    for umbral in (2, 3):
        ce = g[f"consenso{umbral}"]
        c = O.combinar_volumenes(vols["axial"], vols["coronal"], vols["sagital"], umbral)
        assert c.dtype == np.uint8
        assert sha(np.ascontiguousarray(c.transpose(2, 1, 0))) == ce["sha"]
        counts = O.confusion_counts(gt, c)
        assert list(counts) == ce["counts"]
        assert O.metricas_desde_conteos(*counts) == ce["metricas"]


def test_literal_metrics_small(golden):
    # full-array restatement == counts restatement == reference (random small cases)
    for case in golden["random_metricas"]:
        tp, fp, fn, tn = case["counts"]
        yt = np.r_[np.ones(tp + fn), np.zeros(fp + tn)]
        yp = np.r_[np.ones(tp), np.zeros(fn), np.ones(fp), np.zeros(tn)]
        assert O.generar_diccionario_metricas(yt, yp) == case["metricas"]
        assert O.metricas_desde_conteos(tp, fp, fn, tn) == case["metricas"]


def test_edge_metrics(golden):
    def norm(d):
        return {k: (None if (isinstance(v, float) and math.isnan(v)) else v) for k, v in d.items()}
    e = golden["edge_metricas"]
    assert norm(O.metricas_desde_conteos(0, 1, 0, 119)) == e["empty_gt"]
    assert norm(O.metricas_desde_conteos(0, 0, 2, 118)) == e["empty_pred"]
    assert norm(O.metricas_desde_conteos(2, 0, 0, 118)) == e["perfect"]


def test_auc_restatement_vs_sklearn():
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(11)
    for _ in range(200):
        tp, fp, fn, tn = (int(x) for x in rng.integers(0, 5000, 4))
        if tp + fn == 0 or fp + tn == 0:
            continue
        w = np.array([tn, fp, fn, tp], dtype=np.float64)
        keep = w > 0
        want = sk.roc_auc_score(np.array([0, 0, 1, 1])[keep], np.array([0, 1, 0, 1])[keep], sample_weight=w[keep])
        assert O.auc_binary_from_counts(tp, fp, fn, tn) == want


def test_fold_statistics(golden):
    assert O.calcular_promedio(golden["promedio_fold"]["in"]) == golden["promedio_fold"]["out"]
    assert O.calcular_resumen_experimento(golden["resumen_experimento"]["in"]) == golden["resumen_experimento"]["out"]
    for k, table in golden["calcular_fold"].items():
        for pid, fold in table.items():
            assert O.calcular_fold(pid, int(k)) == fold
    with pytest.raises(ValueError):
        O.calcular_fold("P54", 5)
    assert O.calcular_fold("P75", 5, n_ids=75) == 5
    assert O.num_cortes_percentil(golden["percentil"]["in"], 50) == golden["percentil"]["P50"]
    assert O.num_cortes_percentil(golden["percentil"]["in"], 25) == golden["percentil"]["P25"]


def test_tables_and_cv2_agreement():
    cv2 = pytest.importorskip("cv2")
    g = np.arange(256, dtype=np.uint8).reshape(16, 16)
    lut_l = cv2.cvtColor(cv2.cvtColor(g, cv2.COLOR_GRAY2BGR), cv2.COLOR_BGR2LAB)[..., 0].ravel()
    assert np.array_equal(lut_l, O.LUT_L)
    lab = np.stack([g, np.full_like(g, 128), np.full_like(g, 128)], axis=-1)
    lut_out = cv2.cvtColor(cv2.cvtColor(lab, cv2.COLOR_LAB2BGR), cv2.COLOR_BGR2GRAY).ravel()
    assert np.array_equal(lut_out, O.LUT_OUT)
    rng = np.random.default_rng(3)
    for shape in ((182, 218), (218, 182), (182, 182), (64, 64), (37, 91), (8, 8)):
        for kind in range(3):
            if kind == 0:
                u = rng.integers(0, 256, shape, dtype=np.uint8)
            elif kind == 1:
                u = (rng.integers(0, 6, shape) * 50).astype(np.uint8)
            else:
                u = np.full(shape, 77, np.uint8)
            assert np.array_equal(O.equalize_hist(u), cv2.equalizeHist(u))
            assert np.array_equal(O.clahe_apply(u), cv2.createCLAHE(2.0, (8, 8)).apply(u))


def test_lt_table_is_slice_independent():
    # E1 maps the slice maximum to exactly 255 whenever ptp > 0, so LT collapses to one table
    t = O.lt_table(255)
    assert t[0] == 0 and t[255] == 255 and t[1] == 31
    assert O.lt_table(0)[0] == 0          # blank slice: inf*0 = NaN -> 0 (only entry 0 is reachable)
    assert O.gc_table()[16] == 1 and O.gc_table()[255] == 255


def test_prediction_postprocessing_golden():
    """R0: combinar_predicciones / normalizar_prediccion restated vs the reference's frozen outputs."""
    import json, os
    from oracle.make_golden_pred import CASES, instance_masks, sha as sha_
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_pred_v1.json")))
    assert len(g["cases"]) == len(CASES)
    for case, (seed, n, mh, mw, h, w) in zip(g["cases"], CASES):
        masks = instance_masks(seed, n, mh, mw)
        comb = O.combinar_predicciones(list(masks), (h, w))
        assert sha_(comb) == case["combined_sha"] and int(comb.sum()) == case["combined_sum"], seed
        norm = O.normalizar_prediccion(comb)
        assert sha_(norm) == case["normalised_sha"] and list(norm.shape) == case["normalised_shape"], seed


def test_resize_nearest_index_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for _ in range(60):
        sh, sw, dh, dw = (int(v) for v in rng.integers(1, 400, 4))
        a = rng.integers(0, 2, (sh, sw)).astype(np.uint8)
        want = cv2.resize(a, (dw, dh), interpolation=cv2.INTER_NEAREST)
        got = a[O.resize_nearest_index(dh, sh)][:, O.resize_nearest_index(dw, sw)]
        assert np.array_equal(got, want), (sh, sw, dh, dw)


def test_png_container_oracle_decodes():
    """The stored-deflate PNG container restated in oracle.png_stored is a valid PNG: zlib and Pillow / cv2 read it back."""
    import io, struct, zlib
    rng = np.random.default_rng(9)
    for shape in [(218, 182, 4), (5, 7, 4), (33, 21), (300, 200, 4)]:
        a = rng.integers(0, 256, shape, dtype=np.uint8)
        b = O.png_stored(a)
        assert b[:8] == b"\x89PNG\r\n\x1a\n"
        # walk the chunks, check every CRC, inflate the IDAT
        off, idat = 8, b""
        while off < len(b):
            n, tag = struct.unpack(">I4s", b[off:off + 8])
            data = b[off + 8:off + 8 + n]
            assert struct.unpack(">I", b[off + 8 + n:off + 12 + n])[0] == zlib.crc32(tag + data) & 0xFFFFFFFF
            if tag == b"IDAT":
                idat += data
            off += 12 + n
        H, W = shape[:2]
        ch = shape[2] if len(shape) == 3 else 1
        raw = zlib.decompress(idat)
        assert len(raw) == H * (1 + W * ch)
        rows = np.frombuffer(raw, np.uint8).reshape(H, 1 + W * ch)
        assert int(rows[:, 0].sum()) == 0 and np.array_equal(rows[:, 1:].reshape(a.shape), a)
        cv2 = pytest.importorskip("cv2")
        d = cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(d[..., [2, 1, 0, 3]] if a.ndim == 3 else d, a)


def test_label_goldens_reproduce_with_installed_cv2():
    """The frozen label text digests of the demo masks (oracle/make_golden_labels.py) are reproduced by the restated
    ultralytics converter on this machine's cv2 - so a cv2 upgrade that changes findContours is noticed on the CPU."""
    import json
    pytest.importorskip("cv2")
    from conftest import GOLDEN_DIR
    from oracle import ref_stubs
    z = np.load(GOLDEN_DIR / "demo_label_masks.npz")
    gold = json.loads((GOLDEN_DIR / "demo_labels_v1.json").read_text())
    for key in ("P18_axial", "P39_sagital"):
        shape = tuple(int(d) for d in z[key + "_shape"])
        masks = np.unpackbits(z[key + "_bits"])[:int(np.prod(shape))].reshape(shape)
        for i, m in enumerate(masks):
            text = "".join(ln + "\n" for ln in ref_stubs.yolo_seg_lines(m, 1))
            assert hashlib.sha256(text.encode()).hexdigest() == gold[key]["sha"][i]
