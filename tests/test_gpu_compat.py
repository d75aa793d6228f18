"""GPU: the drop-in shims (mslesseg_b200.compat) against the golden vectors of the real reference / the oracle."""
import hashlib
import warnings

import numpy as np
import pytest

from oracle import oracle as O
from mslesseg_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def C(cuda_device):
    from mslesseg_b200 import _lib
    _lib.load()
    from mslesseg_b200 import compat
    from mslesseg_b200.compat import mejora_imagen, utils, Paciente, reconstruir_volumen, generar_consenso, eval, extraer_dataset
    return compat


def test_algoritmo_classes_on_demo_slices(C, demo_slices):
    MI, U = C.mejora_imagen, C.utils
    for k in [k for k in demo_slices.files if k.endswith("_raw")][:6]:
        raw = demo_slices[k].astype(np.float64)
        for cls in (MI.HE, MI.CLAHE, MI.GC, MI.LT):
            out = cls().aplicar(raw)
            assert out.shape == raw.shape + (3,) and out.dtype == np.uint8
            assert np.array_equal(U.verificar_grises(out), demo_slices[k[:-4] + "_" + repr(cls())]), (k, repr(cls()))
            assert np.array_equal(cls().aplicar_gris(raw), demo_slices[k[:-4] + "_" + repr(cls())])
        assert np.array_equal(U.normalizar_a_uint8(raw), O.normalizar_a_uint8(raw))
    with pytest.raises(NotImplementedError):
        MI.Algoritmo().aplicar(np.zeros((4, 4)))
    with pytest.raises(ValueError):
        MI.CLAHE(clip_limit=3.0)


def test_clahe_three_channels_match_cv2(C, demo_slices):
    cv2 = pytest.importorskip("cv2")
    raw = demo_slices["P39_axial_84_raw"].astype(np.float64)
    u = O.normalizar_a_uint8(raw)
    lab = cv2.cvtColor(cv2.cvtColor(u, cv2.COLOR_GRAY2BGR), cv2.COLOR_BGR2LAB)
    l, a, b = cv2.split(lab)
    want = cv2.cvtColor(cv2.merge((cv2.createCLAHE(2.0, (8, 8)).apply(l), a, b)), cv2.COLOR_LAB2BGR)
    assert np.array_equal(C.mejora_imagen.CLAHE().aplicar(raw), want)


def test_paciente_shim_against_reference(C, golden):
    P = C.Paciente.Paciente
    g = golden["synthetic_enhance"]["P1"]
    pat = S.make_patient(1, config_id=1, num_cortes=20)
    vol = np.asfortranarray(S.as_xyz(pat.flair).astype(np.float64))
    gt = np.asfortranarray(S.as_xyz(pat.gt).astype(np.float64))
    for plano in O.PLANOS:
        pe = g["planes"][plano]
        p = P("P1", plano, modalidad=["FLAIR"], mejora="GC", gt_mask=gt)
        p._volumenes["FLAIR"] = vol
        lesion = p.indices_cortes_con_lesion()
        assert len(lesion) == pe["n_lesion"] and sha(np.asarray(lesion, dtype=np.int32)) == pe["lesion_sha"]
        assert p.indices_a_usar(20) == pe["usar20"] and p.indices_a_usar(7) == pe["usar7"]
        cl = p.cortes_con_lesion_img(7)["FLAIR"]
        assert [i for i, _ in cl] == pe["usar7"] and cl[0][1].shape[2] == 3
        assert sha(np.stack([C.utils.verificar_grises(c) for _, c in cl])) == pe["cortes_img_GC7_sha"]
        cm = p.cortes_con_lesion_mask(7)
        assert sha(np.stack([c for _, c in cm]).astype(np.uint8)) == pe["cortes_mask7_sha"]
        idx, st = p.cortes_con_lesion_gris(7)["FLAIR"]
        assert sha(st) == pe["cortes_img_GC7_sha"]
    # no enhancement: the raw float64 slices, like the reference
    p = P("P1", "axial", modalidad=["FLAIR"], mejora=None, gt_mask=gt)
    p._volumenes["FLAIR"] = vol
    i, s = p.cortes_con_lesion_img(3)["FLAIR"][0]
    assert s.dtype == np.float64 and np.array_equal(s, vol[:, :, i])


def test_output_side_shims_against_reference(C, golden, tmp_path):
    from PIL import Image
    pid = "P54"
    ge = golden["synthetic_eval"][pid]
    pat = S.make_patient(54, config_id=2, num_cortes=20)
    gt = np.asfortranarray(S.as_xyz(pat.gt).astype(np.float64))
    aff = np.diag([1.0, 1.0, 1.0, 1.0])
    gt_path = tmp_path / "GT" / f"{pid}_MASK.nii.gz"
    C.utils.guardar_volumen(gt.astype(np.float32), aff, gt_path)
    vols = {}
    for plano in O.PLANOS:
        d = tmp_path / plano
        d.mkdir()
        for i, q in zip(pat.pred_indices[plano], pat.pred_slices[plano]):
            Image.fromarray(q).save(d / f"{pid}_FLAIR_{i}.png")
        out = tmp_path / f"{pid}_{plano}.nii.gz"
        vol = C.reconstruir_volumen.reconstruir_volumen(d, gt_path, out, plano)
        assert vol.dtype == np.float32 and vol.shape == S.SHAPE_XYZ
        assert sha(np.ascontiguousarray(vol.transpose(2, 1, 0))) == ge["planes"][plano]["recon_f32_sha"]
        assert C.utils.reconstruccion_valida(out, gt_path)
        vols[plano] = C.utils.cargar_volumen(out)
        assert C.eval.generar_diccionario_metricas(gt, vols[plano]) == ge["planes"][plano]["metricas"]
        assert C.eval.calcular_metricas(gt_path, out) == ge["planes"][plano]["metricas"]
    for umbral in (2, 3):
        c = C.generar_consenso.combinar_volumenes(vols["axial"], vols["coronal"], vols["sagital"], umbral)
        assert c.dtype == np.uint8 and c.shape == S.SHAPE_XYZ
        assert sha(np.ascontiguousarray(c.transpose(2, 1, 0))) == ge[f"consenso{umbral}"]["sha"]
        assert C.eval.generar_diccionario_metricas(gt, c) == ge[f"consenso{umbral}"]["metricas"]
    cons_path = tmp_path / f"{pid}_consenso.nii.gz"
    C.generar_consenso.generar_consenso(tmp_path / f"{pid}_axial.nii.gz", tmp_path / f"{pid}_coronal.nii.gz",
                                        tmp_path / f"{pid}_sagital.nii.gz", cons_path, umbral=2)
    assert C.eval.calcular_metricas(gt_path, cons_path) == ge["consenso2"]["metricas"]
    # error behaviour of the reference: out-of-range index / wrong slice shape -> ValueError
    bad = tmp_path / "bad"; bad.mkdir()
    Image.fromarray(np.zeros((182, 218), np.uint8)).save(bad / f"{pid}_FLAIR_999.png")
    with pytest.raises(ValueError, match="fuera de rango"):
        C.reconstruir_volumen.reconstruir_volumen(bad, gt_path, tmp_path / "x.nii.gz", "axial")
    with pytest.raises(ValueError, match="incorrectas"):
        C.reconstruir_volumen.reconstruir_volumen(tmp_path / "axial", gt_path, tmp_path / "x.nii.gz", "sagital")
    with pytest.raises(ValueError):
        C.eval.generar_diccionario_metricas(gt, vols["axial"] * 0.5)       # non-integral "mask"
    with pytest.raises(ValueError):
        C.eval.generar_diccionario_metricas(gt * 3, vols["axial"])         # non-binary mask
    assert np.isnan(C.utils.AUC(np.zeros((4, 4)), np.zeros((4, 4))))        # single-class GT -> nan, like the reference


def test_guardar_cortes_writes_imsave_pixels(C, tmp_path):
    from PIL import Image
    pat = S.make_patient(2, config_id=1, num_cortes=5)
    vol = np.asfortranarray(S.as_xyz(pat.flair).astype(np.float64))
    gt = np.asfortranarray(S.as_xyz(pat.gt).astype(np.float64))
    for plano, mejora in (("axial", "HE"), ("sagital", "CLAHE"), ("coronal", None)):
        p = C.Paciente.Paciente("P2", plano, modalidad=["FLAIR"], mejora=mejora, gt_mask=gt)
        p._volumenes["FLAIR"] = vol
        imgs, masks = tmp_path / f"{plano}_images", tmp_path / f"{plano}_masks"
        imgs.mkdir(); masks.mkdir()
        C.extraer_dataset.guardar_cortes(p, imgs, masks, 5)
        idx = p.indices_a_usar(5)
        assert len(list(imgs.glob("*.png"))) == len(idx) == len(list(masks.glob("*.png")))
        for i in idx:
            s = O.slice_of(vol, plano, i)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want = O.imsave_rgba(O.enhance_slice(s, mejora) if mejora else s)
            assert np.array_equal(np.array(Image.open(imgs / f"P2_FLAIR_{i}.png")), want), (plano, mejora, i)
            assert np.array_equal(np.array(Image.open(masks / f"P2_{i}.png")), O.imsave_rgba(O.slice_of(gt, plano, i)))


@pytest.mark.gpu
def test_generar_predicciones_shim(C):
    """compat.generar_predicciones mirrors the reference names (scripts/generar_predicciones.py:123-140)."""
    from mslesseg_b200.compat import generar_predicciones as GP
    from oracle.make_golden_pred import instance_masks
    masks = instance_masks(21, 3, 160, 136, )
    comb = GP.combinar_predicciones(list(masks), (218, 182))
    assert comb.dtype == np.uint8 and comb.shape == (218, 182)
    assert np.array_equal(comb, O.combinar_predicciones(list(masks), (218, 182)))
    assert np.array_equal(GP.normalizar_prediccion(comb), O.normalizar_prediccion(comb))
    assert int(GP.combinar_predicciones([], (218, 182)).sum()) == 0


@pytest.mark.gpu
def test_verificar_grises_matches_cv2(C):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (40, 50, 3), dtype=np.uint8)
    assert np.array_equal(C.utils.verificar_grises(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    img[..., 1] = 255; img[..., 0] = 255; img[..., 2] = 255
    assert int(C.utils.verificar_grises(img).min()) == 255
