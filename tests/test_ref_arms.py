"""CPU: the two CPU-baseline arms bench.py can time - oracle/ref_path.py (the port) and oracle/ref_real.py (the real
reference functions from the staged copy) - are pinned to the golden vectors of the real reference and to each other."""
import warnings

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_path as RP
from oracle import ref_real as RR
from mslesseg_b200 import synthetic as S


def test_ref_path_enhancements_equal_reference_goldens(demo_slices):
    pytest.importorskip("cv2")
    keys = [k for k in demo_slices.files if k.endswith("_raw")]
    for k in keys:
        raw = demo_slices[k].astype(np.float64)
        for mej in O.MEJORAS:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                got = RP.verificar_grises(RP.aplicar_mejora(raw, mej))
            assert np.array_equal(got, demo_slices[k[:-4] + "_" + mej]), (k, mej)


def test_ref_path_output_side_equals_reference_goldens(golden):
    g = golden["cohort22"]["patients"][0] if "cohort22" in golden and "patients" in golden["cohort22"] else None
    pat = S.make_patient(54, config_id=2, num_cortes=20)
    gt = S.as_xyz(pat.gt).astype(np.float64)
    vols = [RP.reconstruir(pat.pred_slices[pl], pat.pred_indices[pl], S.SHAPE_XYZ, pl).astype(np.float64) for pl in O.PLANOS]
    cons = RP.combinar_volumenes(*vols, 2)
    assert np.array_equal(cons, O.combinar_volumenes(*vols, 2))
    for v in vols + [cons.astype(np.float64)]:
        assert RP.generar_diccionario_metricas(gt, v) == O.generar_diccionario_metricas(gt, v)
    if g is not None and "metrics" in g:
        assert RP.generar_diccionario_metricas(gt, cons.astype(np.float64)) == g["metrics"]["consenso"]


@pytest.mark.skipif(not RR.available(), reason="reference not staged")
def test_ref_real_equals_port_and_oracle():
    pat = S.make_patient(3, config_id=1, num_cortes=12)
    vol = S.as_xyz(pat.flair).astype(np.float64)
    gt = S.as_xyz(pat.gt).astype(np.float64)
    for plano in O.PLANOS:
        n = vol.shape[O.plane_axis(plano)]
        idx = [0, n // 3, n // 2, n - 1]
        for mej in O.MEJORAS:
            a = RR.enhance_plane(vol, plano, mej, idx)
            b = RP.enhance_plane(vol, plano, mej, idx)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                c = [O.png_orient(O.enhance_slice(O.slice_of(vol, plano, i), mej)) for i in idx]
            for x, y, z in zip(a, b, c):
                assert np.array_equal(x, y) and np.array_equal(x, z), (plano, mej)
    vols = []
    for plano in O.PLANOS:
        r = RR.reconstruir(pat.pred_slices[plano], pat.pred_indices[plano], S.SHAPE_XYZ, plano)
        assert r.dtype == np.float32
        assert np.array_equal(r, RP.reconstruir(pat.pred_slices[plano], pat.pred_indices[plano], S.SHAPE_XYZ, plano))
        vols.append(r.astype(np.float64))
    cons = RR.combinar_volumenes(*vols, 2)
    assert cons.dtype == np.uint8 and np.array_equal(cons, RP.combinar_volumenes(*vols, 2))
    assert RR.generar_diccionario_metricas(gt, cons.astype(np.float64)) == RP.generar_diccionario_metricas(gt, cons.astype(np.float64))


@pytest.mark.skipif(not RR.available(), reason="reference not staged")
def test_cpu_baseline_uses_the_real_reference():
    from oracle import cpu_baseline as CB
    assert CB.KIND == "reference" and CB.R is RR
