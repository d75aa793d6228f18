"""CPU: host-side pieces of the drop-in shims (no kernel launch)."""
import numpy as np
import pytest

from mslesseg_b200 import nifti
from mslesseg_b200.compat import utils as U
from mslesseg_b200.compat import reconstruir_volumen as RV
from mslesseg_b200.compat.Paciente import Paciente


def test_gray_inputs_pass_through_on_the_host():
    rng = np.random.default_rng(1)
    g = rng.integers(0, 256, (8, 9), dtype=np.uint8)
    assert U.verificar_grises(g) is g
    assert np.array_equal(U.normalizar_a_uint8(g), g)          # uint8 passes through without touching the GPU


def test_nifti_roundtrip_and_reference_helpers(tmp_path):
    rng = np.random.default_rng(2)
    aff = np.array([[-1., 0, 0, 90], [0, 1, 0, -126], [0, 0, 1, -72], [0, 0, 0, 1]])
    v = rng.integers(0, 2, (6, 7, 5)).astype(np.float32)
    p = tmp_path / "a" / "P1_axial.nii.gz"
    import torch
    if not torch.cuda.is_available():          # the voxel payload is encoded / decoded on the GPU only: no CPU fallback
        with pytest.raises(RuntimeError):
            U.guardar_volumen(v, aff, p)
    nifti.save(v, aff, p)                      # host-side writer (test tool); the header helpers below read only 348 bytes
    got = nifti.load(p, np.float64)[0]
    assert got.dtype == np.float64 and got.flags["F_CONTIGUOUS"] and np.array_equal(got, v)
    shape, aff2 = U.cargar_referencia_nifti(p)
    assert shape == (6, 7, 5) and np.allclose(aff2, aff)
    q = tmp_path / "b.nii.gz"
    nifti.save(np.zeros((6, 7, 4), np.uint8), aff, q)
    assert U.reconstruccion_valida(p, p) and not U.reconstruccion_valida(q, p)
    with pytest.raises(FileNotFoundError):
        U.cargar_referencia_nifti(tmp_path / "missing.nii.gz")
    bad = tmp_path / "bad.nii.gz"
    bad.write_bytes(b"not a nifti")
    with pytest.raises(ValueError):
        U.cargar_referencia_nifti(bad)


def test_patient_listing_and_small_helpers(tmp_path):
    for n in ("P10", "P2", "P1", ".DS_Store", "x.tmp"):
        (tmp_path / n).mkdir()
    assert U.listar_pacientes(tmp_path) == ["P1", "P2", "P10"]
    assert U.int_o_percentil("12") == 12 and U.int_o_percentil("p50") == "P50"
    with pytest.raises(Exception):
        U.int_o_percentil("abc")
    assert U.evaluar_resultados([]) is None and U.evaluar_resultados([None, None]) is None
    assert U.evaluar_resultados([True, True]) is True and U.evaluar_resultados([True, None]) == "parcial"
    U.escribir_json({"a": 1.5}, tmp_path / "m.json")
    assert U.leer_json(tmp_path / "m.json") == {"a": 1.5}
    with pytest.raises(FileNotFoundError):
        U.leer_json(tmp_path / "nope.json")


def test_recon_file_logic(tmp_path):
    from PIL import Image
    with pytest.raises(FileNotFoundError):
        RV.extraer_indices_png(tmp_path / "nope")
    with pytest.raises(FileNotFoundError):
        RV.extraer_indices_png(tmp_path)
    for name in ("P1_FLAIR_12.png", "P1_FLAIR_3.png", "P1_FLAIR_7_pred.png", "README.png"):
        Image.fromarray(np.zeros((4, 5), np.uint8)).save(tmp_path / name)
    assert RV.extraer_indices_png(tmp_path) == [("P1_FLAIR_3.png", 3), ("P1_FLAIR_7_pred.png", 7), ("P1_FLAIR_12.png", 12)]
    shape = (182, 218, 182)
    RV.validar_corte(0, np.zeros((182, 218)), shape, "axial")
    with pytest.raises(ValueError, match="fuera de rango"):
        RV.validar_corte(182, np.zeros((182, 218)), shape, "axial")
    with pytest.raises(ValueError, match="incorrectas"):
        RV.validar_corte(5, np.zeros((218, 182)), shape, "coronal")
    rgb = np.zeros((4, 5, 3), np.uint8); rgb[1, 2, 0] = 255
    Image.fromarray(rgb).save(tmp_path / "rgb.png")
    assert RV.cargar_mascara_png(tmp_path / "rgb.png").shape == (4, 5)


def test_paciente_argument_validation():
    with pytest.raises(ValueError):
        Paciente("X1", "axial", modalidad=["FLAIR"])
    with pytest.raises(ValueError):
        Paciente("P1", "oblicuo", modalidad=["FLAIR"])
    with pytest.raises(ValueError):
        Paciente("P1", "axial", modalidad=["FLAIR"], mejora="XX")
    with pytest.raises(TypeError):
        Paciente("P1", "axial", modalidad="FLAIR")
    with pytest.raises(ValueError):
        Paciente("P1", "axial", modalidad=["PET"])
    p = Paciente("P7", "coronal", modalidad=["FLAIR", "T1", "FLAIR"], mejora="HE", gt_mask=np.zeros((4, 5, 6)))
    assert p.modalidad == ["FLAIR", "T1"] and p.modalidad_str == "T1FLAIR" and p.num_cortes == 5
    assert str(p) == "P7" and repr(p) == "Paciente(P7)"
    assert p.indice_plano(3) == (slice(None), 3, slice(None))
    with pytest.raises(ValueError):
        Paciente("P7", "consenso", modalidad=["FLAIR"], gt_mask=np.zeros((4, 5, 6))).indice_plano(0)
    with pytest.raises(FileNotFoundError):
        Paciente("P7", "axial", modalidad=["FLAIR"]).cargar_volumen("FLAIR")
