"""CPU: host-side logic of the product and the C-ABI surface (no kernel is launched here)."""
import ctypes
import math
import re
from pathlib import Path

import numpy as np
import pytest

from mslesseg_b200 import _lib, metrics as M, tables as T

ROOT = Path(__file__).resolve().parents[1]


def header_functions():
    text = (ROOT / "include" / "mslesseg.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msl_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()                       # loads libmslesseg.so without touching a GPU
    names = header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mslesseg.h but not exported"
    assert sorted(_lib.exported_symbols()) == names
    assert lib.msl_version() == 1


def test_header_constants_match_binding():
    text = (ROOT / "include" / "mslesseg.h").read_text()
    defs = dict(re.findall(r"#define\s+(MSL_[A-Z_0-9]+)\s+(-?\d+)\b", text))
    assert int(defs["MSL_AXIAL"]) == _lib.AXIAL and int(defs["MSL_SAGITAL"]) == _lib.SAGITAL
    assert [int(defs[f"MSL_MEJORA_{m}"]) for m in ("NONE", "HE", "CLAHE", "GC", "LT")] == [0, 1, 2, 3, 4]
    assert int(defs["MSL_OUT_PNG_RGBA"]) == _lib.OUT_PNG_RGBA
    assert int(defs["MSL_TAB_LT"]) == T.TAB_LT and int(defs["MSL_TAB_CM"]) == T.TAB_CM
    assert int(defs["MSL_ERR_WORKSPACE"]) == _lib.ERR_WORKSPACE


def test_argument_errors_without_gpu():
    lib = _lib.load()
    # NULL pointers / bad enums are rejected before any CUDA call
    rc = lib.msl_confusion_counts(None, None, 1, 8, None, None)
    assert rc == _lib.ERR_ARG and b"NULL" in lib.msl_last_error()
    rc = lib.msl_enhance_images(ctypes.c_void_p(16), 0, 1, 8, 8, 64, 9, ctypes.c_void_p(16), 64, 0, ctypes.c_void_p(16), None)
    assert rc == _lib.ERR_ARG and b"mejora" in lib.msl_last_error()
    rc = lib.msl_enhance_images(ctypes.c_void_p(16), 0, 1, 8, 8, 32, 0, ctypes.c_void_p(16), 64, 0, ctypes.c_void_p(16), None)
    assert rc == _lib.ERR_ARG and b"pitch" in lib.msl_last_error()
    assert lib.msl_workspace_bytes(_lib.WS_RECON, 2, 182, 218, 182) == 2 * (218 + 2) * 4
    assert lib.msl_workspace_bytes(_lib.WS_ENHANCE_VOLUMES, 1, 182, 218, 182) > 3 * 7221032
    with pytest.raises(_lib.MslError):
        _lib.check(rc)


def test_argument_errors_of_the_widened_entry_points_without_gpu():
    lib = _lib.load()
    P = ctypes.c_void_p(4096)
    assert lib.msl_png_bytes(218, 182, 4) == 159000 and lib.msl_png_bytes(33, 21, 1) == 794 and lib.msl_png_bytes(5, 5, 3) == 0
    assert lib.msl_png_pack(P, 1, 218, 182, 3, P, 160000, None) == _lib.ERR_ARG and b"channels" in lib.msl_last_error()
    assert lib.msl_png_pack(P, 1, 218, 182, 4, P, 1000, None) == _lib.ERR_ARG and b"pitch" in lib.msl_last_error()
    assert lib.msl_png_pack(None, 0, 218, 182, 4, None, 0, None) == 0                       # nothing to do
    assert lib.msl_combine_predictions(P, P, 1, 640, 544, 182, 218, _lib.OUT_PNG_RGBA, P, None) == _lib.ERR_ARG
    assert b"layout" in lib.msl_last_error()
    assert lib.msl_combine_predictions(None, None, 0, 1, 1, 182, 218, _lib.OUT_G, None, None) == 0
    assert lib.msl_slice_counts(P, None, 1, 8, 8, 8, P, None) == _lib.ERR_ARG and b"NULL" in lib.msl_last_error()
    assert lib.msl_copy_box_d2h(P, P, 10, 10, 10, 3, 2, 0, 10, None) == _lib.ERR_ARG and b"box" in lib.msl_last_error()
    assert lib.msl_copy_box_d2h(P, P, 10, 10, 10, 2, 2, 0, 10, None) == 0                   # empty box: no copy issued
    assert lib.msl_nonzero_flags(P, 1, 0, 4, 4, P, P, None) == _lib.ERR_ARG
    assert lib.msl_bgr_to_gray(None, 0, None, None) == 0


def test_argument_errors_of_the_round2_entry_points_without_gpu():
    lib = _lib.load()
    P, Q = ctypes.c_void_p(4096), ctypes.c_void_p(4100)
    # staged-stack route: pitch / alignment / outputs are checked before any CUDA call
    assert lib.msl_enhance_stack_workspace_bytes(182, 218) > 0 and lib.msl_enhance_stack_workspace_bytes(182, 218) % 16 == 0
    assert lib.msl_enhance_stack(P, 39680, 4, 182, 218, None, None, None, None, P, P, 1 << 20, None) == _lib.ERR_ARG
    assert b"no output" in lib.msl_last_error()
    assert lib.msl_enhance_stack(P, 39676, 4, 182, 218, P, None, None, None, P, P, 1 << 20, None) == _lib.ERR_ARG
    assert b"pitch" in lib.msl_last_error()
    assert lib.msl_enhance_stack(Q, 39680, 4, 182, 218, P, None, None, None, P, P, 1 << 20, None) == _lib.ERR_ARG
    assert lib.msl_enhance_stack(P, 39680, 0, 182, 218, P, None, None, None, P, P, 1 << 20, None) == 0          # nothing to do
    assert lib.msl_enhance_stack(P, 4 * 3001 * 3001, 1, 3001, 3001, P, None, None, None, P, P, 1 << 20, None) == _lib.ERR_ARG   # pitch multiple of 16 fails first
    assert lib.msl_stage_slices(P, 1, 182, 218, 182, 7, None, None, 182, P, 39680, None) == _lib.ERR_ARG and b"Plano" in lib.msl_last_error()
    assert lib.msl_stage_slices(P, 1, 182, 218, 182, 0, None, None, 5, P, 39680, None) == _lib.ERR_ARG and b"dense mode" in lib.msl_last_error()
    assert lib.msl_stage_slices(P, 1, 182, 218, 182, 0, P, P, 5, P, 100, None) == _lib.ERR_ARG and b"pitch" in lib.msl_last_error()
    assert lib.msl_stage_slices(P, 1, 181, 218, 182, 0, P, P, 5, P, 39680, None) == _lib.ERR_UNSUPPORTED      # odd rows: the slice kernel's job
    assert lib.msl_stage_slices(P, 1, 182, 218, 182, 0, P, P, 0, P, 39680, None) == 0
    assert lib.msl_selftest_norm_division(None, None, 4, P, None) == _lib.ERR_ARG
    # codec: sizes, distances, workspace
    assert lib.msl_deflate_bound(4, _lib.Z_GZIP, 16384) >= 4 * 16384 and lib.msl_deflate_workspace_bytes(4, _lib.Z_GZIP, 16384) > lib.msl_deflate_bound(4, _lib.Z_GZIP, 16384)
    assert lib.msl_profile_timeline(0, None, None, None, None) == 0


def test_host_box_logic():
    from mslesseg_b200 import ops
    a = np.zeros(10, np.uint8); b = np.zeros(7, np.uint8)
    assert ops.box_from_flags(a, b) == (0, 0, 0, 0)
    a[3] = a[8] = 1; b[0] = 1
    assert ops.box_from_flags(a, b) == (3, 9, 0, 1)
    b[:] = 0
    assert ops.box_from_flags(a, b) == (0, 0, 0, 0)


def test_tables():
    t = T.host_tables()
    assert t.shape == (T.TABLES_BYTES,)
    assert t[T.TAB_GC + 16] == 1 and t[T.TAB_GC + 255] == 255
    lt255 = t[T.TAB_LT + 255 * 256:T.TAB_LT + 256 * 256]
    assert lt255[:9].tolist() == [0, 31, 50, 63, 74, 82, 89, 95, 101] and lt255[255] == 255
    assert t[T.TAB_LT] == 0
    cm = t[T.TAB_CM:T.TAB_CM + 256]
    off = [i for i in range(256) if cm[i] != i]
    assert off == [33, 37, 41, 45, 49, 53, 57, 61, 66, 74, 82, 90, 98, 106, 114, 122, 132, 148, 164, 180, 196, 212, 228, 244]
    assert all(cm[i] == i - 1 for i in off)
    cv2 = pytest.importorskip("cv2")
    g = np.arange(256, dtype=np.uint8).reshape(16, 16)
    assert np.array_equal(cv2.cvtColor(cv2.cvtColor(g, cv2.COLOR_GRAY2BGR), cv2.COLOR_BGR2LAB)[..., 0].ravel(), T.LUT_L)
    lab = np.stack([g, np.full_like(g, 128), np.full_like(g, 128)], axis=-1)
    assert np.array_equal(cv2.cvtColor(cv2.cvtColor(lab, cv2.COLOR_LAB2BGR), cv2.COLOR_BGR2GRAY).ravel(), T.LUT_OUT)


def test_metrics_from_counts_against_reference(golden):
    for case in golden["random_metricas"]:
        assert M.metricas_desde_conteos(*case["counts"]) == case["metricas"]
    for pid, e in golden["synthetic_eval"].items():
        for plano, pe in e["planes"].items():
            assert M.metricas_desde_conteos(*pe["counts"]) == pe["metricas"]
        for u in (2, 3):
            assert M.metricas_desde_conteos(*e[f"consenso{u}"]["counts"]) == e[f"consenso{u}"]["metricas"]

    def norm(d):
        return {k: (None if (isinstance(v, float) and math.isnan(v)) else v) for k, v in d.items()}
    assert norm(M.metricas_desde_conteos(0, 1, 0, 119)) == golden["edge_metricas"]["empty_gt"]
    assert norm(M.metricas_desde_conteos(0, 0, 2, 118)) == golden["edge_metricas"]["empty_pred"]
    assert norm(M.metricas_desde_conteos(2, 0, 0, 118)) == golden["edge_metricas"]["perfect"]


def test_fold_logic_against_reference(golden):
    assert M.calcular_promedio(golden["promedio_fold"]["in"]) == golden["promedio_fold"]["out"]
    assert M.calcular_resumen_experimento(golden["resumen_experimento"]["in"]) == golden["resumen_experimento"]["out"]
    for k, table in golden["calcular_fold"].items():
        for pid, fold in table.items():
            assert M.calcular_fold(pid, int(k)) == fold
    with pytest.raises(ValueError):
        M.calcular_fold("P54")
    assert M.num_cortes_percentil(golden["percentil"]["in"], 50) == golden["percentil"]["P50"]
    with pytest.raises(ValueError):
        M.calcular_promedio({})
    d = golden["demo"]["P39"]["axial"]
    assert M.ventana_central(list(range(50, 151)), 20) == list(range(90, 110))
    assert M.ventana_central([1, 2, 3], 20) == [1, 2, 3] and M.ventana_central([1, 2, 3], None) == [1, 2, 3]
    assert len(d["usar20"]) == 20


def test_iou_and_rango_global():
    from mslesseg_b200 import metrics as M
    assert M.iou_desde_conteos(3, 1, 2) == 0.5
    assert M.iou_desde_conteos(0, 0, 0) == 0.0
    assert M.iou_desde_conteos(7, 0, 0) == 1.0
    m = M.metricas_desde_conteos(30, 10, 20, 940)
    assert "IoU" not in m                                   # reference JSONs stay identical by default
    m2 = M.metricas_desde_conteos(30, 10, 20, 940, con_iou=True)
    assert m2["IoU"] == 0.5 and {k: v for k, v in m2.items() if k != "IoU"} == m
    # Jaccard and Dice of the same counts: J = D / (2 - D) up to the rounding of both
    assert abs(m2["IoU"] - m2["DSC"] / (2 - m2["DSC"])) < 2e-3
    r = np.array([[0.0, 10.0], [-3.0, 4.0], [1.0, 99.0]])
    assert M.calcular_rango_global(r) == (-3.0, 99.0)
    assert M.calcular_rango_global(r, [0, 2]) == (0.0, 99.0)
    with pytest.raises(ValueError):
        M.calcular_rango_global(r, [])
