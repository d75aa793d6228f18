"""CPU: the reference's own stage scripts (staged copy / original checkout) reproduce the frozen digests of
tests/golden/stock_pipeline_v1.json, and the import swap refuses to install without the CUDA path."""
import json

import pytest

from conftest import GOLDEN_DIR
from oracle import build_ref

pytestmark = pytest.mark.skipif(not build_ref.available(), reason="reference neither staged (oracle/_ref) nor at /root/reference")


@pytest.fixture(scope="module")
def stock_golden():
    return json.loads((GOLDEN_DIR / "stock_pipeline_v1.json").read_text())


def test_unswapped_reference_matches_golden(stock_golden):
    from oracle import make_golden_stock as G
    ns = build_ref.load()
    _, _, states, snap = G.run_case(ns, "HE_12", **G.CASES["HE_12"])
    got = json.loads(json.dumps(G.digest_of(snap, states), default=str))
    want = stock_golden["cases"]["HE_12"]
    assert got["count"] == want["count"]
    assert got["states"] == want["states"]
    assert got["sha"] == want["sha"]
    assert all(v is True for v in states.values())


def test_staged_copy_is_verbatim():
    if not (build_ref.staged_available() and build_ref.source_available()):
        pytest.skip("needs both the staged copy and the original checkout")
    man = json.loads((build_ref.STAGED / "MANIFEST.json").read_text())["files"]
    assert man == build_ref._manifest(build_ref.SOURCE / build_ref.PKG)
    assert len(man) == 29


def test_swap_names_exist_in_the_reference():
    """Every name the installer rebinds is a real attribute of the reference module it is looked up in
    (anotar_mascaras / normalizar_mascara_binaria included), so a rename upstream is noticed here."""
    import importlib
    build_ref.load()
    from mslesseg_b200.compat import install as I
    for ref_mod, names in I._SWAPS.items():
        mod = importlib.import_module(f"yolo_mslesseg.{ref_mod}")
        for name in names:
            assert hasattr(mod, name), (ref_mod, name)


def test_install_needs_the_cuda_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from mslesseg_b200.compat import install as I
    with pytest.raises(RuntimeError):
        I.install()
    assert not I.installed()
