"""GPU: the drop-in proven under the reference's own stage scripts.

The UNMODIFIED `ejecutar_{dataset,reconstrucciones,eval,consenso,promediar_folds}_pipeline` run on a synthetic
5-patient / 3-fold tree twice: stock (CPU, third-party stand-ins) and after `compat.install.install()` rebound the
hot-path functions to libmslesseg.so.  Every artefact must decode to the same content: PNG pixels (images, GT masks),
label text, NIfTI dtype / shape / voxels / affine, per-patient / fold / global JSON values; the stages' tri-state
results (None = skipped, True, "parcial") must agree too.  The stock arm is also pinned by the frozen digests of
tests/golden/stock_pipeline_v1.json (generated from /root/reference by oracle/make_golden_stock.py)."""
import json
import math
import shutil
from pathlib import Path

import pytest

from conftest import GOLDEN_DIR
from oracle import build_ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not build_ref.available(), reason="reference neither staged (oracle/_ref) nor at /root/reference")]


@pytest.fixture(scope="module")
def env(cuda_device):
    from mslesseg_b200 import _lib
    _lib.load()
    from oracle import make_golden_stock as G, stock_pipeline as SP
    from mslesseg_b200.compat import install as fast
    ns = build_ref.load()
    golden = json.loads((GOLDEN_DIR / "stock_pipeline_v1.json").read_text())
    yield ns, G, SP, fast, golden
    fast.uninstall()


def same(a, b):
    if isinstance(a, float) and isinstance(b, float):
        return a == b or (math.isnan(a) and math.isnan(b))
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(same(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    return a == b


def diff(sa, sb):
    bad = [k for k in sorted(set(sa) | set(sb)) if k not in sa or k not in sb or not same(sa[k], sb[k])]
    return bad


@pytest.mark.parametrize("case", ["CLAHE_P50", "HE_12", "Base_P50"])
def test_swapped_pipeline_equals_stock(env, case):
    ns, G, SP, fast, golden = env
    kw = G.CASES[case]
    fast.uninstall()
    _, _, st_ref, snap_ref = G.run_case(ns, case + "_ref", **kw)
    got = json.loads(json.dumps(G.digest_of(snap_ref, st_ref), default=str))
    assert got["sha"] == golden["cases"][case]["sha"], "stock arm drifted from the frozen reference digests"
    n = fast.install()
    assert n >= 20, n                                   # bindings replaced across the reference's modules
    try:
        root, pats, st_gpu, snap_gpu = G.run_case(ns, case + "_gpu", **kw)
    finally:
        fast.uninstall()
    assert st_gpu == st_ref
    bad = diff(snap_ref, snap_gpu)
    assert not bad, (len(bad), bad[:8])
    assert json.loads(json.dumps(G.digest_of(snap_gpu, st_gpu), default=str))["sha"] == golden["cases"][case]["sha"]


def test_skip_if_exists_and_parcial(env):
    ns, G, SP, fast, golden = env
    kw = G.CASES["HE_12"]
    fast.install()
    try:
        root, pats, st1, snap1 = G.run_case(ns, "HE_12_rerun", **kw)
        assert all(v is True for v in st1.values())
        # second run over the finished tree: every stage skips itself, nothing changes
        st2 = SP.run_all(ns, root, pats, k_folds=3, **kw)
        assert all(v is None for v in st2.values()), st2
        assert not diff(snap1, SP.snapshot(root))
        # remove one patient's dataset and one reconstructed volume: the stages report "parcial" and restore them
        base = Path(root) / "datasets" / "HE" / "FLAIR_12c_3folds"
        shutil.rmtree(base / "fold1" / "P3" / "axial" / "images")
        (base / "fold1" / "P3" / "axial" / "images").mkdir()
        vol = Path(root) / "pred_vols" / "HE" / "FLAIR_12c_3folds_50epochs" / "fold1" / "P9" / "P9_coronal.nii.gz"
        vol.unlink()
        st3 = SP.run_all(ns, root, pats, k_folds=3, **kw)
        assert st3[("dataset", "axial", 0)] == "parcial"
        assert st3[("recon", "coronal", 1)] == "parcial"
        assert st3[("recon", "coronal", 2)] is None
        assert not diff(snap1, SP.snapshot(root))
    finally:
        fast.uninstall()
