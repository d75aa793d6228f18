"""GPU: YOLO label polygons - msl_mask_contours against cv2.findContours and the frozen label text of
the demo patients' ground-truth masks (tests/golden/demo_labels_v1.json, oracle/make_golden_labels.py)."""
import hashlib
import json

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda_device):
    from mslesseg_b200 import _lib, ops
    _lib.load()
    return ops


def _cv2_contours(m):
    import cv2
    cs, _ = cv2.findContours(np.ascontiguousarray(m), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    return [c.reshape(-1, 2) for c in cs]


def test_contours_equal_cv2_on_random_masks(ops, cuda_device):
    import torch
    pytest.importorskip("cv2")
    from scipy import ndimage
    rng = np.random.default_rng(0)
    for H, W in ((37, 53), (182, 218), (218, 182), (5, 3), (1, 1), (64, 64)):
        masks = []
        for t in range(24):
            k = t % 4
            if k == 0:
                m = (rng.random((H, W)) < rng.uniform(0.02, 0.7)).astype(np.uint8)
            elif k == 1:
                m = (ndimage.gaussian_filter(rng.random((H, W)), 1.5) > 0.5).astype(np.uint8)
            elif k == 2:
                m = np.zeros((H, W), np.uint8)
                yy, xx = np.mgrid[0:H, 0:W]
                for _ in range(int(rng.integers(1, 7))):
                    cy, cx, r = rng.integers(0, H), rng.integers(0, W), rng.integers(1, 9)
                    m |= ((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r).astype(np.uint8)
                cy, cx, r = rng.integers(0, H), rng.integers(0, W), rng.integers(1, 5)
                m[((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r)] = 0          # holes, rings, things inside holes
                m[min(cy, H - 1), min(cx, W - 1)] = 1
            else:
                m = np.zeros((H, W), np.uint8) if t % 8 == 3 else np.ones((H, W), np.uint8)
            masks.append(m * (255 if t % 3 == 0 else 1))
        st = np.stack(masks)
        got = ops.mask_contours(torch.from_numpy(st).to(cuda_device), value=0, max_contours=8, max_points=64)   # tiny capacities: regrown
        for m, g in zip(st, got):
            want = _cv2_contours((m != 0).astype(np.uint8))
            assert len(want) == len(g)
            for a, b in zip(want, g):
                assert np.array_equal(a, b)
        # value = 1 selects exactly the pixels equal to 1
        got1 = ops.mask_contours(torch.from_numpy(st).to(cuda_device), value=1)
        for m, g in zip(st, got1):
            want = _cv2_contours((m == 1).astype(np.uint8))
            assert len(want) == len(g) and all(np.array_equal(a, b) for a, b in zip(want, g))


def test_demo_masks_label_text_equals_golden(ops, cuda_device):
    import torch
    from mslesseg_b200.compat.extraer_dataset import lineas_yolo
    z = np.load(GOLDEN_DIR / "demo_label_masks.npz")
    gold = json.loads((GOLDEN_DIR / "demo_labels_v1.json").read_text())
    for key, g in gold.items():
        shape = tuple(int(d) for d in z[key + "_shape"])
        masks = np.unpackbits(z[key + "_bits"])[:int(np.prod(shape))].reshape(shape)
        cont = ops.mask_contours(torch.from_numpy(masks).to(cuda_device), value=1)
        n, h, w = shape
        total = 0
        for i in range(n):
            lines = lineas_yolo(cont[i], w, h)
            total += len(lines)
            text = "".join(ln + "\n" for ln in lines)
            assert hashlib.sha256(text.encode()).hexdigest() == g["sha"][i], (key, i)
        assert total == g["contours"]


def test_anotar_mascaras_shim(ops, cuda_device, tmp_path):
    """Directory in, directory out: RGBA {0, 255} mask PNGs (what guardar_cortes writes) -> binary gray PNGs + labels."""
    import torch
    from PIL import Image
    from oracle import ref_stubs
    from mslesseg_b200.compat import extraer_dataset as ED, utils as U
    rng = np.random.default_rng(4)
    gdir, ldir = tmp_path / "GT_masks", tmp_path / "labels"
    gdir.mkdir(); ldir.mkdir()
    masks = {}
    for i in range(5):
        m = np.zeros((218, 182), np.uint8)
        for _ in range(int(rng.integers(0, 5))):
            cy, cx, r = rng.integers(5, 210), rng.integers(5, 175), rng.integers(1, 9)
            yy, xx = np.mgrid[0:218, 0:182]
            m |= ((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r).astype(np.uint8)
        masks[f"P7_{40 + i}"] = m
        rgba = np.stack([m * 255] * 3 + [np.full_like(m, 255)], axis=-1)
        Image.fromarray(rgba, "RGBA").save(gdir / f"P7_{40 + i}.png")
    ED.anotar_mascaras(gdir, ldir)
    for stem, m in masks.items():
        im = Image.open(gdir / f"{stem}.png")
        assert im.mode == "L" and np.array_equal(np.array(im), m)
        want = "".join(ln + "\n" for ln in ref_stubs.yolo_seg_lines(m, 1))
        assert (ldir / f"{stem}.txt").read_text() == want
    # single-file form
    Image.fromarray(np.stack([masks["P7_40"] * 255] * 3 + [np.full((218, 182), 255, np.uint8)], axis=-1), "RGBA").save(tmp_path / "one.png")
    U.normalizar_mascara_binaria(tmp_path / "one.png")
    assert np.array_equal(np.array(Image.open(tmp_path / "one.png")), masks["P7_40"])
    with pytest.raises(FileNotFoundError):
        ED.anotar_mascaras(ldir, ldir)
