"""GPU: the device byte-stream codec (msl_codec.cu / msl_inflate.cu) against zlib, gzip and Pillow on the host."""
import gzip
import io
import struct
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda_device):
    from mslesseg_b200 import _lib, ops
    _lib.load()
    return ops


def _payloads():
    rng = np.random.default_rng(5)
    out = {}
    out["zeros"] = np.zeros(200_000, np.uint8)
    out["noise"] = rng.integers(0, 256, 70_001, dtype=np.uint8)
    mixed = np.zeros(300_000, np.uint8)
    mixed[50_000:90_000] = rng.integers(0, 256, 40_000, dtype=np.uint8)
    mixed[150_000:150_700] = 7
    mixed[200_000:260_000:3] = 255
    out["mixed"] = mixed
    f = np.zeros(60_000, np.float32)
    f[10_000:40_000] = np.round(rng.uniform(1, 1500, 30_000))
    out["float32"] = f.view(np.uint8)
    out["tiny"] = np.array([1, 2, 3], np.uint8)
    out["one_run"] = np.full(65_536 * 2 + 5, 9, np.uint8)
    out["empty"] = np.zeros(0, np.uint8)
    return out


@pytest.mark.parametrize("container", ["raw", "zlib", "gzip"])
def test_deflate_chunks_decode_with_zlib(ops, cuda_device, container):
    import torch
    for name, data in _payloads().items():
        for chunk, d2 in ((65536, 0), (65536, 4), (20_000, 4), (1 << 20, 0)):
            src = torch.from_numpy(data.copy()).to(cuda_device)
            ps = ops.deflate_chunks(src, chunk_len=chunk, container=container, dist2=d2)
            packed, off = ps.to_host()
            meta = ps.meta.cpu().numpy().astype(np.int64) & 0xffffffff
            n = len(off) - 1
            assert n == max(1, -(-data.size // chunk))
            got = []
            for i in range(n):
                b = packed[off[i]:off[i + 1]].tobytes()
                assert meta[i, 0] == len(b)
                if container == "raw":
                    got.append(zlib.decompress(b, wbits=-15))
                elif container == "zlib":
                    got.append(zlib.decompress(b))
                    assert meta[i, 2] == zlib.adler32(got[-1])
                else:
                    got.append(gzip.decompress(b))
                    assert meta[i, 2] == zlib.crc32(got[-1])
                    assert b[12:14] == b"MS" and int.from_bytes(b[16:20], "little") == len(b) and int.from_bytes(b[20:24], "little") == len(got[-1])
                assert meta[i, 1] == len(got[-1])
            assert b"".join(got) == data.tobytes(), (name, chunk, d2)
            if container == "gzip":          # the members back to back are one valid .gz file
                assert gzip.decompress(packed[:off[-1]].tobytes()) == data.tobytes()
            if name in ("zeros", "one_run") and data.size > 100_000:
                assert off[-1] < data.size // 40, (name, off[-1])
            if name == "noise":
                assert off[-1] < data.size * 1.07 + 64 * n


def test_png_encode_decodes_with_pillow(ops, cuda_device):
    import torch
    from PIL import Image
    rng = np.random.default_rng(9)
    for shape in ((5, 218, 182), (3, 182, 218, 4), (2, 37, 53, 3), (2, 16, 20, 2), (1, 1, 1), (4, 182, 182, 4)):
        px = np.zeros(shape, np.uint8)
        if px.ndim == 3:
            px[:, 5:-5, 7:-7] = rng.integers(0, 256, px[:, 5:-5, 7:-7].shape, dtype=np.uint8) if shape[1] > 10 else 200
            px[0] = 0
        else:
            g = np.zeros(shape[:3], np.uint8)
            g[:, 3:-3, 4:-4] = rng.integers(0, 256, g[:, 3:-3, 4:-4].shape, dtype=np.uint8)
            px[..., :] = g[..., None]
            if shape[3] in (2, 4):
                px[..., -1] = 255
        files = ops.png_encode(torch.from_numpy(px).to(cuda_device)).files()
        assert len(files) == shape[0]
        for i, f in enumerate(files):
            im = Image.open(io.BytesIO(f))
            im.load()
            a = np.array(im)
            assert im.mode == {3: "L", 4: {4: "RGBA", 3: "RGB", 2: "LA"}.get(shape[-1])}[px.ndim] if px.ndim == 4 else im.mode == "L"
            assert np.array_equal(a, px[i]), (shape, i)
        stored = sum(len(f) for f in files)
        assert stored < px.size * 1.08 + 200 * shape[0]
    # a blank stack compresses to almost nothing; cv2 reads the files too
    cv2 = pytest.importorskip("cv2")
    blank = torch.zeros((3, 218, 182), dtype=torch.uint8, device=cuda_device)
    files = ops.png_encode(blank).files()
    assert all(len(f) < 1000 for f in files)
    assert np.array_equal(cv2.imdecode(np.frombuffer(files[0], np.uint8), cv2.IMREAD_UNCHANGED), np.zeros((218, 182), np.uint8))


@pytest.fixture(scope="module")
def codec(ops):
    from mslesseg_b200 import codec
    return codec


def test_inflate_decodes_zlib_gzip_and_own_streams(ops, codec, cuda_device):
    import torch
    pl = _payloads()
    # streams written by zlib at several levels (dynamic, fixed and stored blocks), containers raw / zlib / gzip
    for container, wbits in (("raw", -15), ("zlib", 15), ("gzip", 31)):
        pieces, sizes, want = [], [], []
        for name, data in pl.items():
            for level in (0, 1, 6, 9):
                c = zlib.compressobj(level, zlib.DEFLATED, wbits)
                pieces.append(c.compress(data.tobytes()) + c.flush())
                sizes.append(data.size)
                want.append(data)
            c = zlib.compressobj(6, zlib.DEFLATED, wbits, 9, zlib.Z_FIXED)
            pieces.append(c.compress(data.tobytes()) + c.flush())
            sizes.append(data.size)
            want.append(data)
        out, off = codec.inflate(pieces, sizes, container, cuda_device)
        host = out.cpu().numpy()
        for i, w in enumerate(want):
            assert np.array_equal(host[off[i]:off[i] + w.size], w), (container, i)
    # multi-member gzip as ONE stream (a foreign file: `cat a.gz b.gz`), and our own members one warp each
    a, b = pl["mixed"], pl["noise"]
    cat = gzip.compress(a.tobytes(), 6) + gzip.compress(b.tobytes(), 1)
    out, off = codec.inflate([cat], [a.size + b.size], "gzip", cuda_device)
    assert np.array_equal(out.cpu().numpy()[:a.size + b.size], np.concatenate([a, b]))
    ps = ops.deflate_chunks(torch.from_numpy(pl["float32"].copy()).to(cuda_device), chunk_len=65536, container="gzip", dist2=4)
    data, o = ps.to_host()
    blob = data[:o[-1]].tobytes()
    members = codec.gzip_members(blob)
    assert members is not None and len(members) == len(o) - 1
    out, off = codec.inflate([blob[p:p + m] for p, m, _ in members], [r for _, _, r in members], "gzip", cuda_device)
    got = np.concatenate([out.cpu().numpy()[off[i]:off[i] + members[i][2]] for i in range(len(members))])
    assert np.array_equal(got, pl["float32"])
    # far matches (a 20 KB block repeated: distances beyond the 2 KB output ring, served from global memory), overlapping
    # short-period matches of every period 1..9, and a gzip header with FEXTRA + FNAME + FCOMMENT + FHCRC
    rng = np.random.default_rng(9)
    block = rng.integers(0, 256, 20_000, dtype=np.uint8)
    far = np.concatenate([block, block, block[:7000], rng.integers(0, 256, 3000, dtype=np.uint8), block[5000:]])
    periods = np.concatenate([np.resize(rng.integers(0, 256, p_, dtype=np.uint8), 3000 + 37 * p_) for p_ in range(1, 10)])
    for data in (far, periods):
        for level in (1, 6, 9):
            out, off = codec.inflate([zlib.compress(data.tobytes(), level)], [data.size], "zlib", cuda_device)
            assert np.array_equal(out.cpu().numpy()[:data.size], data), (data.size, level)
    raw_member = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = raw_member.compress(far.tobytes()) + raw_member.flush()
    hdr = bytearray(b"\x1f\x8b\x08\x1e\x00\x00\x00\x00\x00\xff")                       # FLG = FEXTRA | FNAME | FCOMMENT | FHCRC
    hdr += struct.pack("<H", 6) + b"XX\x02\x00ab" + b"name.nii\x00" + b"a comment\x00"
    hdr += struct.pack("<H", zlib.crc32(bytes(hdr)) & 0xffff)
    member = bytes(hdr) + body + struct.pack("<II", zlib.crc32(far.tobytes()), far.size)
    assert gzip.decompress(member) == far.tobytes()
    out, off = codec.inflate([member], [far.size], "gzip", cuda_device)
    assert np.array_equal(out.cpu().numpy()[:far.size], far)
    # damaged input is reported, not decoded
    bad = bytearray(zlib.compress(pl["mixed"].tobytes(), 6))
    bad[40] ^= 0x55
    with pytest.raises(codec.CodecError):
        codec.inflate([bytes(bad)], [pl["mixed"].size], "zlib", cuda_device)
    with pytest.raises(codec.CodecError):
        codec.inflate([zlib.compress(pl["noise"].tobytes())], [100], "zlib", cuda_device)      # output region too small


def test_png_decode_first_channel(ops, codec, cuda_device):
    import torch
    from PIL import Image
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(12)
    masks = (rng.random((6, 182, 218)) > 0.97).astype(np.uint8) * 255
    masks[0] = 0
    files = []
    for i, m in enumerate(masks):       # cv2.imwrite like guardar_prediccion (scripts/generar_predicciones.py:153)
        ok, buf = cv2.imencode(".png", m, [cv2.IMWRITE_PNG_COMPRESSION, 3])
        files.append(buf.tobytes())
    got = codec.png_decode_first_channel(files, cuda_device).cpu().numpy()
    assert np.array_equal(got, masks)
    # Pillow with its adaptive filters, gray / RGB / RGBA / LA, smooth content so that every filter type shows up
    yy, xx = np.mgrid[0:97, 0:131]
    for mode, bpp in (("L", 1), ("RGB", 3), ("RGBA", 4), ("LA", 2)):
        imgs, files = [], []
        for k in range(4):
            base = ((np.sin(xx / (7.0 + k)) + np.cos(yy / (5.0 + k))) * 60 + 128 + rng.integers(0, 3 + 20 * k, xx.shape)).clip(0, 255).astype(np.uint8)
            arr = base if bpp == 1 else np.stack([np.roll(base, c * 3, axis=1) for c in range(bpp)], axis=-1)
            bio = io.BytesIO()
            Image.fromarray(arr, mode).save(bio, format="PNG", compress_level=k + 1)
            files.append(bio.getvalue())
            imgs.append(arr if bpp == 1 else arr[..., 0])
        got = codec.png_decode_first_channel(files, cuda_device).cpu().numpy()
        assert np.array_equal(got, np.stack(imgs)), mode
    # hand-built files: noise images, a random None / Sub / Up filter per scanline (the row-parallel path; 257 columns: serial)
    def png_of(img, fts):
        rows = []
        for y, ft in enumerate(fts):
            cur = img[y].astype(np.int16)
            ref = np.zeros_like(cur) if ft == 0 else (np.concatenate([[0], cur[:-1]]) if ft == 1 else (img[y - 1].astype(np.int16) if y else np.zeros_like(cur)))
            rows.append(bytes([ft]) + ((cur - ref) & 0xff).astype(np.uint8).tobytes())
        def chunk(t, d):
            return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
        h, w = img.shape
        return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"".join(rows), 6))
                + chunk(b"IEND", b""))
    for w in (182, 256, 257, 31):
        imgs = rng.integers(0, 256, (5, 73, w), dtype=np.uint8)
        files = [png_of(im, rng.integers(0, 3, 73) if k else np.full(73, 2)) for k, im in enumerate(imgs)]
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(files[1]))), imgs[1])
        assert np.array_equal(codec.png_decode_first_channel(files, cuda_device).cpu().numpy(), imgs), w
    # our own encoder's files
    px = torch.from_numpy(masks).to(cuda_device)
    assert np.array_equal(codec.png_decode_first_channel(ops.png_encode(px).files(), cuda_device).cpu().numpy(), masks)
    with pytest.raises(codec.CodecError):
        bio = io.BytesIO()
        Image.fromarray(masks[1]).convert("P").save(bio, format="PNG")
        codec.png_decode_first_channel([bio.getvalue()], cuda_device)


def test_nifti_gz_round_trip_on_device(ops, codec, cuda_device, tmp_path):
    import torch
    from mslesseg_b200 import nifti, synthetic as S
    pat = S.make_patient(2, config_id=1, num_cortes=10)
    aff = np.array([[1.0, 0, 0, -90], [0, 1.0, 0, -126], [0, 0, 1.0, -72], [0, 0, 0, 1.0]])
    # a file written by the host writer with Python's gzip (one member, dynamic Huffman) -> device reader
    p1 = tmp_path / "P2_T1_FLAIR.nii.gz"
    nifti.save(S.as_xyz(pat.flair), aff, p1)
    vol, shape, affine = codec.nifti_load_device(p1, cuda_device, torch.float32)
    assert shape == (182, 218, 182) and np.allclose(affine, aff)
    assert np.array_equal(vol.cpu().numpy(), pat.flair)
    p2 = tmp_path / "P2_MASK.nii.gz"
    nifti.save(S.as_xyz(pat.gt).astype(np.float32), aff, p2)
    m, _, _ = codec.nifti_load_device(p2, cuda_device, torch.uint8)
    assert np.array_equal(m.cpu().numpy(), pat.gt)
    # device writer -> Python's gzip / host reader, and back through the parallel member path
    p3 = tmp_path / "out" / "P2_axial.nii.gz"
    size = codec.nifti_save_device(vol, aff, p3)
    assert size == p3.stat().st_size and size < pat.flair.nbytes // 2
    arr, aff2 = nifti.load(p3)
    assert arr.dtype == np.float32 and np.array_equal(S.as_zyx(arr) if hasattr(S, "as_zyx") else arr.transpose(2, 1, 0), pat.flair) and np.allclose(aff2, aff)
    raw = gzip.decompress(p3.read_bytes())
    assert len(raw) == 352 + pat.flair.nbytes
    assert codec.gzip_members(p3.read_bytes()) is not None
    vol2, _, _ = codec.nifti_load_device(p3, cuda_device, torch.float32)
    assert torch.equal(vol2, vol)
    p4 = tmp_path / "P2_consenso.nii.gz"
    codec.nifti_save_device(m, aff, p4)
    assert p4.stat().st_size < 250_000
    arr, _ = nifti.load(p4)
    assert arr.dtype == np.uint8 and np.array_equal(arr.transpose(2, 1, 0), pat.gt)
    assert codec.nifti_read_header(p4)[0] == (182, 218, 182)
    # values a uint8 volume cannot hold are refused
    with pytest.raises(codec.CodecError):
        codec.nifti_load_device(p1, cuda_device, torch.uint8)
    # dataset preparation: a foreign single-member file is rewritten with the member index, its decoded bytes unchanged
    before = gzip.decompress(p1.read_bytes())
    assert codec.gzip_member_table(p1.read_bytes()) is None
    assert codec.reindex_gz(p1) is True and codec.reindex_gz(p1) is False
    assert codec.gzip_member_table(p1.read_bytes()) is not None and gzip.decompress(p1.read_bytes()) == before
    vol3, _, _ = codec.nifti_load_device(p1, cuda_device, torch.float32)
    assert torch.equal(vol3, vol)


def test_deflate_files_prefix_expand_and_index(ops, codec, cuda_device):
    """Several files in one launch: shared header prefix + body per file, uint8 masks stored as float32, and the index
    member that lets a reader split the file without hopping from member to member."""
    import torch
    rng = np.random.default_rng(3)
    masks = (rng.random((3, 20, 30, 40)) > 0.98).astype(np.uint8)
    masks[1] = 0
    aff = np.diag([1.0, 1.0, 1.0, 1.0])
    ps = codec.nifti_gz_device(torch.from_numpy(masks).to(cuda_device), aff, como_float32=True)
    assert ps.streams_per_file == -(-(352 + masks[0].size * 4) // codec.CHUNK)
    for f in range(3):
        blob = codec.nifti_gz_bytes(ps, f)
        raw = gzip.decompress(blob)
        assert len(raw) == 352 + masks[f].size * 4
        assert raw[:352] == codec.nifti_header_bytes((40, 30, 20), np.float32, aff)
        assert np.array_equal(np.frombuffer(raw, "<f4", offset=352), masks[f].reshape(-1).astype(np.float32))
        tab = codec.gzip_member_table(blob)
        assert tab is not None and len(tab) == ps.streams_per_file and int(tab[:, 2].sum()) == len(raw)
    # plain float32 volumes, per-file prefixes
    vols = np.round(rng.uniform(0, 900, (2, 9, 11, 13))).astype(np.float32)
    pref = np.stack([np.frombuffer(codec.nifti_header_bytes((13, 11, 9), np.float32, aff * (k + 1)), np.uint8) for k in range(2)])
    ps = ops.deflate_files(torch.from_numpy(vols).to(cuda_device), prefix=torch.from_numpy(pref.copy()).to(cuda_device), chunk_len=1000, dist2=4)
    data, off = ps.to_host()
    spv = ps.streams_per_file
    for f in range(2):
        raw = gzip.decompress(data[off[f * spv]:off[(f + 1) * spv]].tobytes())
        assert raw == pref[f].tobytes() + vols[f].tobytes()


def test_own_streams_round_trip_all_distances(ops, codec, cuda_device):
    """What msl_deflate_* writes (stored blocks + fixed-Huffman run blocks, one match distance 1..4) decodes on the device
    and with zlib, for every container, chunk size and distance."""
    import torch
    pl = _payloads()
    for container in ("raw", "zlib", "gzip"):
        for name, data in pl.items():
            for chunk, d2 in ((16384, 0), (16384, 4), (5000, 2), (3001, 3), (1 << 18, 1)):
                ps = ops.deflate_chunks(torch.from_numpy(data.copy()).to(cuda_device), chunk_len=chunk, container=container, dist2=d2)
                packed, off = ps.to_host()
                meta = ps.meta.cpu().numpy().astype(np.int64) & 0xffffffff
                pieces = [packed[off[i]:off[i + 1]].tobytes() for i in range(len(off) - 1)]
                out, o = codec.inflate(pieces, [int(r) for r in meta[:, 1]], container, cuda_device)
                host = out.cpu().numpy()
                got = np.concatenate([host[o[i]:o[i] + meta[i, 1]] for i in range(len(pieces))]) if len(pieces) else np.zeros(0, np.uint8)
                assert np.array_equal(got, data), (container, name, chunk, d2)


def test_run_groups_of_every_length(ops, cuda_device):
    """A run group of k segments (k = 1 .. 256: one fixed block of 258-byte matches + the rest, 1808 bytes being the case whose
    rest would be two bytes) between literal segments, at every position class of a 4 KB tile and across tile borders:
    zlib on the host decodes what the device wrote."""
    import torch
    rng = np.random.default_rng(77)
    streams, chunk = [], 3 * 4096
    for k in list(range(1, 257)) + [113, 113, 256, 255, 17, 33]:
        for d, lead in ((1, int(rng.integers(0, 40))), (4, int(rng.integers(0, 40)))):
            buf = rng.integers(1, 255, chunk, dtype=np.uint8)
            # literal bytes that never repeat at distance d inside a segment (so that only the planted run is a run)
            buf[::2] = (np.arange(chunk // 2) * 7 + 3) % 251 + 1
            a = 16 * lead
            b = min(a + 16 * k, chunk)
            pattern = rng.integers(1, 255, d, dtype=np.uint8)
            buf[a:b] = np.resize(pattern, b - a)
            if a >= d:
                buf[a - d:a] = pattern                      # the run has something to refer back to
            streams.append((d, buf))
    for d in (1, 4):
        data = np.concatenate([b for dd, b in streams if dd == d])
        for container in ("zlib", "gzip"):
            ps = ops.deflate_chunks(torch.from_numpy(data).to(cuda_device), chunk_len=chunk, container=container, dist2=d)
            packed, off = ps.to_host()
            back = b"".join((zlib.decompress(packed[off[i]:off[i + 1]].tobytes()) if container == "zlib"
                             else gzip.decompress(packed[off[i]:off[i + 1]].tobytes())) for i in range(len(off) - 1))
            assert back == data.tobytes(), (d, container)
            assert off[-1] < data.size                                   # the planted runs were found
