"""JPEG loader, CPU side: the arithmetic shared with the device kernels (csrc/jpeg_core.h, built for
the CPU by oracle/jpeg_check.py) against PIL -- the decoder the reference calls (compute_feats.py:28) -- bit for
bit; and the host header parser of libdsmil_host.so."""
import io

import numpy as np
import pytest
from PIL import Image

from oracle import jpeg_check
import jpeg_cases as jc


@pytest.mark.parametrize("hw", [(224, 224), (256, 256), (223, 217), (17, 33), (8, 8), (1, 1), (5, 3), (100, 2), (64, 250)])
@pytest.mark.parametrize("subsampling", [0, 1, 2])
def test_core_arithmetic_matches_pil_bit_for_bit(hw, subsampling):
    h, w = hw
    for kind, img in (("noise", jc.noise(h, w, 1)), ("tissue", jc.histology_like(h, w, 2))):
        for q in (30, 70, 95, 100):
            for extra in ({}, {"restart_marker_blocks": 5}):
                data = jc.encode(img, quality=q, subsampling=subsampling, **extra)
                rc, ours = jpeg_check.decode(data)
                assert rc == 0, (kind, q, extra)
                assert np.array_equal(ours, jc.pil_rgb(data)), (kind, q, extra)


def test_optimised_huffman_tables_and_grey_files():
    img = jc.histology_like(96, 80, 3)
    data = jc.encode(img, quality=70, optimize=True)           # per-file Huffman tables, codes longer than 9 bits
    rc, ours = jpeg_check.decode(data)
    assert rc == 0 and np.array_equal(ours, jc.pil_rgb(data))
    b = io.BytesIO()
    Image.fromarray(img).convert("L").save(b, format="JPEG", quality=70)
    rc, ours = jpeg_check.decode(b.getvalue())
    assert rc == 0 and np.array_equal(ours, jc.pil_rgb(b.getvalue()))        # grey -> replicated, as convert("RGB")


def test_reference_patch_format_224_q70():
    """The files of the reference pipeline: 224x224, quality 70, PIL defaults (deepzoom_tiler.py)."""
    for data in jc.patch_files(8):
        rc, ours = jpeg_check.decode(data)
        assert rc == 0 and np.array_equal(ours, jc.pil_rgb(data))


def _parse(files):
    from dsmil_wsi_b200 import jpeg
    return jpeg.parse_batch(files)


def test_parse_batch_geometry_and_offsets():
    files = jc.patch_files(5, 64, 48)
    pb = _parse(files)
    assert pb.n == 5 and pb.bad == 0 and (pb.H, pb.W) == (64, 48)
    assert pb.offsets.tolist() == np.concatenate([[0], np.cumsum([len(f) for f in files])]).tolist()
    assert pb.statuses.tolist() == [0] * 5
    assert bytes(pb.blob.numpy()[pb.offsets[2]:pb.offsets[3]]) == files[2]


def test_parse_flags_what_the_device_path_does_not_take():
    img = jc.histology_like(32, 32, 0)
    ok = jc.encode(img, quality=70)
    progressive = jc.encode(img, quality=70, progressive=True)
    b = io.BytesIO()
    Image.fromarray(img).convert("CMYK").save(b, format="JPEG")
    cmyk = b.getvalue()
    truncated_header = ok[:100]
    not_jpeg = b"\x89PNG\r\n\x1a\n" + bytes(64)
    other_size = jc.encode(jc.histology_like(40, 32, 0), quality=70)
    pb = _parse([ok, progressive, cmyk, truncated_header, not_jpeg])
    assert pb.statuses.tolist() == [0, -2, -2, -1, -1] and pb.bad == 4
    pb = _parse([ok, other_size])
    assert pb.statuses.tolist() == [0, 0] and pb.bad == 1        # decodable, but not one batch geometry


def test_parse_empty_batch():
    pb = _parse([])
    assert pb.n == 0 and pb.bad == 0


def test_parse_paths_reads_files_natively(tmp_path):
    from dsmil_wsi_b200 import jpeg
    files = jc.patch_files(9, 48, 40)
    names = []
    for i, data in enumerate(files):
        p = tmp_path / f"{i}_{i}.jpeg"
        p.write_bytes(data)
        names.append(str(p))
    for threads in (1, 3, 16):
        pb = jpeg.parse_paths(names, threads=threads)
        ref = jpeg.parse_batch(files)
        assert pb.bad == 0 and (pb.H, pb.W) == (48, 40) and pb.offsets.tolist() == ref.offsets.tolist()
        assert bytes(pb.blob.numpy()[:pb.blob_bytes]) == b"".join(files)
        assert pb.file_bytes(4) == files[4]
    with pytest.raises(FileNotFoundError):
        jpeg.parse_paths(names[:2] + [str(tmp_path / "missing.jpeg")])
    assert jpeg.parse_paths([]).n == 0


def test_mutated_files_never_crash_the_shared_decoder():
    """Random damage to header and entropy data: a status or pixels, never a crash (the ASAN/UBSAN harness
    tools/jpeg_fuzz.c ran 450 000 such cases clean; this keeps a small sample in the suite)."""
    rng = np.random.default_rng(7)
    seeds = [jc.encode(jc.histology_like(40, 56, 1), quality=70),
             jc.encode(jc.noise(33, 17, 2), quality=90, subsampling=0, restart_marker_blocks=3)]
    seen = set()
    for it in range(600):
        b = bytearray(seeds[it % 2])
        if it % 4 == 0:
            b = b[: int(rng.integers(1, len(b)))]
        span = min(len(b), 700) if it % 3 == 0 else len(b)
        for _ in range(int(rng.integers(0, 6))):
            b[int(rng.integers(0, span))] = int(rng.choice([0, 255, int(rng.integers(0, 256))]))
        rc, out = jpeg_check.decode(bytes(b)) if _small(bytes(b)) else (-2, None)
        assert rc in (0, -1, -2)
        seen.add(rc)
    assert seen == {0, -1, -2}


def _small(data):
    import ctypes as C
    lib = jpeg_check.load()
    w, h, n = C.c_int32(), C.c_int32(), C.c_int32()
    rc = lib.jpegcheck_size(data, len(data), w, h, n)
    return rc != 0 or w.value * h.value <= 1 << 22
