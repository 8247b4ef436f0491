"""Chunking + byte shuffle ahead of the chunk codec (first step): the host-side
helpers and the two CPU restatements (NumPy, following compute_cratio's own loop —
utils/img_util.py:427-438 — and the C one behind the oracle's C ABI) against each other."""
import ctypes

import numpy as np
import pytest

SHAPES = [((64, 64, 64), (64, 64, 64)), ((70, 66, 130), (64, 64, 64)), ((20, 33, 35), (8, 16, 4)), ((5, 3, 2), (64, 64, 64))]


def _c_restatement(oracle_lib, vol, chunk):
    lib = oracle_lib.load()
    npieces = int(np.prod([-(-s // c) for s, c in zip(vol.shape, chunk)]))
    out = np.empty(2 * vol.size, np.uint8)
    hist = np.empty((npieces, 2, 256), np.uint32)
    i3 = ctypes.c_int64 * 3
    rc = lib.b4d_chunk_shuffle_u16(None, ctypes.c_void_p(vol.ctypes.data), i3(*vol.shape), i3(*chunk),
                                   ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(hist.ctypes.data), 0, 0)
    assert rc == 0
    return out, hist


@pytest.mark.parametrize("shape,chunk", SHAPES)
def test_two_restatements_agree_and_the_shuffle_inverts(shape, chunk, oracle_lib):
    from b4d import codec

    rng = np.random.default_rng(sum(shape))
    vol = rng.integers(0, 65536, shape, dtype=np.uint16)
    by, hist = oracle_lib.chunk_shuffle_reference(vol, chunk)
    cby, chist = _c_restatement(oracle_lib, vol, chunk)
    assert np.array_equal(by, cby) and np.array_equal(hist, chist)
    grid = codec.chunk_grid(shape, chunk)
    assert len(grid) == hist.shape[0] and by.size == 2 * vol.size
    back = np.zeros_like(vol)
    for (z0, y0, x0), d, pos in grid:  # every piece sits where chunk_grid says and unshuffles to the slice
        back[z0 : z0 + d[0], y0 : y0 + d[1], x0 : x0 + d[2]] = codec.unshuffle_piece(by[pos:], d)
    assert np.array_equal(back, vol)
    assert int(hist.sum()) == 2 * vol.size


def test_shuffle_layout_by_hand(oracle_lib):
    vol = np.array([[[0x0102, 0x0304], [0x0506, 0x0708]]], np.uint16)  # one piece, C order
    by, hist = oracle_lib.chunk_shuffle_reference(vol, (64, 64, 64))
    assert by.tolist() == [0x02, 0x04, 0x06, 0x08, 0x01, 0x03, 0x05, 0x07]  # low bytes, then high bytes
    assert hist[0, 0, 2] == 1 and hist[0, 1, 7] == 1 and hist.sum() == 8


def test_entropy_bound():
    from b4d import codec

    h = np.zeros((3, 2, 256), np.uint32)
    h[0, 0, 7] = 1000  # constant planes: 0 bits
    h[0, 1, 0] = 1000
    h[1, 0, :] = 4  # uniform low plane: 8 bits per byte; constant high plane
    h[1, 1, 3] = 1024
    h[2, 0, :2] = 512  # two equiprobable values: 1 bit per byte in each plane
    h[2, 1, :2] = 512
    e = codec.entropy_bytes(h)
    assert e[0] == 0.0 and e[1] == 1024.0 and e[2] == 2 * 1024 / 8.0
    assert codec.estimate_cratio(h[1:]) == pytest.approx((2048 + 2048) / (1024.0 + 256.0))


def test_compute_cratio_with_a_codec_object_follows_the_reference_loop():
    """img_util.py:401-441: every piece goes through codec.encode as a contiguous array; ratio rounded to 2."""
    from b4d import codec

    class Half:  # "compresses" every piece to half its bytes and records what it was given
        def __init__(self):
            self.seen = []

        def encode(self, piece):
            assert piece.flags.c_contiguous and piece.dtype == np.uint16
            self.seen.append(piece.shape)
            return bytes(piece.nbytes // 2)

    vol = np.arange(70 * 66 * 65, dtype=np.uint16).reshape(70, 66, 65)
    c = Half()
    assert codec.compute_cratio(vol, c) == 2.0
    assert c.seen == [(64, 64, 64), (64, 64, 1), (64, 2, 64), (64, 2, 1), (6, 64, 64), (6, 64, 1), (6, 2, 64), (6, 2, 1)]
    assert codec.compute_cratio(vol[None, None], Half()) == 2.0  # 5-D input: img[0, 0] (img_util.py:420-421)


def test_host_zstd_round_trip():
    from b4d import codec

    if not codec.zstd_available():
        pytest.skip("libzstd not loadable here")
    rng = np.random.default_rng(0)
    buf = rng.integers(0, 4, 100_000, dtype=np.uint8)
    c = codec.zstd_compress(buf, 6)
    assert c.size < buf.size // 3
    assert np.array_equal(codec.zstd_decompress(c, buf.size), buf)
