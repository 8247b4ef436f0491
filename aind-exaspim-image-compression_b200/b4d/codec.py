"""Chunking + byte shuffle ahead of the chunk codec (first step).

The reference measures compression with ``compute_cratio(img, codec)``
(utils/img_util.py:401-441): the uint16 volume is cut into a C-order grid of 64^3
pieces (ragged at the far faces), each piece is made contiguous and handed to
``Blosc(cname="zstd", clevel=6, shuffle=SHUFFLE)`` (evaluate.py:40, :132; train.py:105).
This module does the part of that which is pure byte movement on the device —
the gather of every piece and Blosc's 2-byte shuffle (kernel K9,
``b4d_chunk_shuffle_u16``) — and returns, per piece, the 256-bin counts of its
low-byte and high-byte planes, from which an order-0 entropy bound of the
compressed size follows without touching the host codec.

The entropy coder itself stays on the host.  ``numcodecs`` (the reference's
codec) is not installed here, so Blosc frames cannot be produced or checked;
when the system's ``libzstd.so.1`` is loadable, ``compute_cratio`` compresses
the shuffled pieces with plain zstd at the same level and adds Blosc's fixed
framing (16-byte header + one 4-byte block offset; at clevel 6 c-blosc's block
size for zstd is 512 KiB, i.e. one block per 64^3 uint16 piece).  That figure
is labelled an APPROXIMATION of the reference's cratio wherever it is printed.
"""
import ctypes
import ctypes.util

import numpy as np

BLOSC_HEADER = 16  # bytes: version, versionlz, flags, typesize, nbytes, blocksize, cbytes
BLOSC_BSTART = 4   # one int32 offset per block


def chunk_grid(shape, chunk=(64, 64, 64)):
    """Pieces of the compute_cratio loop in its own order (img_util.py:427-431):
    list of ((z0, y0, x0), (dz, dy, dx), byte_offset) — byte_offset into the shuffled stream."""
    shape = tuple(int(s) for s in shape)
    chunk = tuple(int(c) for c in chunk)
    out, pos = [], 0
    for z0 in range(0, shape[0], chunk[0]):
        for y0 in range(0, shape[1], chunk[1]):
            for x0 in range(0, shape[2], chunk[2]):
                d = (min(chunk[0], shape[0] - z0), min(chunk[1], shape[1] - y0), min(chunk[2], shape[2] - x0))
                out.append(((z0, y0, x0), d, pos))
                pos += 2 * d[0] * d[1] * d[2]
    return out


def chunk_shuffle(x, chunk=(64, 64, 64), want_bytes=True, want_hist=True, device=None):
    """uint16 volume -> (shuffled bytes of all pieces back to back, hist[pieces, 2, 256]).
    NumPy in -> NumPy out; a CUDA torch tensor in -> CUDA tensors out (nothing crosses PCIe)."""
    from .api import get_denoiser

    return get_denoiser(device).chunk_shuffle(x, chunk, want_bytes, want_hist)


def unshuffle_piece(buf, dims):
    """Inverse of the 2-byte shuffle for one piece: bytes -> uint16 array of shape `dims`."""
    ne = int(dims[0]) * int(dims[1]) * int(dims[2])
    b = np.frombuffer(buf, dtype=np.uint8, count=2 * ne)
    return (b[:ne].astype(np.uint16) | (b[ne:].astype(np.uint16) << 8)).reshape(dims)


def entropy_bytes(hist):
    """Order-0 entropy bound per piece, in bytes: sum over the two byte planes of n * H(plane) / 8.
    float64 on the host from the integer counts (identical for every implementation of the counts)."""
    h = np.asarray(hist, dtype=np.float64).reshape(-1, 2, 256)
    n = h.sum(axis=2, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        p = np.where(h > 0, h / n, 1.0)
        bits = -(h * np.log2(p)).sum(axis=2)
    return bits.sum(axis=1) / 8.0


def estimate_cratio(hist):
    """Uncompressed bytes / entropy bound, over all pieces (an optimistic order-0 model:
    a real coder pays table overhead but also exploits runs in the high plane)."""
    h = np.asarray(hist).reshape(-1, 2, 256)
    raw = float(h.sum())
    est = float(np.maximum(entropy_bytes(h), 1.0).sum())
    return raw / est


# ---------------------------------------------------------------- host zstd ----
_zstd = None


def _load_zstd():
    global _zstd
    if _zstd is None:
        name = ctypes.util.find_library("zstd") or "libzstd.so.1"
        try:
            lib = ctypes.CDLL(name)
            lib.ZSTD_compressBound.restype = ctypes.c_size_t
            lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
            lib.ZSTD_compress.restype = ctypes.c_size_t
            lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
            lib.ZSTD_decompress.restype = ctypes.c_size_t
            lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
            lib.ZSTD_isError.restype = ctypes.c_uint
            lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
            _zstd = lib
        except OSError:
            _zstd = False
    return _zstd or None


def zstd_available():
    return _load_zstd() is not None


def zstd_compress(buf, level=6):
    lib = _load_zstd()
    if lib is None:
        raise RuntimeError("libzstd is not loadable on this host")
    src = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8))
    cap = lib.ZSTD_compressBound(src.size)
    dst = np.empty(cap, dtype=np.uint8)
    n = lib.ZSTD_compress(dst.ctypes.data, cap, src.ctypes.data, src.size, int(level))
    if lib.ZSTD_isError(n):
        raise RuntimeError("ZSTD_compress failed")
    return dst[:n]


def zstd_decompress(buf, nbytes):
    lib = _load_zstd()
    if lib is None:
        raise RuntimeError("libzstd is not loadable on this host")
    src = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8))
    dst = np.empty(int(nbytes), dtype=np.uint8)
    n = lib.ZSTD_decompress(dst.ctypes.data, dst.size, src.ctypes.data, src.size)
    if lib.ZSTD_isError(n) or n != dst.size:
        raise RuntimeError("ZSTD_decompress failed")
    return dst


def compute_cratio(img, codec=None, patch_shape=(64, 64, 64), clevel=6, device=None, sample_every=1):
    """Chunked compression ratio in the shape of the reference's ``compute_cratio``
    (img_util.py:401-441): total uncompressed / total compressed over the pieces, 2 decimals.

    codec given  -> every piece (made contiguous, unshuffled) goes through ``codec.encode`` exactly
                    as in the reference; nothing runs on the device.
    codec None   -> pieces are gathered and byte-shuffled on the device, zstd-compressed at `clevel`
                    on the host, plus Blosc's fixed framing — an APPROXIMATION of the reference's
                    Blosc(zstd, SHUFFLE) figure (module docstring).  `sample_every` > 1 compresses
                    every k-th piece only (ratio over the sampled pieces).
    """
    if hasattr(img, "ndim") and img.ndim == 5:  # img_util.py:420-421
        img = img[0, 0]
    img = np.ascontiguousarray(img, dtype=np.uint16)
    grid = chunk_grid(img.shape, patch_shape)
    total_c = total_u = 0
    if codec is not None:
        for (z0, y0, x0), (dz, dy, dx), _ in grid:
            piece = np.ascontiguousarray(img[z0 : z0 + dz, y0 : y0 + dy, x0 : x0 + dx])
            total_c += len(codec.encode(piece))
            total_u += piece.nbytes
        return round(total_u / total_c, 2)
    shuffled, _ = chunk_shuffle(img, patch_shape, want_bytes=True, want_hist=False, device=device)
    for i, (_, (dz, dy, dx), pos) in enumerate(grid):
        if i % int(sample_every):
            continue
        nb = 2 * dz * dy * dx
        total_c += zstd_compress(shuffled[pos : pos + nb], clevel).size + BLOSC_HEADER + BLOSC_BSTART
        total_u += nb
    return round(total_u / total_c, 2)
