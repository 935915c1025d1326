"""Plain (unframed) .lz4 / .zst streams — what the reference routes to st_decompress (lz4-mt_decompress.c:391-483,
zstd-mt_decompress.c:552-687).  Inputs are produced by the codec libraries the reference links
(liblz4 1.9.4 / libzstd of the image, called directly through ctypes), decoded through LZ4MT_/ZSTDCB_decompressDCtx."""
import ctypes

import numpy as np
import pytest

import zstdmt_b200 as z

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


class LZ4FPrefs(ctypes.Structure):
    _fields_ = [("blockSizeID", ctypes.c_int), ("blockMode", ctypes.c_int), ("contentChecksumFlag", ctypes.c_int), ("frameType", ctypes.c_int),
                ("contentSize", ctypes.c_ulonglong), ("dictID", ctypes.c_uint), ("blockChecksumFlag", ctypes.c_int),
                ("compressionLevel", ctypes.c_int), ("autoFlush", ctypes.c_uint), ("favorDecSpeed", ctypes.c_uint), ("reserved", ctypes.c_uint * 3)]


def lz4f(data, **kw):
    try:
        L = ctypes.CDLL("liblz4.so.1")
    except OSError:
        pytest.skip("liblz4.so.1 not present")
    L.LZ4F_compressFrameBound.restype = ctypes.c_size_t; L.LZ4F_compressFrameBound.argtypes = [ctypes.c_size_t, ctypes.c_void_p]
    L.LZ4F_compressFrame.restype = ctypes.c_size_t
    L.LZ4F_compressFrame.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    p = LZ4FPrefs()
    for k, v in kw.items():
        setattr(p, k, v)
    cap = L.LZ4F_compressFrameBound(data.size, ctypes.byref(p))
    out = np.empty(cap, np.uint8)
    n = L.LZ4F_compressFrame(out.ctypes.data, cap, data.ctypes.data, data.size, ctypes.byref(p))
    assert n < (1 << 62)
    return out[:n].copy()


def zstd1(data, level):
    try:
        L = ctypes.CDLL("libzstd.so.1")
    except OSError:
        pytest.skip("libzstd.so.1 not present")
    L.ZSTD_compressBound.restype = ctypes.c_size_t; L.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    L.ZSTD_compress.restype = ctypes.c_size_t; L.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    cap = L.ZSTD_compressBound(data.size)
    out = np.empty(cap, np.uint8)
    n = L.ZSTD_compress(out.ctypes.data, cap, data.ctypes.data, data.size, level)
    assert n < (1 << 62)
    return out[:n].copy()


@pytest.mark.parametrize("prefs", [dict(), dict(blockMode=1, contentSize=1, contentChecksumFlag=1), dict(blockSizeID=7, blockMode=1),
                                   dict(blockSizeID=5, contentChecksumFlag=1, blockChecksumFlag=1), dict(compressionLevel=9, blockSizeID=6)])
def test_plain_lz4_frame(gpu, prefs):
    src = z.gen_stream(z.GEN_MIX, (5 << 20) + 321, 1 << 20)
    if prefs.get("contentSize"):
        prefs = dict(prefs, contentSize=int(src.size))
    frame = lz4f(src, **prefs)
    rc, back, st = z.decompress_mem(z.CODEC_LZ4, frame, src.size + 16, threads=4)
    assert rc == 0, z.lib().LZ4MT_getErrorString(rc)
    assert back.size == src.size and np.array_equal(back, src)
    assert st["insize"] == frame.size and st["outsize"] == src.size


def test_plain_lz4_concatenated_frames_and_skippable(gpu):
    a = z.gen_stream(z.GEN_TEXT, 700000, 1 << 20)
    b = z.gen_stream(z.GEN_MIX, 300000, 1 << 20, first=5)
    skip = np.frombuffer(bytes.fromhex("512a4d18") + (5).to_bytes(4, "little") + b"hello", np.uint8)     # user skippable frame
    stream = np.concatenate([lz4f(a), skip, lz4f(b, blockMode=1, contentChecksumFlag=1), lz4f(np.zeros(0, np.uint8))])
    rc, back, st = z.decompress_mem(z.CODEC_LZ4, stream, a.size + b.size + 16, threads=2)
    assert rc == 0 and np.array_equal(back, np.concatenate([a, b]))
    bad = stream.copy(); bad[-1] ^= 0xFF; bad = np.concatenate([stream, np.frombuffer(b"\x01\x02\x03\x04\x05", np.uint8)])
    rc, _, _ = z.decompress_mem(z.CODEC_LZ4, bad, a.size + b.size + 16)
    assert z.lib().LZ4MT_isError(rc)                                        # trailing garbage is not silently accepted


@pytest.mark.parametrize("level", [1, 3, 12, 19])
def test_plain_zstd_frames(gpu, level):
    a = z.gen_stream(z.GEN_MIX, (6 << 20) + 17, 1 << 20)           # one frame, window > 1 MiB, 128 KiB blocks
    b = z.gen_stream(z.GEN_TEXT, 1 << 20, 1 << 20)
    stream = np.concatenate([zstd1(a, level), zstd1(b, level), zstd1(np.zeros(0, np.uint8), level)])
    rc, back, st = z.decompress_mem(z.CODEC_ZSTD, stream, a.size + b.size + 16, threads=4)
    assert rc == 0, z.lib().ZSTDCB_getErrorString(rc)
    assert np.array_equal(back, np.concatenate([a, b]))
    assert st["outsize"] == a.size + b.size


def test_plain_zstd_small_and_empty(gpu):
    for n in (0, 1, 5, 200):
        src = z.gen_stream(z.GEN_TEXT, n, 1 << 20)
        fr = zstd1(src, 3)
        rc, back, st = z.decompress_mem(z.CODEC_ZSTD, fr, n + 16)
        assert rc == 0 and np.array_equal(back, src), n
