"""ctypes bindings to the TEST-ONLY oracle: oracle/liboracle.so (CPU restatement) and
oracle/_ref/libzstdmt_ref.so (the unmodified reference wrapper + liblz4/libzstd)."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
c_sz = ctypes.c_size_t
c_vp = ctypes.c_void_p

CODEC_LZ4, CODEC_ZSTD = 1, 2
_orc = None
_ref = None


def orc():
    global _orc
    if _orc is None:
        L = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        L.orc_xxh32.restype = ctypes.c_uint32; L.orc_xxh32.argtypes = [c_vp, c_sz, ctypes.c_uint32]
        L.orc_mt_decode.restype = ctypes.c_int; L.orc_mt_decode.argtypes = [ctypes.c_int, c_vp, c_sz, c_vp, c_sz, ctypes.POINTER(c_sz)]
        L.orc_mt_encode_lz4_b200.restype = c_sz; L.orc_mt_encode_lz4_b200.argtypes = [c_vp, c_sz, c_sz, c_vp, c_sz]
        L.orc_lz4f_decode.restype = ctypes.c_int; L.orc_lz4f_decode.argtypes = [c_vp, c_sz, c_vp, c_sz, ctypes.POINTER(c_sz), ctypes.POINTER(c_sz)]
        L.orc_zstd_decode.restype = ctypes.c_int; L.orc_zstd_decode.argtypes = [c_vp, c_sz, c_vp, c_sz, ctypes.POINTER(c_sz), ctypes.POINTER(c_sz)]
        L.orc_lz4_block_compress_b200.restype = c_sz; L.orc_lz4_block_compress_b200.argtypes = [c_vp, c_sz, c_vp, c_sz]
        L.orc_lz4_block_bound.restype = c_sz; L.orc_lz4_block_bound.argtypes = [c_sz]
        L.orc_lz4_block_decode.restype = ctypes.c_long; L.orc_lz4_block_decode.argtypes = [c_vp, c_sz, c_vp, c_sz, c_sz]
        L.orc_mt_scan.restype = ctypes.c_long; L.orc_mt_scan.argtypes = [c_vp, c_sz, c_vp, c_vp, c_sz]
        _orc = L
    return _orc


def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libzstdmt_ref.so"))


def ref():
    global _ref
    if _ref is None:
        L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libzstdmt_ref.so"))
        for f in ("ref_lz4_compress_mem", "ref_zstd_compress_mem"):
            getattr(L, f).restype = c_sz
            getattr(L, f).argtypes = [ctypes.c_int] * 3 + [c_vp, c_sz, c_vp, c_sz, ctypes.POINTER(c_sz)]
        for f in ("ref_lz4_decompress_mem", "ref_zstd_decompress_mem"):
            getattr(L, f).restype = c_sz
            getattr(L, f).argtypes = [ctypes.c_int] * 2 + [c_vp, c_sz, c_vp, c_sz, ctypes.POINTER(c_sz)]
        _ref = L
    return _ref


def _arr(data):
    return np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data, dtype=np.uint8)


def ref_compress(codec, data, threads=1, level=1, chunk=1 << 20):
    """The real reference: {LZ4MT,ZSTDCB}_compressCCtx -> liblz4 / libzstd."""
    data = _arr(data)
    cap = data.size + data.size // 64 + 65536 + 64 * (data.size // chunk + 2)
    out = np.empty(cap, np.uint8); st = (c_sz * 5)()
    fn = ref().ref_lz4_compress_mem if codec == CODEC_LZ4 else ref().ref_zstd_compress_mem
    rc = fn(threads, level, chunk, data.ctypes.data, data.size, out.ctypes.data, cap, st)
    return rc, out[: st[0]].copy(), list(st)


def ref_decompress(codec, data, out_cap, threads=1):
    data = _arr(data)
    out = np.empty(out_cap + 1, np.uint8); st = (c_sz * 5)()
    fn = ref().ref_lz4_decompress_mem if codec == CODEC_LZ4 else ref().ref_zstd_decompress_mem
    rc = fn(threads, 0, data.ctypes.data, data.size, out.ctypes.data, out_cap + 1, st)
    return rc, out[: st[0]].copy(), list(st)


def orc_decode(codec, data, out_cap):
    data = _arr(data)
    out = np.empty(out_cap + 1, np.uint8); got = c_sz(0)
    rc = orc().orc_mt_decode(codec, data.ctypes.data, data.size, out.ctypes.data, out_cap + 1, ctypes.byref(got))
    return rc, out[: got.value]


def orc_encode_lz4(data, chunk=1 << 20):
    data = _arr(data)
    cap = data.size + data.size // 100 + 4096 + 64 * (data.size // chunk + 2)
    out = np.empty(cap, np.uint8)
    n = orc().orc_mt_encode_lz4_b200(data.ctypes.data, data.size, chunk, out.ctypes.data, cap)
    return out[:n].copy()


def xxh32(data, seed=0):
    data = _arr(data)
    return orc().orc_xxh32(data.ctypes.data, data.size, seed)
