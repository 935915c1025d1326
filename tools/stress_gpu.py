"""Randomised parity stress on the GPU (not part of the pytest suite: run for a time budget).
usage: stress_gpu.py <seconds> [seed]
Each round draws a data recipe (generator classes, random / repeated / zero spans glued at random places), a total
size and a chunk size, then checks:
  LZ4 : GPU frames == oracle twin, byte for byte; GPU decode and the oracle decoder restore the input
  zstd: GPU frames decode with the oracle's RFC 8878 decoder, with the real libzstd (when oracle/_ref is built) and on the GPU
  API : every third round, {LZ4MT,ZSTDCB}_compressCCtx / _decompressDCtx through in-memory callbacks with a random thread count:
        same bytes as the device path, round trip, and streams made by the real reference (random level) decode through our API"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import zstdmt_b200 as z, _oracle as o

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
if torch.cuda.is_available(): torch.cuda.set_device(0)

def recipe(n):
    parts = []; left = n
    while left > 0:
        k = int(rng.integers(0, 6)); m = int(min(left, rng.choice([1, 7, 100, 4096, 65536, 300000, 1 << 20]) * rng.integers(1, 4)))
        if k == 0: p = rng.integers(0, 256, m, dtype=np.uint8)
        elif k == 1: p = np.zeros(m, np.uint8)
        elif k == 2: p = np.resize(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8), m)
        elif k == 3: p = z.gen_stream(z.GEN_TEXT, m, 1 << 20, first=int(rng.integers(0, 50)))
        elif k == 4: p = z.gen_stream(z.GEN_MIX, m, 1 << 20, first=int(rng.integers(0, 50)))
        else:
            if parts:
                cur = np.concatenate(parts); back = int(rng.integers(1, min(cur.size, 70000) + 1)); p = np.resize(cur[-back:], m)
            else: p = rng.integers(0, 4, m, dtype=np.uint8)
        parts.append(p); left -= m
    return np.concatenate(parts)[:n]

t_end = time.time() + budget; rounds = 0; nbytes = 0
while time.time() < t_end:
    n = int(rng.choice([0, 1, 13, 5000, 65536, 65537, 1 << 20, (3 << 20) + 17, (9 << 20) + 1, 20 << 20]))
    n = max(0, n + int(rng.integers(-3, 4)) if n > 3 else n)
    chunk = int(rng.choice([65536, 100000, 1 << 20, 4 << 20, (1 << 20) + 4096]))
    src = recipe(n) if n else np.empty(0, np.uint8)
    d_in = torch.from_numpy(src).cuda() if n else torch.empty(1, dtype=torch.uint8, device="cuda")
    # ---- LZ4
    comp = z.Lz4DeviceCompressor(n, chunk); out, foff = comp.run(d_in); torch.cuda.synchronize()
    f = out[: int(foff[-1])].cpu().numpy(); e = o.orc_encode_lz4(src, chunk)
    assert f.size == e.size and np.array_equal(f, e), ("lz4 twin mismatch", n, chunk, seed, rounds)
    rc, back = o.orc_decode(o.CODEC_LZ4, f, n); assert rc == 0 and np.array_equal(back, src), ("lz4 oracle decode", n, chunk)
    offs, sizes = z.scan_frames(f)
    osz = [min(chunk, n - i * chunk) for i in range(max(1, -(-n // chunk)))]
    dec = z.Lz4DeviceDecompressor(offs, sizes, osz); dout, st = dec.run(torch.from_numpy(np.ascontiguousarray(f)).cuda()); torch.cuda.synchronize()
    assert int(st.abs().sum().item()) == 0 and np.array_equal(dout[: dec.out_total].cpu().numpy(), src), ("lz4 gpu decode", n, chunk)
    # ---- zstd
    zc = z.ZstdDeviceCompressor(n, chunk); zo, zf = zc.run(d_in); torch.cuda.synchronize()
    g = zo[: int(zf[-1])].cpu().numpy()
    rc, back = o.orc_decode(o.CODEC_ZSTD, g, n); assert rc == 0 and np.array_equal(back, src), ("zstd oracle decode", n, chunk, seed, rounds)
    if o.have_ref():
        rc, b2, _ = o.ref_decompress(o.CODEC_ZSTD, g, n, threads=2); assert rc == 0 and np.array_equal(b2, src), ("libzstd decode", n, chunk)
    zd = z.ZstdDeviceDecompressor(g); zout, zst = zd.run(torch.from_numpy(g).cuda() if g.size else torch.empty(1, dtype=torch.uint8, device="cuda")); torch.cuda.synchronize()
    assert zd.out_total == n and int(zst.abs().sum().item()) == 0 and np.array_equal(zout[:n].cpu().numpy(), src), ("zstd gpu decode", n, chunk)
    # ---- callback API (host pipeline): same bytes as the device path, round trip, and reference-made streams
    if n <= (24 << 20) and rounds % 3 == 0:
        T = int(rng.integers(1, 9))
        for codec, dev_bytes in ((z.CODEC_LZ4, f), (z.CODEC_ZSTD, g)):
            rc, fr, _ = z.compress_mem(codec, src, threads=T, level=int(rng.integers(1, 4)), chunk=chunk)
            assert rc == 0 and fr.size == dev_bytes.size and np.array_equal(fr, dev_bytes), ("api compress != device path", codec, n, chunk, T)
            rc, bk, _ = z.decompress_mem(codec, fr, n + 16, threads=T)
            assert rc == 0 and bk.size == n and np.array_equal(bk, src), ("api decompress", codec, n, chunk, T)
            if o.have_ref() and n:
                rc, rf, _ = o.ref_compress(codec, src, threads=int(rng.integers(1, 5)), level=int(rng.choice([1, 3, 9])), chunk=chunk)
                assert rc == 0
                rc, bk, _ = z.decompress_mem(codec, rf, n + 16, threads=T)
                assert rc == 0 and bk.size == n and np.array_equal(bk, src), ("api decompress of reference stream", codec, n, chunk, T)
    rounds += 1; nbytes += n
print("stress ok: %d rounds, %.1f MiB, seed %d" % (rounds, nbytes / 2**20, seed))
