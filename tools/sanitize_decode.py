"""Small decode workloads for compute-sanitizer (memcheck / racecheck): LZ4 own + reference-framed (linked) streams, zstd own +
libzstd-framed streams, through the device API and the callback API.  python tools/sanitize_decode.py [MiB]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import zstdmt_b200 as z, _oracle as o
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 6
chunk, n = 1 << 20, (mib << 20) + 12345
src = z.gen_stream(z.GEN_MIX, n, chunk)
d_src = torch.from_numpy(src).cuda()
outs = [min(chunk, n - i * chunk) for i in range(-(-n // chunk))]
def lz4_dec(framed, label):
    offs, sizes = z.scan_frames(framed)
    dec = z.Lz4DeviceDecompressor(offs, sizes, outs)
    out, st = dec.run(torch.from_numpy(framed).cuda()); torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0 and torch.equal(out[:n], d_src), label
    print(label, "ok", flush=True)
def zstd_dec(framed, label):
    dec = z.ZstdDeviceDecompressor(framed)
    out, st = dec.run(torch.from_numpy(framed).cuda()); torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0 and torch.equal(out[:n], d_src), label
    print(label, "ok", flush=True)
c = z.Lz4DeviceCompressor(n, chunk); out, foff = c.run(d_src); torch.cuda.synchronize()
lz4_dec(out[: int(foff[-1])].cpu().numpy(), "lz4 own stream")
rc, rf, _ = o.ref_compress(o.CODEC_LZ4, src, threads=4, level=1, chunk=chunk); assert rc == 0
lz4_dec(rf, "lz4 reference-framed (linked blocks)")
zc = z.ZstdDeviceCompressor(n, chunk); zout, zf = zc.run(d_src); torch.cuda.synchronize()
zstd_dec(zout[: int(zf[-1])].cpu().numpy(), "zstd own stream")
rc, zrf, _ = o.ref_compress(o.CODEC_ZSTD, src, threads=4, level=3, chunk=chunk); assert rc == 0
zstd_dec(zrf, "zstd reference-framed (libzstd level 3)")
for codec, fr in ((z.CODEC_LZ4, rf), (z.CODEC_ZSTD, zrf)):
    rc, back, st = z.decompress_mem(codec, fr, n + 16, threads=4)
    assert rc == 0 and np.array_equal(back, src)
print("callback API ok")
