"""Device-timed zstd decode of (a) the GPU encoder's own stream and (b) a stream framed by the reference (libzstd level 3,
frame-sequential blocks) — python tools/zstd_decode_probe.py [GiB] [kind: text|mix]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import zstdmt_b200 as z, _oracle as o
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
kind = z.GEN_TEXT if (len(sys.argv) < 3 or sys.argv[2] == "text") else z.GEN_MIX
chunk, n = 1 << 20, int(gib * (1 << 30))
src = z.gen_stream(kind, n, chunk)
d_src = torch.from_numpy(src).cuda()
def timed(dec, d_framed, total, label):
    out, st = dec.run(d_framed); torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0 and torch.equal(out[:n], d_src), label
    for _ in range(2): dec.run(d_framed)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): dec.run(d_framed)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%s: %.2f ms, %.1f GB/s in+out (%.1f GB/s out), ratio %.3f" % (label, ms, total / ms / 1e6, n / ms / 1e6, n / (total - n)), flush=True)
zc = z.ZstdDeviceCompressor(n, chunk); zout, zf = zc.run(d_src); torch.cuda.synchronize()
fr = int(zf[-1].item()); host = zout[:fr].cpu().numpy()
timed(z.ZstdDeviceDecompressor(host), zout, n + fr, "own stream")
if o.have_ref():
    rc, rf, _ = o.ref_compress(o.CODEC_ZSTD, src, threads=min(os.cpu_count() or 1, 64), level=3, chunk=chunk)
    assert rc == 0
    timed(z.ZstdDeviceDecompressor(rf), torch.from_numpy(rf).cuda(), n + rf.size, "reference-framed (libzstd level 3)")
