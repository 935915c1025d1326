"""Localise a mismatch of the zstdmt-style / reference-framed zstd decode: first differing byte, per-frame status."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import zstdmt_b200 as z
from test_gpu_plain_streams import zstd1
n, chunk = (3 << 20) + 77, 1 << 20
src = z.gen_stream(z.GEN_MIX, n, chunk)
frames = [zstd1(src[o:o + chunk], 3) for o in range(0, n, chunk)]
hdr = lambda f: np.frombuffer((0x184D2A50).to_bytes(4, "little") + (4).to_bytes(4, "little") + int(f.size).to_bytes(4, "little"), np.uint8)
framed = np.concatenate([x for f in frames for x in (hdr(f), f)])
for it in range(6):
    dec = z.ZstdDeviceDecompressor(framed)
    out, st = dec.run(torch.from_numpy(framed).cuda()); torch.cuda.synchronize()
    o = out[:n].cpu().numpy()
    bad = np.nonzero(o != src)[0]
    print("device it%d status %s sizes %s flags %s first-bad %s nbad %d" % (it, st.cpu().tolist(), dec.out_size.cpu().tolist(), dec.d_fseq.cpu().tolist(), bad[:3].tolist(), bad.size), flush=True)
stream = np.concatenate([np.frombuffer(bytes.fromhex("28b52ffd2000010000"), np.uint8), framed])
for it in range(4):
    rc, back, stt = z.decompress_mem(z.CODEC_ZSTD, stream, n + 16, threads=4)
    bad = np.nonzero(back[:min(back.size, n)] != src[:min(back.size, n)])[0] if back.size else np.zeros(0)
    print("callback it%d rc %d size %d frames %d first-bad %s" % (it, rc, back.size, stt["frames"], bad[:3].tolist()), flush=True)
