import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import zstdmt_b200 as z, _oracle as o
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048575
first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if os.environ.get("WARM"):
    w = z.gen_stream(z.GEN_MIX, 65536, 1 << 20, first=first)
    c0 = z.Lz4DeviceCompressor(65536, 1 << 20); c0.run(torch.from_numpy(w).cuda()); torch.cuda.synchronize(); print("warm ok")
src = z.gen_stream(z.GEN_MIX, n, 1 << 20, first=first)
d_in = torch.from_numpy(src).cuda()
comp = z.Lz4DeviceCompressor(n, 1 << 20)
out, foff = comp.run(d_in)
torch.cuda.synchronize()
f = out[: int(foff[-1])].cpu().numpy()
e = o.orc_encode_lz4(src, 1 << 20)
print("sizes", f.size, e.size, "equal", f.size == e.size and np.array_equal(f, e))
if os.environ.get("TIME"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2): comp.run(d_in)
    e0.record()
    for _ in range(5): comp.run(d_in)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("ms/step %.3f  in GB/s %.1f  in+out GB/s %.1f" % (ms, n / ms / 1e6, (n + f.size) / ms / 1e6))
if f.size == e.size and not np.array_equal(f, e):
    d = np.nonzero(f != e)[0]; print("first diff", d[:5])
