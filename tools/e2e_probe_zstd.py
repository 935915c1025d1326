import sys, os, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import zstdmt_b200 as z
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2
kind = int(sys.argv[2]) if len(sys.argv) > 2 else z.GEN_TEXT
threads = 4; chunk = 1 << 20; n = int(gib * (1 << 30))
torch.cuda.set_device(0)
src = z.gen_stream(kind, n, chunk)
L = z.memio_lib(); cap = z.mt_bound(n, chunk); out = np.empty(cap, np.uint8); st = (ctypes.c_size_t * 5)()
for it in range(3):
    t = time.perf_counter(); rc = L.zmt_zstd_compress_mem(threads, 3, chunk, src.ctypes.data, n, out.ctypes.data, cap, st); dt = time.perf_counter() - t
    print("zstd compress e2e it%d rc=%d %.3fs  in %.2f GB/s  in+out %.2f GB/s" % (it, rc, dt, n / dt / 1e9, (n + st[0]) / dt / 1e9), flush=True)
framed = out[: st[0]].copy(); back = np.empty(n + 16, np.uint8)
for it in range(3):
    t = time.perf_counter(); rc = L.zmt_zstd_decompress_mem(threads, 0, framed.ctypes.data, framed.size, back.ctypes.data, n + 16, st); dt = time.perf_counter() - t
    print("zstd decompress e2e it%d rc=%d %.3fs  out %.2f GB/s  in+out %.2f GB/s" % (it, rc, dt, n / dt / 1e9, (n + framed.size) / dt / 1e9), flush=True)
print("roundtrip ok", np.array_equal(back[:n], src))
