"""CPU emulation of lz4_compress_blocks_kernel's speculative-chain parse (debug tool)."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import zstdmt_b200 as z, _oracle as o

TILE, NT, SEG, END = 4096, 256, 16, 0xFFFF

def offsets(blk):
    off = np.zeros(len(blk), np.uint16)
    L = o.orc(); L.orc_lz4_b200_offsets.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]; L.orc_lz4_b200_offsets.restype = None
    L.orc_lz4_b200_offsets(blk.ctypes.data, len(blk), off.ctypes.data)
    return off

def mlen(b, q, c, limit):
    L = 4
    while q + L < limit and b[q + L] == b[c + L]: L += 1
    return L

def greedy(b, off):
    n = len(b); p = 0; seqs = []
    while p < n:
        q = p
        while q < n and off[q] == 0: q += 1
        if q >= n: break
        L = mlen(b, q, q - off[q], n - 5); seqs.append((q, L)); p = q + L
    return seqs

def emulate(b, off):
    n = len(b); limit = n - 5; e = 0; seqs = []
    for t0 in range(0, n, TILE):
        t1 = t0 + TILE
        M = np.zeros(TILE, bool); M[: min(TILE, n - t0)] = off[t0:t1] != 0
        if not M.any() or e >= t1:
            if e < t1: e = t1
            continue
        V = np.zeros(TILE, bool); Sel = np.zeros(TILE, bool); ln = np.zeros(TILE, np.int64)
        link = [0] * NT; mpos = [0] * NT; xfree = [0] * NT; min_ = [0] * NT
        k0 = (e - t0) // SEG

        def walk(mode, k, p):
            while True:
                if p >= t1:
                    if mode == 1: link[k] = END; mpos[k] = p
                    if mode == 0: xfree[k] = p
                    return
                rel = p - t0; j = rel // SEG
                if mode == 0 and j != k: xfree[k] = p; return
                if mode != 0 and j != k and rel % SEG == 0:
                    if mode == 1: link[k] = j; mpos[k] = p
                    return
                seg_end = (j + 1) * SEG
                cand = [r for r in range(rel, seg_end) if M[r]]
                if not cand:
                    nx = t0 + seg_end
                    if mode == 0: xfree[k] = nx; return
                    if j + 1 == NT:
                        if mode == 1: link[k] = END; mpos[k] = t1
                        return
                    if j != k:
                        if mode == 1: link[k] = j + 1; mpos[k] = nx
                        return
                    p = nx; continue
                qr = cand[0]; q = t0 + qr
                if mode != 0 and j != k and V[qr]:
                    if mode == 1: link[k] = j; mpos[k] = q
                    return
                if mode == 2:
                    Sel[qr] = True; L = ln[qr]
                    assert L >= 4, ("len cache missing", t0, k, qr)
                else:
                    L = mlen(b, q, q - int(off[q]), limit); ln[qr] = L
                    if mode == 0: V[qr] = True
                p = q + L

        for k in range(NT):
            if k >= k0: walk(0, k, e if k == k0 else t0 + k * SEG)
            else: link[k] = k
        for k in range(k0, NT): walk(1, k, xfree[k])
        # reachability
        reach = [False] * NT; k = k0; e_next = None
        while True:
            reach[k] = True
            if link[k] == END: e_next = mpos[k]; break
            assert link[k] > k, ("non-forward link", t0, k, link[k])
            min_[link[k]] = mpos[k]; k = link[k]
        min_[k0] = e
        for k in range(NT):
            if reach[k]: walk(2, k, min_[k])
        e = e_next
        for r in np.nonzero(Sel)[0]: seqs.append((t0 + int(r), int(ln[r])))
    return seqs

if __name__ == "__main__":
    n = int(sys.argv[1]); first = int(sys.argv[2]); blkidx = int(sys.argv[3])
    src = z.gen_stream(z.GEN_MIX, n, 1 << 20, first=first)
    b = src[blkidx * 65536: (blkidx + 1) * 65536]
    off = offsets(b)
    g = greedy(b, off); m = emulate(b, off)
    print("greedy", len(g), "emu", len(m), "equal", g == m)
    if g != m:
        for i, (x, y) in enumerate(zip(g, m)):
            if x != y: print("first diff at seq", i, x, y, g[i-2:i+3], m[i-2:i+3]); break
