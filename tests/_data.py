"""Synthetic inputs shared by the CPU and GPU suites."""
import numpy as np


def long_runs_stream(seed):
    """Blocks built to hit the length-field corner cases of the LZ4 block format: literal runs of 14/15/269/270/271
    and thousands of bytes (255-runs inside the literal length), matches of 18/19/1289/1290 and tens of thousands of
    bytes (255-runs inside the match length), long runs ending a block and a sequence pending across tiles."""
    rng = np.random.default_rng(seed)
    parts = []
    def rnd(n): parts.append(rng.integers(0, 256, n, dtype=np.uint8))
    def rep(back, n):
        cur = np.concatenate(parts); parts.clear(); parts.append(cur)
        src = cur[-back:]; parts.append(np.resize(src, n))
    rnd(5000)
    for lit, back, ml in [(14, 3000, 18), (15, 2000, 19), (269, 4000, 1289), (270, 100, 1290), (271, 1, 1291), (3000, 4500, 40000),
                          (33, 7, 5000), (32, 2, 300), (700, 9000, 15 + 255 * 5 + 4), (0, 16, 70000), (20000, 64, 9)]:
        rnd(lit); rep(back, ml)
    rnd(12345)
    rep(1, 200000)          # a run crossing several blocks: the last sequence of each block ends at the literal tail
    rnd(40000)
    return np.concatenate(parts)
