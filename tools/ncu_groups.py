"""Group warp-stall samples of the pipelined LZ4 compressor by source region (uses tools/ncu_lines.py output).
usage: ncu_groups.py <report.ncu-rep>"""
import os, re, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(here, "..", "zstdmt_b200", "csrc", "lz4_pipe.cuh")).read().splitlines()
def find(s): return next(i + 1 for i, l in enumerate(src) if s in l)
L = {k: find(v) for k, v in dict(exscan="uint32_t team_exscan1(", kern="lz4_blocks_pipe_kernel(const", teamA="if (team == 0) {", rounds="for (uint32_t r = 0; r < C_TILE / C_ROUND",
     anyM="if (lane == 0 && anyM)", walks0="if (alive) c_walk<0>", p3="const uint32_t lk = S.link[ttid];", arriveF="bar_arrive<PB_FULL0 + 1", teamB="back end: mark, merge", waitF="bar_sync<PB_FULL0 + 1",
     mark="c_walk<2>(W", p5="uint32_t np;", flush="flush the pending sequence", final="both teams drained").items()}
env = dict(os.environ, TOP="1000")
out = subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), sys.argv[1], "lz4_blocks_pipe"], capture_output=True, text=True, env=env).stdout
G = {}; tot = 0
for ln in out.splitlines():
    m = re.match(r"\s*(\d+)\s+[\d.]+%\s+inst=\s*(\d+).*?(\S+):(\d+)\s", ln)
    if not m: continue
    s, f, l = int(m.group(1)), m.group(3), int(m.group(4)); tot += s
    if f == "lz4_pipe.cuh":
        if l < L["exscan"]: g = "barrier helpers (bar.sync / bar.arrive)"
        elif l < L["kern"]: g = "B: team_exscan1"
        elif l < L["teamA"]: g = "stage block (all)"
        elif l < L["rounds"]: g = "A: wait EMPTY + tile setup"
        elif l < L["anyM"]: g = "A: phase 1 rounds (candidates + table)"
        elif l < L["p3"]: g = "A: walk call sites / barriers"
        elif l < L["teamB"]: g = "A: path resolution (P3) + arrive"
        elif l <= L["waitF"] + 1: g = "B: wait FULL"
        elif l < L["p5"]: g = "B: mark call site / barriers"
        elif l < L["flush"]: g = "B: merge + emit (P5)"
        elif l < L["final"]: g = "B: flush"
        else: g = "final literals (all)"
    elif f == "lz4_kernels.cu": g = "c_walk / c_extend / c_emit_seq (both teams)"
    else: g = "intrinsics (shfl, atomics, lds32u)"
    G[g] = G.get(g, 0) + s
for k, v in sorted(G.items(), key=lambda kv: -kv[1]): print(f"{k:48s} {v:9d} {v / tot * 100:5.1f}%")
