"""Aggregate an ncu SASS-level source page by CUDA source line.
usage: ncu_lines.py <report.ncu-rep> <kernel-substring> [lib.so]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, kname = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "zstdmt_b200", "libzstdmt_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
line_of = {}   # instruction offset -> (file, line)
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    asm = subprocess.run(["nvdisasm", "-g", "-c", f], cwd=tmp, capture_output=True, text=True).stdout
    cur = None; infn = False
    for ln in asm.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m: infn = kname in m.group(1); continue
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and cur: line_of[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kname], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
base = None; agg = collections.defaultdict(lambda: [0, 0, 0]); tot = 0
for r in rows[hi + 1:]:
    if len(r) <= iT or r[0] == "Address" or not r[0].startswith("0x"): continue
    a = int(r[0], 16)
    if base is None: base = a
    k = line_of.get(a - base, ("?", 0))
    s = int(r[iS] or 0); agg[k][0] += s; agg[k][1] += int(r[iI] or 0); agg[k][2] += int(r[iT] or 0); tot += s
print("total samples", tot)
srcs = {}
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "40"))]:
    if f not in srcs:
        p = os.path.join(os.path.dirname(lib), "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        if not srcs[f]:
            p2 = os.path.join(os.path.dirname(os.path.dirname(lib)), "include", f)
            srcs[f] = open(p2).read().splitlines() if os.path.exists(p2) else []
    text = srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
    print(f"{v[0]:8d} {v[0]/max(tot,1)*100:5.1f}%  inst={v[1]:11d} thr/inst={v[2]/max(v[1],1):5.1f}  {f}:{l:<4d} {text}")

if os.environ.get("BYINST"):
    print("---- by instructions executed")
    ti = sum(v[1] for v in agg.values())
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("TOP", "40"))]:
        text = srcs.get(f, [])
        if f not in srcs:
            p = os.path.join(os.path.dirname(lib), "csrc", f); srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        t = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
        print(f"{v[1]:11d} {v[1]/ti*100:5.1f}%  samples={v[0]:7d} thr/inst={v[2]/max(v[1],1):5.1f}  {f}:{l:<4d} {t}")
    print("total inst", ti)
