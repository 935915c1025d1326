set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for g in 1 4 32; do python bench.py --mode lz4-decompress --size-gib $g --steps 5 --warmup 3 --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('${g}GiB', round(d['value'],1), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms'].items()})"; done
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_decode.py 3 2>&1 | tail -12
timeout 500 compute-sanitizer --tool racecheck python tools/sanitize_decode.py 2 2>&1 | tail -8
