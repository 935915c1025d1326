set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for st in 0 1; do for g in 4 32; do ZSTDMT_B200_LZ4_STAGED=$st python bench.py --mode lz4-decompress --size-gib $g --steps 5 --warmup 3 --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('STAGED=$st ${g}GiB', round(d['value'],1), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms'].items()})"; done; done
