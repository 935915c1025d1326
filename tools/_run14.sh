set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pm in 0 2 4 8; do ZSTDMT_B200_D2H_PIECE_MB=$pm python bench.py --mode lz4-decompress --size-gib 8 --steps 2 --warmup 3 --e2e-steps 4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('PIECE_MB=$pm e2e', round(d['e2e']['value'],2), 'ms', round(d['e2e']['ms_per_step'],1))"; done
ZSTDMT_B200_TRACE=1 python bench.py --mode lz4-decompress --size-gib 8 --steps 2 --warmup 3 --e2e-steps 3 2>&1 >/dev/null | grep decompress | tail -2
