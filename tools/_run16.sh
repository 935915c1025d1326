set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_zstd.py tests/test_gpu_plain_streams.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -3
python bench.py --mode zstd-compress --size-gib 4 --steps 5 --warmup 3 --no-e2e --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('zstd text', round(d['value'],2), round(d['ms_per_step'],2), 'ratio', round(d['ratio'],4))"
python bench.py --mode zstd-mix --size-gib 4 --steps 5 --warmup 3 --no-e2e --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('zstd mix', round(d['value'],2), round(d['ms_per_step'],2), 'ratio', round(d['ratio'],4))"
timeout 150 python tools/stress_gpu.py 45 23 2>&1 | tail -1
