set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_lz4.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-extra --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('SIDE value', round(d['value'],2), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()})"
ZSTDMT_B200_NO_SIDE_STREAM=1 python bench.py --steps 10 --warmup 3 --no-extra --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NOSIDE value', round(d['value'],2), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms'].items()})"
