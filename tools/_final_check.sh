set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'], {k:round(v['device_gbs'],1) for k,v in d['extra'].items() if isinstance(v,dict) and 'device_gbs' in v}, d['cpu_baseline']['value'])"
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-200
