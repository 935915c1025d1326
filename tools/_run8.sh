set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/zstd_decode_probe.py 1 text 2>&1 | tail -2
python tools/zstd_decode_probe.py 1 mix 2>&1 | tail -2
python bench.py --mode lz4-decompress --steps 5 --warmup 3 --no-e2e > gpurun_out/r2e_c3.json 2> gpurun_out/r2e_c3.err; tail -c 300 gpurun_out/r2e_c3.err; python -c "
import json; d=json.load(open('gpurun_out/r2e_c3.json')); print('C3', d['value'], d['ms_per_step'], d['kernel_ms'])"
python bench.py --mode lz4-decompress --size-gib 4 --steps 5 --warmup 3 --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C3 4GiB', d['value'], d['ms_per_step'], d['kernel_ms'])"
timeout 200 python tools/stress_gpu.py 90 11 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'zstd_|lz77_' -c 400 --csv --log-file gpurun_out/r2_launches_zstd_probe.csv python tools/zstd_decode_probe.py 0.25 text > gpurun_out/ncu_l3.log 2>&1
