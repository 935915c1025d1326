set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/zstd_decode_probe.py 1 text 2>&1 | tail -2
python tools/zstd_decode_probe.py 1 mix 2>&1 | tail -2
python tools/zstd_decode_probe.py 4 text 2>&1 | tail -2
python bench.py --mode lz4-decompress --steps 5 --warmup 3 --no-e2e > gpurun_out/r2f_c3.json 2> gpurun_out/r2f_c3.err; tail -c 300 gpurun_out/r2f_c3.err; python -c "
import json; d=json.load(open('gpurun_out/r2f_c3.json')); print('C3', d['value'], d['ms_per_step'], d['kernel_ms'])"
for st in 0 1; do for g in 1 4 8; do ZSTDMT_B200_LZ4_STAGED=$st python bench.py --mode lz4-decompress --size-gib $g --steps 5 --warmup 3 --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('STAGED=$st ${g}GiB', round(d['value'],1), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms'].items()})"; done; done
timeout 200 python tools/stress_gpu.py 60 13 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'zstd_|lz77_' -c 400 --csv --log-file gpurun_out/r2_launches_zstd_probe.csv python tools/zstd_decode_probe.py 0.25 text > gpurun_out/ncu_l3.log 2>&1
