set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/zstd_decode_probe.py 1 text 2>&1 | tail -2
python tools/zstd_decode_probe.py 1 mix 2>&1 | tail -2
python bench.py --mode lz4-decompress --steps 5 --warmup 3 --no-e2e > gpurun_out/r2c_c3.json 2> gpurun_out/r2c_c3.err; tail -c 400 gpurun_out/r2c_c3.err; python -c "
import json; d=json.load(open('gpurun_out/r2c_c3.json')); print('C3', d['value'], d['ms_per_step'], d['kernel_ms'])"
# e2e batch-size experiments (decompress of reference-framed + own stream; compress)
for mb in 32 8 4 2; do echo "DBATCH_MB=$mb"; ZSTDMT_B200_DBATCH_MB=$mb python bench.py --mode lz4-decompress --size-gib 8 --steps 2 --warmup 3 --e2e-steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('  e2e', d['e2e']['value'], 'device', d['value'])"; done
for mb in 8 4 16; do echo "BATCH_MB=$mb"; ZSTDMT_B200_BATCH_MB=$mb python bench.py --size-gib 4 --steps 2 --warmup 3 --e2e-steps 3 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('  e2e', d['e2e']['value'], 'device', d['value'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'zstd_|lz77_' -c 400 --csv --log-file gpurun_out/r2_launches_zstd_probe.csv python tools/zstd_decode_probe.py 0.25 text > gpurun_out/ncu_l3.log 2>&1
# ncu: zstd decode kernels (own stream 512 MiB text) and the updated lz4 decode passes
ncu --set full --clock-control none --import-source on -k regex:'zstd_seq_predef|zstd_literals|zstd_execute' -s 3 -c 3 -o gpurun_out/prof_zstdd_r2a python tools/zstd_decode_probe.py 0.5 text > gpurun_out/ncu_z1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'lz4_parse_blocks|lz4_exec_blocks' -s 2 -c 2 -o gpurun_out/prof_lz4d_r2b python bench.py --mode lz4-decompress --size-gib 32 --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
