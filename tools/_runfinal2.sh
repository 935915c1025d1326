set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -1
ZSTDMT_B200_D2H_PIECE_MB=1 python bench.py --mode lz4-decompress --size-gib 8 --steps 2 --warmup 3 --e2e-steps 4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('PIECE_MB=1 e2e', round(d['e2e']['value'],2))"
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_reference_arm.json 2> gpurun_out/r2_reference_arm.err
ZSTDMT_B200_TRACE=1 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; grep "compress:" gpurun_out/r2_bench_n1.err | tail -8
python bench.py --impl reference --mode lz4-decompress --steps 5 --warmup 2 > gpurun_out/r2_reference_arm_config3.json 2>> gpurun_out/r2_reference_arm.err
python bench.py --mode lz4-decompress --steps 20 --warmup 3 > gpurun_out/r2_bench_config3_n1.json 2> gpurun_out/r2_bench_config3_n1.err; tail -c 300 gpurun_out/r2_bench_config3_n1.err
python bench.py --mode zstd-compress --steps 10 --warmup 3 --no-extra > gpurun_out/r2_bench_config4_n1.json 2> gpurun_out/r2_bench_config4_n1.err
python bench.py --mode zstd-mix --steps 10 --warmup 3 --no-extra > gpurun_out/r2_bench_config5_n1.json 2> gpurun_out/r2_bench_config5_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lz4_|xxh32|scan_u64|lz77_|zstd_' -c 200 --csv --log-file gpurun_out/r2_launches_config2.csv python bench.py --size-gib 2 --steps 2 --warmup 3 --no-e2e --no-extra > gpurun_out/ncu_l2.log 2>&1
