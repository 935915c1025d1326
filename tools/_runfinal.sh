set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_reference_arm.json 2> gpurun_out/r2_reference_arm.err
ZSTDMT_B200_TRACE=1 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 1500 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --mode lz4-decompress --steps 5 --warmup 2 > gpurun_out/r2_reference_arm_config3.json 2>> gpurun_out/r2_reference_arm.err
python bench.py --mode lz4-decompress --steps 20 --warmup 3 > gpurun_out/r2_bench_config3_n1.json 2> gpurun_out/r2_bench_config3_n1.err; tail -c 500 gpurun_out/r2_bench_config3_n1.err
python bench.py --impl reference --mode zstd-compress --steps 3 --warmup 1 > gpurun_out/r2_reference_arm_config4.json 2>> gpurun_out/r2_reference_arm.err
python bench.py --mode zstd-compress --steps 10 --warmup 3 > gpurun_out/r2_bench_config4_n1.json 2> gpurun_out/r2_bench_config4_n1.err; tail -c 500 gpurun_out/r2_bench_config4_n1.err
python bench.py --impl reference --mode zstd-mix --steps 3 --warmup 1 > gpurun_out/r2_reference_arm_config5.json 2>> gpurun_out/r2_reference_arm.err
python bench.py --mode zstd-mix --steps 10 --warmup 3 > gpurun_out/r2_bench_config5_n1.json 2> gpurun_out/r2_bench_config5_n1.err; tail -c 500 gpurun_out/r2_bench_config5_n1.err
# ncu evidence of the final kernels: launch lists + full captures
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lz4_|xxh32|scan_u64|lz77_|zstd_' -c 200 --csv --log-file gpurun_out/r2_launches_config3.csv python bench.py --mode lz4-decompress --size-gib 4 --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lz4_|xxh32|scan_u64|lz77_|zstd_' -c 200 --csv --log-file gpurun_out/r2_launches_config2.csv python bench.py --size-gib 2 --steps 2 --warmup 3 --no-e2e --no-extra > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'lz4_parse_blocks|lz4_exec_blocks' -s 2 -c 2 -o gpurun_out/prof_lz4d_r2final python bench.py --mode lz4-decompress --size-gib 32 --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'zstd_seq_predef|zstd_literals|zstd_execute' -s 3 -c 3 -o gpurun_out/prof_zstdd_r2final python tools/zstd_decode_probe.py 0.5 text > gpurun_out/ncu_z1.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
