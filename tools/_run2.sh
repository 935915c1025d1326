set -x
cd $GRAFT_REPO_ROOT
python __graft_entry__.py --smoke 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ZSTDMT_B200_TRACE=1 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 3000 gpurun_out/r2_bench_a.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_ref_a.json 2>> gpurun_out/r2_bench_a.err
ZSTDMT_B200_TRACE=1 ZSTDMT_B200_NUMA=0 python bench.py --steps 3 --warmup 3 --no-extra --no-bind > gpurun_out/r2_bench_nonuma.json 2> gpurun_out/r2_bench_nonuma.err
ZSTDMT_B200_TRACE=1 python bench.py --mode lz4-decompress --steps 5 --warmup 3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; tail -c 1500 gpurun_out/r2_bench_c3.err
python bench.py --impl reference --mode lz4-decompress --steps 3 --warmup 1 > gpurun_out/r2_ref_c3.json 2>> gpurun_out/r2_bench_c3.err
nproc; numactl -H 2>/dev/null | head -5; lscpu | grep -i "numa\|model name" 
