set -x
cd $GRAFT_REPO_ROOT
N=${1:-2}
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -c 1200 gpurun_out/r2_bench_n$N.err; cut -c1-700 gpurun_out/r2_bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 --mode lz4-decompress > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err; tail -c 600 gpurun_out/r2_bench_c3_n$N.err; cut -c1-500 gpurun_out/r2_bench_c3_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 5 --warmup 3 --mode zstd-mix --no-extra > gpurun_out/r2_bench_c5_n$N.json 2> gpurun_out/r2_bench_c5_n$N.err; tail -c 600 gpurun_out/r2_bench_c5_n$N.err; cut -c1-500 gpurun_out/r2_bench_c5_n$N.json
