set -x
cd $GRAFT_REPO_ROOT
N=${1:-8}
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -c 800 gpurun_out/r2_bench_n$N.err; grep "^{" gpurun_out/r2_bench_n$N.json | cut -c1-400
