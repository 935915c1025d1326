set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/zstd_decode_probe.py 1 text 2>&1 | tail -2
python tools/zstd_decode_probe.py 4 text 2>&1 | tail -2
python tools/zstd_decode_probe.py 4 mix 2>&1 | tail -2
timeout 200 python tools/stress_gpu.py 60 17 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'zstd_|lz77_' -c 400 --csv --log-file gpurun_out/r2_launches_zstd_probe.csv python tools/zstd_decode_probe.py 0.25 text > gpurun_out/ncu_l3.log 2>&1
