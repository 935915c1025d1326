set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/zstd_decode_probe.py 1 text 2>&1 | tail -3
ZSTDMT_B200_NO_FAST_ENTROPY=1 python tools/zstd_decode_probe.py 1 text 2>&1 | tail -3
python tools/zstd_decode_probe.py 1 mix 2>&1 | tail -3
timeout 300 python tools/stress_gpu.py 120 7 2>&1 | tail -4
