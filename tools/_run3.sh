set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/zstd_decode_probe.py 1 text 2>&1 | tail -3
ZSTDMT_B200_NO_FAST_ENTROPY=1 python tools/zstd_decode_probe.py 1 text 2>&1 | tail -3
python tools/zstd_decode_probe.py 1 mix 2>&1 | tail -3
python bench.py --mode lz4-decompress --steps 5 --warmup 3 --no-e2e > gpurun_out/r2b_c3.json 2> gpurun_out/r2b_c3.err; tail -c 600 gpurun_out/r2b_c3.err; cut -c1-600 gpurun_out/r2b_c3.json
# launch lists (cold, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lz4_|xxh32|scan_u64|lz77_|zstd_' -c 200 --csv --log-file gpurun_out/r2_launches_c3.csv python bench.py --mode lz4-decompress --size-gib 4 --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lz4_|xxh32|scan_u64|lz77_|zstd_' -c 200 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --size-gib 2 --steps 2 --warmup 3 --no-e2e --no-extra > gpurun_out/ncu_l2.log 2>&1
# full captures of the two decode passes (4 GiB reference-framed)
ncu --set full --clock-control none --import-source on -k regex:'lz4_parse_blocks|lz4_exec_blocks' -s 2 -c 2 -o gpurun_out/prof_lz4d_r2a python bench.py --mode lz4-decompress --size-gib 4 --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_f1.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
