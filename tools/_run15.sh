set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_lz4.py tests/test_gpu_zstd.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -3
for cfg in "" "ZSTDMT_B200_SLOTS=8" "ZSTDMT_B200_SLOTS=8 ZSTDMT_B200_BATCH_MB=8" "ZSTDMT_B200_SLOTS=16"; do env $cfg python bench.py --size-gib 4 --steps 2 --warmup 3 --e2e-steps 4 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$cfg] e2e', round(d['e2e']['value'],2))"; done
ncu --set full --clock-control none --import-source on -k regex:'lz77_blocks_kernel' -s 2 -c 1 -o gpurun_out/prof_zstdc_r2 python bench.py --mode zstd-compress --size-gib 1 --steps 1 --warmup 3 --no-e2e --no-extra > gpurun_out/ncu_zc.log 2>&1
ls -la gpurun_out/prof_zstdc_r2.ncu-rep
