set -x
cd $GRAFT_REPO_ROOT
python tools/dbg_zstdmt_style.py 2>&1 | tail -12
ZSTDMT_B200_NO_FAST_ENTROPY=1 python tools/dbg_zstdmt_style.py 2>&1 | tail -12
python -m pytest tests -m gpu -q 2>&1 | tail -6
