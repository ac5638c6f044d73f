set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/r2_pytest15.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest15.log
tail -4 gpurun_out/r2_pytest15.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
