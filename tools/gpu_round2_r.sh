set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/r2_pytest14.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest14.log
tail -5 gpurun_out/r2_pytest14.log
