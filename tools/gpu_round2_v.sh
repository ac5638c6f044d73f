set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=200 -p no:cacheprovider -s -k "fused_upsample" > gpurun_out/r2_ups_t1.log 2>&1; echo "exit $?" >> gpurun_out/r2_ups_t1.log
grep -E "fused upsample|passed|failed|FAILED|Error|error|assert" gpurun_out/r2_ups_t1.log | tail -20
