set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=200 -p no:cacheprovider -s -k "fused_blocks_agree" > gpurun_out/r2_ab_t.log 2>&1; echo "exit $?" >> gpurun_out/r2_ab_t.log
grep -E "fused vs|passed|failed|Error|assert" gpurun_out/r2_ab_t.log | tail
