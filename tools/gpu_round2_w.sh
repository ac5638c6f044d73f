set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=200 -p no:cacheprovider -s -k "fused_upsample or fused_blocks" 2>&1 | grep -E "fused upsample|fused vs|passed|failed"
(for i in 1 2; do timeout 300 python tools/layer_speed.py 1024 2>&1 | grep -E "^[0-9]"; done) > gpurun_out/r2_layer_speed_w7.txt 2>&1
cat gpurun_out/r2_layer_speed_w7.txt
