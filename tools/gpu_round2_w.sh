set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=200 -p no:cacheprovider -s -k "fused_upsample" 2>&1 | grep -E "fused upsample|passed|failed"
(for d in 0 64; do TCS_DEBUG=$d TCS_FUSE_UPS=1 timeout 300 python tools/layer_speed.py 1024 2>&1 | grep -E "^[0-9]"; done) > gpurun_out/r2_layer_speed_w5.txt 2>&1
cat gpurun_out/r2_layer_speed_w5.txt
