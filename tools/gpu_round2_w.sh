set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=200 -p no:cacheprovider -s -k "fused_upsample or layers_against" 2>&1 | grep -E "fused upsample|passed|failed"
(for a in 2 3 2 3; do TCS_UPS_A_STAGES=$a TCS_PLAN_LOG=1 timeout 300 python tools/layer_speed.py 1024 2>&1 | grep -E "ups=1|^[0-9]" | sort | uniq; done) > gpurun_out/r2_layer_speed_w6.txt 2>&1
cat gpurun_out/r2_layer_speed_w6.txt
