set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 -p no:cacheprovider -s -k "fused_attention" > gpurun_out/r2_attn_t3.log 2>&1; echo "exit $?" >> gpurun_out/r2_attn_t3.log
grep -E "fused attention|passed|failed|FAILED|Error|worst" gpurun_out/r2_attn_t3.log | tail -20
timeout 300 python tools/attn_profile.py 2048 > gpurun_out/r2_attn_profile.txt 2>&1
cat gpurun_out/r2_attn_profile.txt
timeout 300 python tools/layer_speed.py 1024 > gpurun_out/r2_layer_speed_p.txt 2>&1
cat gpurun_out/r2_layer_speed_p.txt
