set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 -p no:cacheprovider -s -k "fused_attention or layers_against or golden or graph_replay" > gpurun_out/r2_attn_t2.log 2>&1; echo "exit $?" >> gpurun_out/r2_attn_t2.log
grep -E "fused attention|passed|failed|FAILED|Error|worst" gpurun_out/r2_attn_t2.log | tail -20
for f in 1 0; do TCS_FUSE_ATTN=$f timeout 300 python tools/layer_speed.py 1024; done > gpurun_out/r2_layer_speed_o.txt 2>&1
cat gpurun_out/r2_layer_speed_o.txt
