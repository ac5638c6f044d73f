set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=6 --timeout=600 -p no:cacheprovider -s -k "test_conv_layer or golden or tma_store or output_conv or layers_against or full_size_pass" > gpurun_out/r2_pytest9.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest9.log
grep -E "passed|failed|rel-L2|FAILED|Error" gpurun_out/r2_pytest9.log | tail -20
for d in 0 8 16; do TCS_DEBUG=$d timeout 300 python tools/layer_speed.py 1024; done > gpurun_out/r2_layer_speed_i.txt 2>&1
TCS_TB=1 timeout 300 python tools/layer_speed.py 1024 >> gpurun_out/r2_layer_speed_i.txt 2>&1
cat gpurun_out/r2_layer_speed_i.txt
