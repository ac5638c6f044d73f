set -x
mkdir -p gpurun_out
for f in 1 0 1 0; do TCS_GN_GROUP_GRID=$f timeout 300 python tools/layer_speed.py 1024 2>&1 | tail -1; done > gpurun_out/r2_layer_speed_z.txt 2>&1
cat gpurun_out/r2_layer_speed_z.txt
TCS_GN_GROUP_GRID=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 -p no:cacheprovider -k "tma_store or golden or layers_against or determinism or busy or n1024 or full_size" 2>&1 | tail -3
