set -x
mkdir -p gpurun_out
timeout 120 ./build/umma_shift_test > gpurun_out/r2_umma_shift.txt 2>&1; echo "probe exit $?" >> gpurun_out/r2_umma_shift.txt
cat gpurun_out/r2_umma_shift.txt
timeout 300 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider -k "tma_store" > gpurun_out/r2_pytest8.log 2>&1; tail -3 gpurun_out/r2_pytest8.log
for d in 0 8 16 24; do TCS_DEBUG=$d timeout 300 python tools/layer_speed.py 1024; done > gpurun_out/r2_layer_speed_h.txt 2>&1
TCS_NO_COOP_CLUSTER=1 timeout 300 python tools/layer_speed.py 1024 >> gpurun_out/r2_layer_speed_h.txt 2>&1
cat gpurun_out/r2_layer_speed_h.txt
