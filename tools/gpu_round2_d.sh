set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout=600 -p no:cacheprovider -k "test_conv_layer or golden or tma_store or output_conv" > gpurun_out/r2_pytest4.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest4.log
tail -4 gpurun_out/r2_pytest4.log
for d in 0 16; do TCS_DEBUG=$d timeout 300 python tools/layer_speed.py 1024; done > gpurun_out/r2_layer_speed_wide_b.txt 2>&1
cat gpurun_out/r2_layer_speed_wide_b.txt
export TCS_EXCHANGE_TIMEOUT=400000000000
export TCS_NO_COOP_CLUSTER=1
timeout 300 python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_plain.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 16 -c 16 -o gpurun_out/r2_prof_conv -f python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_conv.log 2>&1
tail -3 gpurun_out/r2_ncu_conv.log
ls -la gpurun_out/*.ncu-rep
