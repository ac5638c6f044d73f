set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 --timeout=900 -p no:cacheprovider -k "not test_conv_layer" > gpurun_out/r2_pytest2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest2.log
tail -15 gpurun_out/r2_pytest2.log
# what bounds the conv family at the production pass size (timing experiments: results are garbage by design)
for d in 0 8 16 24 2 1; do TCS_DEBUG=$d timeout 300 python tools/layer_speed.py 1024; done > gpurun_out/r2_layer_speed_debug.txt 2>&1
cat gpurun_out/r2_layer_speed_debug.txt
# ncu: full capture of one 2048-image pass of the conv family + the step kernel at an HBM-bound size
timeout 300 python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 16 -c 16 -o gpurun_out/r2_prof_conv -f python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_conv.log 2>&1
timeout 120 python tools/ncu_target_step.py > gpurun_out/r2_ncu_plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 2 -o gpurun_out/r2_prof_step -f python tools/ncu_target_step.py > gpurun_out/r2_ncu_step.log 2>&1
ls -la gpurun_out/*.ncu-rep
