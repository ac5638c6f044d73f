set -x
mkdir -p gpurun_out
export TCS_EXCHANGE_TIMEOUT=400000000000
export TCS_NO_COOP_CLUSTER=1
# (a) one 2048-image pass of the conv family, full metric set
timeout 300 python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_plain_a.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 16 -c 16 -o gpurun_out/r2b_prof_conv -f python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_conv.log 2>&1
tail -2 gpurun_out/r2_ncu_conv.log
# (b) the non-GEMM kernels of the same pass
timeout 900 ncu --set full --clock-control none -k regex:"upsample|first_conv|attention|gn_image16" -s 10 -c 5 -o gpurun_out/r2b_prof_misc -f python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_misc.log 2>&1
tail -2 gpurun_out/r2_ncu_misc.log
# (c) the fused update kernel at an HBM-bound size
timeout 120 python tools/ncu_target_step.py > gpurun_out/r2_ncu_plain_c.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 2 -o gpurun_out/r2b_prof_step -f python tools/ncu_target_step.py > gpurun_out/r2_ncu_step.log 2>&1
tail -2 gpurun_out/r2_ncu_step.log
unset TCS_EXCHANGE_TIMEOUT
# (d) launch list of bench.py itself
timeout 600 python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/r2_bench_m.json 2> gpurun_out/r2_bench_m.err && TCS_EXCHANGE_TIMEOUT=400000000000 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/r2_ncu_bench.log 2>&1
tail -2 gpurun_out/r2_ncu_bench.log
ls -la gpurun_out/
