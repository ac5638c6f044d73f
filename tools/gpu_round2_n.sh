set -x
mkdir -p gpurun_out
# --- A/B of the TMEM stash (TCS_DEBUG=32 = stash in the staging blocks) + parity of the fused layers
timeout 600 python -m pytest tests -m gpu -q --maxfail=6 --timeout=600 -p no:cacheprovider -s -k "test_conv_layer or golden or tma_store or layers_against or stash" > gpurun_out/r2_pytest12.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest12.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest12.log | tail -5
for d in 0 32 0 32; do TCS_DEBUG=$d timeout 300 python tools/layer_speed.py 1024; done > gpurun_out/r2_layer_speed_n.txt 2>&1
cat gpurun_out/r2_layer_speed_n.txt
export TCS_EXCHANGE_TIMEOUT=400000000000
export TCS_NO_COOP_CLUSTER=1
# (a) one 2048-image pass of the conv family, full metric set
timeout 300 python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_plain_a.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 16 -c 16 -o /tmp/r2b_prof_conv -f python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_conv.log 2>&1
tail -2 gpurun_out/r2_ncu_conv.log
python tools/ncu_conv_summary.py /tmp/r2b_prof_conv.ncu-rep gpurun_out/r2b_conv_ncu_full.json "ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 16 -c 16 python tools/ncu_target.py 1024 1 (one network pass = 2048 images = the production shape; TCS_NO_COOP_CLUSTER=1 because ncu's kernel replay fails launches that carry both the cooperative and the cluster attribute; tap-shift build)" 2048
ncu -i /tmp/r2b_prof_conv.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r2b_conv_ncu_raw.csv.gz
ncu -i /tmp/r2b_prof_conv.ncu-rep --page source --csv --kernel-id :::1 2>/dev/null | gzip > gpurun_out/r2b_conv_ncu_source_down1net3.csv.gz
ls -la /tmp/*.ncu-rep
# (b) the non-GEMM kernels of the same pass
timeout 900 ncu --set full --clock-control none -k regex:"upsample|first_conv|attention|gn_image16" -s 10 -c 5 -o /tmp/r2b_prof_misc -f python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_misc.log 2>&1
python tools/ncu_misc_summary.py /tmp/r2b_prof_misc.ncu-rep gpurun_out/r2b_misc_ncu.json
# (c) the fused update kernel at an HBM-bound size
timeout 120 python tools/ncu_target_step.py > gpurun_out/r2_ncu_plain_c.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 2 -o /tmp/r2b_prof_step -f python tools/ncu_target_step.py > gpurun_out/r2_ncu_step.log 2>&1
python tools/ncu_misc_summary.py /tmp/r2b_prof_step.ncu-rep gpurun_out/r2b_step_ncu.json
ncu -i /tmp/r2b_prof_step.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r2b_step_ncu_raw.csv.gz
unset TCS_EXCHANGE_TIMEOUT
# (d) launch list of bench.py itself
timeout 600 python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/r2_bench_m.json 2> gpurun_out/r2_bench_m.err && TCS_EXCHANGE_TIMEOUT=400000000000 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/r2_ncu_bench.log 2>&1
tail -2 gpurun_out/r2_ncu_bench.log
du -sh gpurun_out; ls -la gpurun_out/ | tail -30
