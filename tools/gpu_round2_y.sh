set -x
mkdir -p gpurun_out
export TCS_EXCHANGE_TIMEOUT=400000000000
export TCS_NO_COOP_CLUSTER=1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|attn_block" -s 15 -c 15 -o /tmp/r2e_prof -f python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_conv3.log 2>&1
tail -2 gpurun_out/r2_ncu_conv3.log
ncu -i /tmp/r2e_prof.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r2e_pass_ncu_raw.csv.gz
ls -la gpurun_out/r2e_pass_ncu_raw.csv.gz
