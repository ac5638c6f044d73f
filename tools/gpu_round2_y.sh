set -x
mkdir -p gpurun_out
export TCS_EXCHANGE_TIMEOUT=400000000000
export TCS_NO_COOP_CLUSTER=1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 520 --csv --log-file gpurun_out/r2d_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/r2_ncu_bench3.log 2>&1
tail -2 gpurun_out/r2_ncu_bench3.log
python tools/launch_summary.py gpurun_out/r2d_launches_bench.csv
# one 2048-image pass of the conv family, full metric set (12 launches per pass now: first conv, 13 convs incl. UPS, attention, eps)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 14 -c 14 -o /tmp/r2d_prof_conv -f python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_conv2.log 2>&1
tail -2 gpurun_out/r2_ncu_conv2.log
python tools/ncu_conv_summary.py /tmp/r2d_prof_conv.ncu-rep gpurun_out/r2d_conv_ncu_full.json "ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 14 -c 14 python tools/ncu_target.py 1024 1 (one network pass = 2048 images; TCS_NO_COOP_CLUSTER=1 because ncu's kernel replay fails launches that carry both the cooperative and the cluster attribute; fused upsample + fused attention build)" 2048
ncu -i /tmp/r2d_prof_conv.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r2d_conv_ncu_raw.csv.gz
