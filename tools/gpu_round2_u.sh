set -x
mkdir -p gpurun_out
export TCS_EXCHANGE_TIMEOUT=400000000000
export TCS_NO_COOP_CLUSTER=1
# (a) the fused attention block of one 2048-image pass, full metric set
timeout 300 python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_plain_attn.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_block" -s 1 -c 2 -o /tmp/r2c_prof_attn -f python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_attn.log 2>&1
tail -2 gpurun_out/r2_ncu_attn.log
python tools/ncu_misc_summary.py /tmp/r2c_prof_attn.ncu-rep gpurun_out/r2c_attn_ncu.json
ncu -i /tmp/r2c_prof_attn.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r2c_attn_ncu_raw.csv.gz
unset TCS_EXCHANGE_TIMEOUT
# (b) launch list of bench.py itself
timeout 600 python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/r2_bench_p.json 2> gpurun_out/r2_bench_p.err && TCS_EXCHANGE_TIMEOUT=400000000000 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/r2_ncu_bench2.log 2>&1
tail -2 gpurun_out/r2_ncu_bench2.log
python tools/launch_summary.py gpurun_out/r2c_launches_bench.csv
