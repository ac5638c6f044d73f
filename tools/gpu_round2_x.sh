set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err; echo "bench exit $?"
tail -c 2500 gpurun_out/r2_bench_q.json
TCS_EXCHANGE_TIMEOUT=400000000000 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 520 --csv --log-file gpurun_out/r2d_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/r2_ncu_bench3.log 2>&1
python tools/launch_summary.py gpurun_out/r2d_launches_bench.csv
