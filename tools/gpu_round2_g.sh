set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
( time timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 --timeout=900 -p no:cacheprovider --durations=20 ) > gpurun_out/r2_pytest7.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest7.log
tail -40 gpurun_out/r2_pytest7.log
TCS_DEBUG=0 timeout 300 python tools/layer_speed.py 1024 > gpurun_out/r2_layer_speed_g.txt 2>&1
cat gpurun_out/r2_layer_speed_g.txt
timeout 300 python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; tail -c 1500 gpurun_out/r2_bench_g.json
