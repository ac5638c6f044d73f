set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout=600 -p no:cacheprovider -k "test_conv_layer or golden or tma_store or output_conv or layers_against" > gpurun_out/r2_pytest6.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest6.log
tail -4 gpurun_out/r2_pytest6.log
for d in 0 8 16 2; do TCS_DEBUG=$d timeout 300 python tools/layer_speed.py 1024; done > gpurun_out/r2_layer_speed_kb64.txt 2>&1
cat gpurun_out/r2_layer_speed_kb64.txt
timeout 300 python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_f.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])"
