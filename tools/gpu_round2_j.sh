set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=6 --timeout=600 -p no:cacheprovider -s -k "test_conv_layer or golden or tma_store or layers_against" > gpurun_out/r2_pytest10.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest10.log
grep -E "passed|failed|rel-L2|FAILED|Error" gpurun_out/r2_pytest10.log | tail -12
for d in 0 8 16; do TCS_DEBUG=$d timeout 300 python tools/layer_speed.py 1024; done > gpurun_out/r2_layer_speed_j.txt 2>&1
cat gpurun_out/r2_layer_speed_j.txt
timeout 300 python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/r2_bench_j.json 2> gpurun_out/r2_bench_j.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_j.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])"
