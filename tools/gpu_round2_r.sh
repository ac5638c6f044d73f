set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/r2_pytest13.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest13.log
tail -5 gpurun_out/r2_pytest13.log
timeout 300 python tools/attn_profile.py 2048 > gpurun_out/r2_attn_profile.txt 2>&1
grep -E "written|image done" gpurun_out/r2_attn_profile.txt
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_o.json 2> gpurun_out/r2_bench_o.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_o.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'], d['clocks'], d['roofline']['achieved'], d['roofline']['per_layer_tflops'], d['roofline']['pass_ms'])
P
