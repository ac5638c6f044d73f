set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['clocks'], d['gpu_launches'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['pass_ms'], d['whole_job']['frac'])
print(d['roofline']['per_layer_tflops'])
print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in d['extra'].items()})
P
