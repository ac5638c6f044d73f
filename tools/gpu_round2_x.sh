set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --skip-e2e > gpurun_out/r2_bench_check.json 2> gpurun_out/r2_bench_check.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_check.json').read().strip().splitlines()[-1])
r=d['roofline']; print(d['value'], r['achieved'], r['frac'], r['frac_of_burst'], r['launch_ms'], r['conv_layers'], r['attention_block_ms'], r['share_of_pass'], r['flop_per_launch_set'], r['traffic'])
P
