set -x
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 --timeout=900 -p no:cacheprovider ) > gpurun_out/r2_pytest11.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest11.log
tail -8 gpurun_out/r2_pytest11.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_l.json 2> gpurun_out/r2_bench_l.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_l.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'], d['extra']['fp32_tensor_core_mode'].get('value'), d['extra']['prior'].get('value'), d['cpu_baseline'])"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_ref_l.json 2> gpurun_out/r2_bench_ref_l.err; echo "ref exit $?"; tail -c 1500 gpurun_out/r2_bench_ref_l.json
