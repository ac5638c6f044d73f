set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 --timeout=900 -p no:cacheprovider > gpurun_out/r2_pytest5.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest5.log
tail -6 gpurun_out/r2_pytest5.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'], d['extra']['fp32_tensor_core_mode'].get('value'), d['extra']['prior'].get('value'))"
timeout 300 python tools/ncu_target.py 1024 1 0 fp32 tcgen05 > gpurun_out/r2_ncu_plain3.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 95 -c 110 --csv --log-file gpurun_out/r2_launches_fp32tc.csv python tools/ncu_target.py 1024 1 0 fp32 tcgen05 > gpurun_out/r2_ncu_fp32tc.log 2>&1
