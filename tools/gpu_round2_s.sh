set -x
mkdir -p gpurun_out
for f in 0 1 0 1; do TCS_FUSE_UPS=$f timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fuse_ups=$f', round(d['value'],2), d['clocks']['sm_mhz'], round(d['roofline']['pass_ms'],3))"; done > gpurun_out/r2_bench_ab_ups.txt 2>&1
grep "^fuse" gpurun_out/r2_bench_ab_ups.txt
