set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --timeout=900 -p no:cacheprovider -s -k "graph or ode_update or bit_identical or sde_update or philox or full_size_pass or stash or neighbour or prior and full_size or seed_alone" > gpurun_out/r2_pytest3.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest3.log
grep -E "passed|failed|rel-L2|launch mode|FAILED" gpurun_out/r2_pytest3.log | tail -30
timeout 300 python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_c.json').read().strip().splitlines()[-1]); print(d['value'], d['step_kernel'], d['roofline']['frac'])"
export TCS_EXCHANGE_TIMEOUT=400000000000
timeout 300 python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_plain.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 16 -c 16 -o gpurun_out/r2_prof_conv -f python tools/ncu_target.py 1024 1 > gpurun_out/r2_ncu_conv.log 2>&1
tail -3 gpurun_out/r2_ncu_conv.log
timeout 300 python tools/ncu_target.py 1024 1 0 fp32 tcgen05 > gpurun_out/r2_ncu_plain3.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 70 --csv --log-file gpurun_out/r2_launches_fp32tc.csv python tools/ncu_target.py 1024 1 0 fp32 tcgen05 > gpurun_out/r2_ncu_fp32tc.log 2>&1
timeout 120 python tools/ncu_target_step.py > gpurun_out/r2_ncu_plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 2 -o gpurun_out/r2_prof_step -f python tools/ncu_target_step.py > gpurun_out/r2_ncu_step.log 2>&1
ls -la gpurun_out/*.ncu-rep
