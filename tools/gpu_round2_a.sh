set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --maxfail=25 -x --timeout=600 -p no:cacheprovider > gpurun_out/r2_pytest1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench exit $?"
tail -c 600 gpurun_out/r2_bench1.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_ref1.json 2> gpurun_out/r2_ref1.err; echo "ref exit $?"
