set -x
mkdir -p gpurun_out
timeout 700 python tools/sweep.py > gpurun_out/r2_sweep_batch.json 2> gpurun_out/r2_sweep_batch.err; echo "sweep exit $?"
tail -30 gpurun_out/r2_sweep_batch.err | cut -c1-220
