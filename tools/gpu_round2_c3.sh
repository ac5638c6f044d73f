# BASELINE configs[2]: n = 65536 reverse-SDE samples sharded over N GPUs of one box, all-gather of the images inside the timed region
set -x
N=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > gpurun_out/r2_c3_smi_n$N.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --scaling strong --n-total 65536 --steps 1 --warmup 1 --warmup-sde-steps 10 --no-cpu-baseline "$@" > gpurun_out/r2_c3_n$N.json 2> gpurun_out/r2_c3_n$N.err; echo "exit $?"
tail -c 1200 gpurun_out/r2_c3_n$N.json
grep -E "nranks|NVLS|comm 0x|\[bench\]" gpurun_out/r2_c3_n$N.err | head -12 > gpurun_out/r2_c3_nccl_n$N.txt; head -5 gpurun_out/r2_c3_nccl_n$N.txt
