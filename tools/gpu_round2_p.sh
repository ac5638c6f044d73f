set -x
mkdir -p gpurun_out
timeout 300 python tools/attn_profile.py 2048 > gpurun_out/r2_attn_profile.txt 2>&1
cat gpurun_out/r2_attn_profile.txt
