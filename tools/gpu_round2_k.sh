set -x
mkdir -p gpurun_out
cp build/libtcs_prof.so vae-diffusion-toy-crystals_b200/toycrystals_b200/libtcs.so
TCS_DEBUG=128 timeout 300 python tools/layer_speed.py 1024 > gpurun_out/r2_kernel_profile_k.txt 2>&1
grep -E "^\[epi|^\[mma" gpurun_out/r2_kernel_profile_k.txt | sort | uniq -c | sort -rn | head -40
