set -x
mkdir -p gpurun_out
for sg in 0 2800 1400; do
echo "== stagger $sg"
TCS_ATTN_STAGGER=$sg timeout 300 python tools/attn_profile.py 2048 2>&1 | grep -E "phase0|head 0 conv|Y ready|rows staged|written|image done"
TCS_ATTN_STAGGER=$sg timeout 300 python tools/layer_speed.py 1024 2>&1 | tail -1 | cut -c1-40,600-
done > gpurun_out/r2_attn_stagger.txt 2>&1
cat gpurun_out/r2_attn_stagger.txt
