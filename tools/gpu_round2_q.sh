set -x
mkdir -p gpurun_out
for f in 0 1 0 1; do TCS_FUSE_ATTN=$f timeout 300 python tools/layer_speed.py 1024; done > gpurun_out/r2_layer_speed_q.txt 2>&1
cat gpurun_out/r2_layer_speed_q.txt
