#!/usr/bin/env bash
# Round-2 profile set of the VQ codebook search (run through gpurun, one GPU):
#   gpurun --timeout 900 -- 'bash scripts/profile_r02_vq.sh r02'
# full ncu capture of vq_argmin_tc2_kernel, its per-tile timeline, the A/B against the round-1 kernel in both
# layouts and the launch list of the PR-DVQVAE2 step with the new kernel.
tag=${1:-r02}
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_argmin_tc2 -s 3 -c 1 \
    -o gpurun_out/${tag}_vq_tc2 -f python tools/vq_bench.py > /dev/null 2>&1
{
  echo "# tools/vq_bench.py: 2^20 positions (4096 frames), 10 launches, CUDA events; v2 = vq_argmin_tc2_kernel, v1 = LVT_VQ_TC1=1"
  echo -n "v2 NCHW          "; timeout 60 python tools/vq_bench.py 2>&1 | tail -1
  echo -n "v1 NCHW          "; LVT_VQ_TC1=1 timeout 60 python tools/vq_bench.py 2>&1 | tail -1
  echo -n "v2 NHWC + zq bf16 "; VQ_NHWC=1 timeout 60 python tools/vq_bench.py 2>&1 | tail -1
  echo -n "v1 NHWC + zq bf16 "; VQ_NHWC=1 LVT_VQ_TC1=1 timeout 60 python tools/vq_bench.py 2>&1 | tail -1
  echo -n "v2 NCHW default-init codebook "; VQ_INIT=default timeout 60 python tools/vq_bench.py 2>&1 | tail -1
  echo -n "v2 NCHW 512 frames (the PR-DVQVAE2 step's size) "; VQ_FRAMES=512 timeout 60 python tools/vq_bench.py 2>&1 | tail -1
  echo -n "v1 NCHW 512 frames "; VQ_FRAMES=512 LVT_VQ_TC1=1 timeout 60 python tools/vq_bench.py 2>&1 | tail -1
  echo -n "v2 NCHW, four 128-column accumulator buffers (LVT_VQ_NQ=4) "; LVT_VQ_NQ=4 timeout 60 python tools/vq_bench.py 2>&1 | tail -1
  echo -n "v2 NCHW, all 32 tests of a chunk on the fma pipe (LVT_VQ_NF=32) "; LVT_VQ_NF=32 timeout 60 python tools/vq_bench.py 2>&1 | tail -1
  echo -n "v2 NCHW, no L2 prefetch of the next tile (LVT_VQ_NOPF=1) "; LVT_VQ_NOPF=1 timeout 60 python tools/vq_bench.py 2>&1 | tail -1
} > gpurun_out/${tag}_vq_ab.txt 2>&1
timeout 60 python tools/debug_vq_clock.py > gpurun_out/${tag}_vq_timeline.txt 2>&1
LVT_VQ_TC1=1 timeout 60 python tools/debug_vq_clock.py > gpurun_out/${tag}_vq_timeline_v1.txt 2>&1
WORKLOAD=vqvae timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv \
    --log-file gpurun_out/${tag}_vqvae_launches.csv python tools/profile_step.py > /dev/null 2>&1
python tools/launch_table.py gpurun_out/${tag}_vqvae_launches.csv 40 > gpurun_out/${tag}_vqvae_launches.txt
cat gpurun_out/${tag}_vq_ab.txt; tail -3 gpurun_out/${tag}_vqvae_launches.txt
