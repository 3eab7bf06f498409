#!/usr/bin/env bash
# Round-2 profile set (run through gpurun, one GPU):  gpurun --timeout 1500 -- 'bash scripts/profile_r02.sh r02'
# launch lists with DRAM traffic of the DSFVT and the PR-DVQVAE2 step, full ncu captures of the fused attention
# backward and of one implicit-GEMM convolution, the attention-backward timeline and the TMEM read micro-benchmark.
tag=${1:-r02}
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv \
    --log-file gpurun_out/${tag}_step_launches.csv python tools/profile_step.py > /dev/null 2>&1
python tools/launch_table.py gpurun_out/${tag}_step_launches.csv 40 --json gpurun_out/${tag}_step_traffic.json > gpurun_out/${tag}_step_launches.txt
WORKLOAD=vqvae timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv \
    --log-file gpurun_out/${tag}_vqvae_launches.csv python tools/profile_step.py > /dev/null 2>&1
python tools/launch_table.py gpurun_out/${tag}_vqvae_launches.csv 40 > gpurun_out/${tag}_vqvae_launches.txt
BATCH=64 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:attn_bwd_kernel -c 1 -o gpurun_out/${tag}_attn_bwd -f python tools/profile_step.py > /dev/null 2>&1
WORKLOAD=vqvae timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:gemm_bf16_kernel -s 2 -c 1 -o gpurun_out/${tag}_conv_gemm -f python tools/profile_step.py > /dev/null 2>&1
timeout 100 python tools/attn_bwd_prof.py > gpurun_out/${tag}_attn_bwd_timeline.txt 2>/dev/null
CAUSAL=1 timeout 100 python tools/attn_bwd_prof.py >> gpurun_out/${tag}_attn_bwd_timeline.txt 2>/dev/null
timeout 60 tools/micro/tmem_bw > gpurun_out/${tag}_tmem_bw.txt 2>&1
timeout 300 python tools/gemm_bench.py qkv ffn ffn_res ffn_mask proj qkv_dgrad ffn_wgrad_auto qkv_wgrad_auto attn_fused_nop attn_bwd_all attn_bwd_all_causal > gpurun_out/${tag}_kernels_in_graph.txt 2>&1
tail -3 gpurun_out/${tag}_step_launches.txt; tail -3 gpurun_out/${tag}_vqvae_launches.txt; ls -la gpurun_out/${tag}_*
