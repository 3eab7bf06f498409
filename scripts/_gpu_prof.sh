mkdir -p gpurun_out
# (1) every launch of one DSFVT step: duration + DRAM bytes (cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1500 -c 481 --csv --log-file gpurun_out/r01c_step_launches.csv python bench.py --quick --no-graph --steps 2 --warmup 3 > gpurun_out/r01c_ncu_bench.log 2>&1
# (2) full captures of the top kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 6 -c 1 -f -o gpurun_out/r01c_gemm_qkv python tools/gemm_bench.py qkv > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 6 -c 1 -f -o gpurun_out/r01c_gemm_softmax python tools/gemm_bench.py softmax > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_argmin_tc -s 3 -c 1 -f -o gpurun_out/r01c_vq_tc python tools/vq_bench.py > /dev/null 2>&1
ls -la gpurun_out/
