timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -k "fused_attention" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_dsfvt_gpu.py -x -q 2>&1 | tail -3
for i in 1 2; do timeout 300 python bench.py --quick --steps 60 --warmup 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step', d['ms_per_step'], d['loss'], d['gpu_launches_per_step'])"; done
