timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q -k "fused_attention" 2>&1 | tail -3
timeout 300 python bench.py --quick --steps 60 --warmup 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step', d['ms_per_step'], d['loss'], d['gpu_launches_per_step'])"
LVT_SPLIT_ATTN=1 timeout 300 python bench.py --quick --steps 60 --warmup 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('split-fwd step', d['ms_per_step'], d['loss'], d['gpu_launches_per_step'])"
