timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_dsfvt_gpu.py -x -q 2>&1 | tail -3
python tools/gemm_bench.py softmax pv qkv
timeout 300 python bench.py --quick --steps 50 --warmup 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step', d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'], d['gpu_launches_per_step'], d['clocks'])"
