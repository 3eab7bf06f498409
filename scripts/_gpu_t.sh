timeout 900 python -m pytest tests/test_dsfvt_gpu.py tests/test_api_gpu.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --quick --steps 50 --warmup 5 2>&1 | tail -12 | cut -c1-400
