timeout 900 python -m pytest tests/test_api_gpu.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r01c_bench.json 2> gpurun_out/r01c_bench.err; tail -3 gpurun_out/r01c_bench.err; cat gpurun_out/r01c_bench.json
