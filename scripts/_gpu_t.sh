for i in 1 2 3 4; do timeout 300 python -m pytest tests/test_vqvae_gpu.py -x -q 2>&1 | tail -1; done
