for i in 1 2 3; do timeout 900 python -m pytest tests/test_vqvae_gpu.py -x -q -k graphed 2>&1 | grep -E "^E|passed|failed" | head -8; done
