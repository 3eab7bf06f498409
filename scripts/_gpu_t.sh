for i in 1 2; do
LVT_SPLIT_ATTN=1 timeout 300 python bench.py --quick --steps 60 --warmup 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('split', d['ms_per_step'], d['gpu_launches_per_step'])"
timeout 300 python bench.py --quick --steps 60 --warmup 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused', d['ms_per_step'], d['gpu_launches_per_step'])"
done
