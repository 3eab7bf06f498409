#!/bin/bash
# 2-GPU sweep of the gradient all-reduce overlap settings (run under gpurun --gpus 2)
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
P=29520
for cfg in "LVT_COMM_SMS=0 LVT_GRAD_PARTS=8" "LVT_COMM_SMS=4 LVT_GRAD_PARTS=8" "LVT_COMM_SMS=16 LVT_GRAD_PARTS=8" "LVT_COMM_SMS=0 LVT_GRAD_PARTS=8 NCCL_MAX_CTAS=8" "LVT_COMM_SMS=0 LVT_GRAD_PARTS=8 NCCL_MAX_CTAS=16"; do
  P=$((P+1))
  extra=""
  if [ "$cfg" = "NOOVERLAP=1" ]; then extra="--no-overlap"; fi
  echo "== $cfg"
  env $cfg timeout 300 $R $P bench.py --gpus 2 --steps 20 --warmup 5 --quick $extra > /tmp/out.json 2> /tmp/err.log
  python - <<PY
import json
try:
    d = json.loads(open("/tmp/out.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d.get("comm"))
except Exception as e:
    print("FAILED", e)
    print(open("/tmp/out.json").read()[-500:])
    print(open("/tmp/err.log").read()[-1500:])
PY
done
timeout 100 python bench.py --quick 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n1', d['ms_per_step'])"
