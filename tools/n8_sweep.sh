#!/bin/bash
# N-GPU sweep of the gradient all-reduce overlap settings (run under gpurun --gpus N): NGPU=8 tools/n8_sweep.sh
N=${NGPU:-8}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
P=29550
for cfg in "LVT_COMM_SMS=16" "LVT_COMM_SMS=0" "LVT_COMM_SMS=0 NCCL_MAX_CTAS=16" "NOOVERLAP=1 LVT_COMM_SMS=0"; do
  P=$((P+1))
  extra=""
  case "$cfg" in NOOVERLAP*) extra="--no-overlap";; esac
  echo "== $cfg"
  env $cfg timeout 300 $R $P bench.py --gpus $N --steps 20 --warmup 5 --quick $extra > /tmp/out.json 2> /tmp/err.log
  python - <<PY
import json
try:
    d = json.loads(open("/tmp/out.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d.get("comm"))
except Exception as e:
    print("FAILED", e); print(open("/tmp/err.log").read()[-800:])
PY
done
