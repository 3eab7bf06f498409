#!/usr/bin/env bash
# Round-end validation on a GPU box (run through gpurun): parity suite, smoke, bench line, reference arm, ncu launch lists.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_validate.sh r01e'
tag=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -2 gpurun_out/${tag}_bench.err; cut -c1-400 gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1400 -c 433 \
    --csv --log-file gpurun_out/${tag}_step_launches.csv python bench.py --quick --no-graph --steps 2 --warmup 3 > /dev/null 2>&1
