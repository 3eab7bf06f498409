timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 1500 --csv --log-file gpurun_out/r01d_sampler_launches.csv python tools/sampler_bench.py > /dev/null 2>&1
