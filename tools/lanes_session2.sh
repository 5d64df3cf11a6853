#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_graphed_step.py -x -q 2>&1 | tail -15 > gpurun_out/lanes_tests2.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r1p_zinc.json 2> gpurun_out/bench_r1p_zinc.err
for w in peptides; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --pool 2 --no-cpu-baseline > gpurun_out/bench_r1p_${w}.json 2> gpurun_out/bench_r1p_${w}.err
done
