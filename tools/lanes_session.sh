#!/bin/bash
# A/B of the two-lane issue: bit-identity tests, then the ZINC headline with lanes off / on
timeout 600 python -m pytest tests/test_gpu_lanes.py -x -q 2>&1 | tail -15 > gpurun_out/lanes_tests.log
for l in off on; do
  timeout 600 python bench.py --lanes $l --no-cpu-baseline > gpurun_out/bench_r1n_zinc_lanes_$l.json 2> gpurun_out/bench_r1n_zinc_lanes_$l.err
done
for w in peptides cifar tsp; do
  for l in off on; do
    timeout 600 python bench.py --workload $w --lanes $l --steps 10 --warmup 3 --pool 2 --no-cpu-baseline > gpurun_out/bench_r1n_${w}_lanes_$l.json 2> gpurun_out/bench_r1n_${w}_lanes_$l.err
  done
done
