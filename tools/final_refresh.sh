#!/bin/bash
# short end-of-round refresh (no ncu): tests, smoke, headline with CPU baseline, reference arm, the long-row workloads
TAG=${1:-r1u}
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py > gpurun_out/bench_${TAG}_zinc.json 2> gpurun_out/bench_${TAG}_zinc.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_zinc_reference.json 2> /dev/null
for w in cifar tsp; do
  python bench.py --workload $w --steps 10 --warmup 3 --pool 2 --no-cpu-baseline > gpurun_out/bench_${TAG}_$w.json 2> gpurun_out/bench_${TAG}_$w.err
done
