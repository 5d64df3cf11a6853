#!/bin/bash
# short validation session (1 GPU): tests, smoke, the five workloads, the reference arm.   usage: tools/quick_session.sh <tag>
TAG=${1:-r2q}
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > $O/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/bench_${TAG}_zinc.json 2> $O/bench_${TAG}_zinc.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_zinc_reference.json 2> /dev/null
for w in zinc_default peptides cifar tsp; do
  timeout 400 python bench.py --workload $w --steps 12 --warmup 4 > $O/bench_${TAG}_$w.json 2> $O/bench_${TAG}_$w.err
done
timeout 300 python tools/timeline_probe.py zinc on > $O/${TAG}_timeline_zinc_on.txt 2> /dev/null
tail -2 $O/${TAG}_gpu_tests.log; tail -2 $O/${TAG}_smoke.log
grep -H -o '"value": [0-9.]*, "unit": "graphs/s", "n_gpus"' $O/bench_${TAG}_*.json
grep -H -o '"ms_per_step": [0-9.]*' $O/bench_${TAG}_*.json
