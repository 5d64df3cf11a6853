#!/bin/bash
# end-of-round measurement session (1 GPU): tests, smoke, every workload, reference arm, dense shape tables, timeline,
# ncu launch list of the bench command, ncu --set full of the persistent GEMM (each ncu run only after the same command
# has exited 0 without it).   usage: tools/final_session.sh <tag>
TAG=${1:-r2z}
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > $O/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/bench_${TAG}_zinc.json 2> $O/bench_${TAG}_zinc.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_zinc_reference.json 2> /dev/null
for w in zinc_default peptides cifar tsp; do
  timeout 400 python bench.py --workload $w --steps 12 --warmup 4 > $O/bench_${TAG}_$w.json 2> $O/bench_${TAG}_$w.err
done
for w in zinc cifar tsp peptides; do
  timeout 300 python tools/dense_shapes_probe.py $w > $O/${TAG}_dense_shapes_$w.txt 2> /dev/null
done
timeout 300 python tools/timeline_probe.py zinc on > $O/${TAG}_timeline_zinc_on.txt 2> /dev/null
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
timeout 120 python tools/gemm_one.py 24144 256 1408 > $O/${TAG}_plain_gemm.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k gemm_tf32x3_persistent_kernel -s 2 -c 1 -f -o $O/${TAG}_gemm_ps_k1408 \
    python tools/gemm_one.py 24144 256 1408 > $O/${TAG}_ncu_gemm.log 2>&1
tail -3 $O/${TAG}_gpu_tests.log; tail -2 $O/${TAG}_smoke.log
grep -H -o '"value": [0-9.]*, "unit": "graphs/s", "n_gpus"' $O/bench_${TAG}_*.json
