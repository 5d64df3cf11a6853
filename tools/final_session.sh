#!/bin/bash
# end-of-round measurement session (1 GPU): tests, smoke, every workload, reference arm, launch list, ncu capture of the GEMM
TAG=${1:-r1s}
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py > gpurun_out/bench_${TAG}_zinc.json 2> gpurun_out/bench_${TAG}_zinc.err
python bench.py --lanes off --no-cpu-baseline > gpurun_out/bench_${TAG}_zinc_lanes_off.json 2> /dev/null
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_zinc_reference.json 2> /dev/null
for w in zinc_default peptides cifar tsp; do
  python bench.py --workload $w --steps 10 --warmup 3 --pool 2 > gpurun_out/bench_${TAG}_$w.json 2> gpurun_out/bench_${TAG}_$w.err
done
python tools/gemm_shapes_probe.py > gpurun_out/gemm_shapes_${TAG}.log 2>&1
python tools/wgrad_shapes_probe.py > gpurun_out/wgrad_shapes_${TAG}.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/gemm_one.py 24144 256 1408 > gpurun_out/plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 --launch-skip 2 -c 2 -o gpurun_out/gemm_${TAG} python tools/gemm_one.py 24144 256 1408 > gpurun_out/ncu_gemm.log 2>&1
