#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/quick_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r1r_zinc.json 2> gpurun_out/bench_r1r_zinc.err
timeout 600 python bench.py --no-cpu-baseline --project-first off > gpurun_out/bench_r1r_zinc_ptt_off.json 2> /dev/null
timeout 600 python bench.py --workload peptides --steps 10 --warmup 3 --pool 2 --no-cpu-baseline > gpurun_out/bench_r1r_peptides.json 2> gpurun_out/bench_r1r_peptides.err
