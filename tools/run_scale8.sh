#!/bin/bash
# All BASELINE.json configurations at 1 and 8 GPUs (+ 2 and 4 for the headline): bench lines into gpurun_out/scale_r2_*.json
# usage (8-GPU box): bash tools/run_scale8.sh
port=29600
run() {  # workload gpus steps
  port=$((port + 1))
  if [ "$2" = "1" ]; then
    timeout 500 python bench.py --workload $1 --gpus 1 --steps $3 --warmup 5 --no-cpu-baseline > gpurun_out/scale_r2_$1_n1.json 2> gpurun_out/scale_r2_$1_n1.err
  else
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $port bench.py --workload $1 --gpus $2 --steps $3 --warmup 5 > gpurun_out/scale_r2_$1_n$2.json 2> gpurun_out/scale_r2_$1_n$2.err
  fi
}
for n in 1 2 4 8; do run zinc $n 30; done
for w in peptides cifar tsp zinc_default; do run $w 1 15; run $w 8 15; done
python - <<'EOF'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_r2_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("scale_r2_")[1][:-5], d["n_gpus"], round(d["value"], 1), "graphs/s", round(d["ms_per_step"], 3), "ms/step  e2e", round(d["e2e"]["value"], 1))
    except Exception as e:
        print(f, "ERR", e)
EOF
