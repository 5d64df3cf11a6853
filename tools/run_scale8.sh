#!/bin/bash
# 8-GPU weak-scaling runs of every BASELINE.json workload (one process per GPU, NCCL all-reduce of the flat gradient bucket)
for w in zinc peptides cifar tsp; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --steps 10 --warmup 3 --pool 2 --workload $w > gpurun_out/bench_r1_n8_$w.json 2> gpurun_out/bench_r1_n8_$w.err
done
