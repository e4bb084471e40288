#!/bin/bash
# 2 GPUs, one short visit: bench.py --parallelism sharded (configs[4] mode) at a size that takes seconds --
# 64 MiB byte corpus in 16 Mi-row data blocks (5 blocks), built by both ranks side by side (build_dist),
# loaded as two shards, counted through the device-initiated exchange, checked against the reference.
mkdir -p gpurun_out
timeout 48 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --parallelism sharded --kind bytes --corpus-mib 64 --block-rows-log2 24 --npats 262144 \
  --steps 3 --warmup 2 --cpu-sample-seconds 2 > gpurun_out/r02_sharded_mode_small.json 2> gpurun_out/r02_sharded_mode_small.log
echo "rc=$?"
tail -c 1500 gpurun_out/r02_sharded_mode_small.json
tail -n 12 gpurun_out/r02_sharded_mode_small.log
