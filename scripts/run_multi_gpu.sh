#!/bin/bash
# usage: scripts/run_multi_gpu.sh N [extra bench args]   -- bench.py under torchrun, one rank per GPU
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"
