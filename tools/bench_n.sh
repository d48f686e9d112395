#!/bin/bash
# tools/bench_n.sh <n_gpus> [bench args...]: bench.py the way the driver launches it
N=$1; shift
if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@"; else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; fi
