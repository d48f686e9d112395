#!/bin/bash
# measurement helper: sharded correctness check + per-level profile on N GPUs for several exchange thresholds
#   tools/shard_probe.sh <n_gpus> <keys_per_gpu> <xchg_bits...>
N=$1; K=$2; shift; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for xb in "$@"; do
  echo "=== N=$N keys/GPU=$K BPHF_XCHG_BITS=$xb shard_check"; BPHF_XCHG_BITS=$xb timeout 300 $TR tools/shard_check.py $K 2>&1 | grep -v "^W\|^\*\*\*" | tail -3
  echo "=== N=$N keys/GPU=$K BPHF_XCHG_BITS=$xb shard_prof"; BPHF_XCHG_BITS=$xb timeout 300 $TR tools/shard_prof.py $K 3 2>&1 | grep -v "^W\|^\*\*\*" | tail -10
done
