#!/bin/bash
# ncu over the kernels that only a sharded build runs: 2 GPUs, exchange path forced onto levels 0 and 1 (per-GPU load =
# what each GPU of an 8 x 1e8-key build sees: 1e8 keys routed, 1e8 offsets marked on a 2 x 21 MB slice).  Only rank 0
# runs under ncu (its kernels are replayed ~40 times each); rank 1 runs free and waits at the device barriers, whose
# timeout is raised for the occasion.   tools/ncu_sharded.sh <out_prefix>
export NCU_OUT=${1:-gpurun_out/prof_r2_sharded}
export BPHF_XCHG_BITS=1e8
export BPHF_BARRIER_TIMEOUT_MS=600000
cat > /tmp/ncu_rank_wrap.sh <<'WRAP'
#!/bin/bash
if [ "$LOCAL_RANK" = "0" ]; then
  exec ncu --set full --clock-control none --import-source on \
       -k regex:"xscatter_kernel|xmark_kernel|xbits_kernel|pass_x_kernel|or_collide_merge_kernel" -c 7 -o $NCU_OUT python "$@"
else
  exec python "$@"
fi
WRAP
chmod +x /tmp/ncu_rank_wrap.sh
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 --no-python \
    /tmp/ncu_rank_wrap.sh tools/shard_loop.py 1e8 1 2>&1 | tail -6
