#!/bin/bash
# Multi-GPU measurement matrix (run under gpurun --gpus N): transports, row layouts and group sizes of fw_multi.
# usage: tools/multi_matrix.sh NGPU ORDER TAG
N=$1; ORDER=$2; TAG=$3
OUT=gpurun_out
run() {  # name, env..., -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py --gpus $N --order $ORDER --steps 2 --warmup 2 --skip-cpu "$@" > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.loads(open('$OUT/${TAG}_${name}.json').read().strip().splitlines()[-1]); print('value %.4g ms %.1f e2e %s G %s' % (d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d['config']['k_blocks_per_bulk_launch']))" 2>&1)"
}
run sp_p2p_cyc  FW_MULTI_TRANSPORT=p2p  -- --single-process --skip-check
run sp_nccl_cyc FW_MULTI_TRANSPORT=nccl -- --single-process --skip-check --skip-e2e
run sp_p2p_contig FW_MULTI_TRANSPORT=p2p FW_MULTI_CYCLIC=0 -- --single-process --skip-check --skip-e2e
run sp_p2p_g2   FW_MULTI_TRANSPORT=p2p FW_MULTI_GROUP=2 -- --single-process --skip-check --skip-e2e
run sp_p2p_g4   FW_MULTI_TRANSPORT=p2p FW_MULTI_GROUP=4 -- --single-process --skip-check --skip-e2e
