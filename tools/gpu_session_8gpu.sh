#!/bin/bash
# Multi-GPU evidence in one gpurun call (N GPUs): rank mode under torchrun exactly as the driver launches it (copy-engine
# IPC transport), a small-order run first so that a transport bug shows up in seconds, then NCCL transport for comparison.
N=$1
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
summ() { python -c "import json,sys; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('value %.4g ms %.1f e2e %s check %s | %s' % (d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d['config']['check'], d['config']['panel_transport']))" 2>&1; }
timeout 300 $TR bench.py --gpus $N --order 16384 --steps 1 --warmup 1 --skip-cpu --skip-e2e > $OUT/r02_g${N}_small.json 2> $OUT/r02_g${N}_small.err; echo "small rc=$? $(summ $OUT/r02_g${N}_small.json)"
timeout 420 $TR bench.py --gpus $N --steps 3 --warmup 3 > $OUT/r02_g${N}_rank_ipc.json 2> $OUT/r02_g${N}_rank_ipc.err; echo "rank ipc rc=$? $(summ $OUT/r02_g${N}_rank_ipc.json)"
FW_MULTI_TRANSPORT=nccl timeout 300 $TR bench.py --gpus $N --steps 2 --warmup 2 --skip-cpu --skip-check --skip-e2e > $OUT/r02_g${N}_rank_nccl.json 2> $OUT/r02_g${N}_rank_nccl.err; echo "rank nccl rc=$? $(summ $OUT/r02_g${N}_rank_nccl.json)"
