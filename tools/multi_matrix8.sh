#!/bin/bash
# 8-GPU (or N-GPU) measurement: rank mode under torchrun (NCCL) and single-process (copy engines), N=65536.
N=$1; TAG=$2
OUT=gpurun_out
summ() { python -c "import json,sys; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('value %.4g ms %.1f e2e %s res %s G %s check %s' % (d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), (d.get('e2e_resident') or {}).get('value'), d['config']['k_blocks_per_bulk_launch'], d['config']['check']))" 2>&1; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 2 --warmup 2 > $OUT/${TAG}_rank.json 2> $OUT/${TAG}_rank.err; echo "rank rc=$? $(summ $OUT/${TAG}_rank.json)"
timeout 600 python bench.py --gpus $N --steps 2 --warmup 2 --single-process --skip-cpu > $OUT/${TAG}_sp.json 2> $OUT/${TAG}_sp.err; echo "sp rc=$? $(summ $OUT/${TAG}_sp.json)"
FW_MULTI_CYCLIC=0 timeout 600 python bench.py --gpus $N --steps 2 --warmup 2 --single-process --skip-cpu --skip-check --skip-e2e > $OUT/${TAG}_sp_contig.json 2> $OUT/${TAG}_sp_contig.err; echo "sp_contig rc=$? $(summ $OUT/${TAG}_sp_contig.json)"
FW_MULTI_GROUP=4 timeout 600 python bench.py --gpus $N --steps 2 --warmup 2 --single-process --skip-cpu --skip-check --skip-e2e > $OUT/${TAG}_sp_g4.json 2> $OUT/${TAG}_sp_g4.err; echo "sp_g4 rc=$? $(summ $OUT/${TAG}_sp_g4.json)"
