#!/bin/bash
# developer aid: N = 2 bench variants on a 2-GPU box (diagnostics of the non-kernel time per step)
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline "${@:3}" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); k=d['device_ms_per_step']
        print('$2', 'ms/step %.1f kernels %.1f' % (d['ms_per_step'], sum(k.values())), 'e2e %.3f f32 %.3f pageable %.3f Gvox/s' % (d['e2e']['value']/1e9, d['e2e_float32']['value']/1e9, d['e2e_pageable']['value']/1e9))
"; }
run 29521 "default          "
run 29522 "no stats         " --no-stats
B4D_EXCHANGE_OVERLAP=0 run 29523 "no overlap       "
run 29524 "no exchange      " --no-exchange
