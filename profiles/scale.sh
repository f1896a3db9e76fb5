#!/bin/bash
# strong-scaling run of bench.py at N GPUs of one box: bash profiles/scale.sh N  -> gpurun_out/scale_N.json
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/scale_$N.err | tail -1 > gpurun_out/scale_$N.json
python -c "
import json; d=json.load(open('gpurun_out/scale_$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks']); print(d['roofline']['kernels_ms_per_step'])"
