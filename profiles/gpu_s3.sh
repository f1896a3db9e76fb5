#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf -p no:cacheprovider ) > gpurun_out/pytest_gpu_r02j.log 2>&1
tail -25 gpurun_out/pytest_gpu_r02j.log
( MIRROR=1 timeout 200 python profiles/e2e_breakdown.py ) > gpurun_out/e2e_breakdown_r02b.txt 2>&1
cat gpurun_out/e2e_breakdown_r02b.txt
for c in C4; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${c}_r02j.json 2> gpurun_out/bench_${c}_r02j.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${c}_r02j.json').read().strip().splitlines()[-1]); print('$c', '%.4g' % d['value'], '%.4f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'], d['e2e'].get('mirror_equals_readback'), d['roofline']['kernels_ms_per_step'])"
done
