#!/bin/bash
# GPU session: parity suite, more scatter shapes, where the end-to-end step goes, bench lines of C4 / C2 / C3
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf -p no:cacheprovider -x ) > gpurun_out/pytest_gpu_r02i.log 2>&1
tail -4 gpurun_out/pytest_gpu_r02i.log
( timeout 300 python profiles/ab_scatter.py C4 dense "384,6,4,1 384,6,2,1 384,6,8,1 192,6,4,2 192,6,8,2" 4
  timeout 300 python profiles/ab_scatter.py C2 sparse "256,8,16,1 256,8,8,1 256,8,12,1 256,8,24,1 256,8,30,1" 4 ) > gpurun_out/ab_scatter_r02b.txt 2>&1
cat gpurun_out/ab_scatter_r02b.txt
( MIRROR=1 timeout 200 python profiles/e2e_breakdown.py; MIRROR=0 timeout 200 python profiles/e2e_breakdown.py; AGENTS=12500000 MIRROR=1 timeout 200 python profiles/e2e_breakdown.py ) > gpurun_out/e2e_breakdown_r02.txt 2>&1
cat gpurun_out/e2e_breakdown_r02.txt
for c in C4 C2 C3; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${c}_r02h.json 2> gpurun_out/bench_${c}_r02h.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${c}_r02h.json').read().strip().splitlines()[-1]); print('$c', '%.4g' % d['value'], '%.4f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'], d['e2e'].get('mirror_equals_readback'), d['roofline']['kernels_ms_per_step'])"
done
