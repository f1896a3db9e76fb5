#!/bin/bash
# GPU session: parity suite on the default and on the new scatter shapes, then the A/B of the shapes (run under gpurun)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf -p no:cacheprovider -x ) > gpurun_out/pytest_gpu_r02h.log 2>&1
tail -4 gpurun_out/pytest_gpu_r02h.log
( QHG_SCATTER_DENSE=192,6,30,2 QHG_SCATTER_SPARSE=128,8,30,2 timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf -p no:cacheprovider -x ) > gpurun_out/pytest_gpu_r02h_v.log 2>&1
tail -4 gpurun_out/pytest_gpu_r02h_v.log
( timeout 300 python profiles/ab_scatter.py C4 dense "384,6,4,1 384,6,16,1 192,6,8,2 192,6,16,2 192,6,30,2 256,5,16,2 128,6,16,3" 4
  timeout 300 python profiles/ab_scatter.py C2 sparse "256,8,4,1 256,8,16,1 128,8,8,2 128,8,16,2 128,8,30,2 192,6,16,2 192,6,30,2" 4
  timeout 300 python profiles/ab_scatter.py C3 sparse "256,8,4,1 256,8,16,1 128,8,16,2 128,8,30,2 192,6,30,2" 4 ) > gpurun_out/ab_scatter_r02.txt 2>&1
cat gpurun_out/ab_scatter_r02.txt
