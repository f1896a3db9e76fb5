#!/bin/bash
# pairing of small cells with one lane per agent + scatter with the next grab prefetched: parity suite on the new default library
# and on the prefetching shapes, then the A/Bs
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf -p no:cacheprovider -x ) > gpurun_out/pytest_gpu_r02m.log 2>&1
tail -4 gpurun_out/pytest_gpu_r02m.log
( QHG_SCATTER_DENSE=192,6,4,2,1 QHG_SCATTER_SPARSE=128,8,16,2,1 timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf -p no:cacheprovider -x ) > gpurun_out/pytest_gpu_r02m_pf.log 2>&1
tail -4 gpurun_out/pytest_gpu_r02m_pf.log
CONFIGS="C2 C3 C4" bash profiles/try_libs.sh 2>&1 | tail -14
( timeout 300 python profiles/ab_scatter.py C4 dense "384,6,4,1,0 384,6,4,1,1 192,6,4,2,0 192,6,4,2,1 192,6,8,2,1 256,5,4,2,1 128,6,4,3,1" 4
  timeout 300 python profiles/ab_scatter.py C2 sparse "256,8,16,1,0 256,8,16,1,1 128,8,16,2,1 128,8,8,2,1 256,8,12,1,0" 4
  timeout 300 python profiles/ab_scatter.py C3 sparse "256,8,16,1,0 256,8,16,1,1 128,8,16,2,1" 4 ) > gpurun_out/ab_scatter_r02c.txt 2>&1
cat gpurun_out/ab_scatter_r02c.txt
