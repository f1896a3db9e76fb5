#!/bin/bash
# 2 GPUs: the big-cell recovery check (both exchanges), then the whole sharded parity set
mkdir -p gpurun_out
WHAT="bigcell" bash profiles/mgpu_round.sh 2 r02d_big 2>&1 | grep -v "^==" | tail -12
