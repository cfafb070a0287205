#!/bin/bash
# round 2, last GPU minutes: the GPU suite and the smoke run at the final commit
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02_last_tests.txt
timeout 40 python __graft_entry__.py smoke 2>&1 | tail -1 | tee -a gpurun_out/r02_last_tests.txt
