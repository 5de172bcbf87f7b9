#!/bin/bash
BWT_DEBUG_SYNC=1 python -m pytest tests/test_gpu_packed.py -q -x -k "indel_reads" > gpurun_out/r2d_dbg.log 2>&1; echo "rc=$?"
grep -n "BwtError\|kernel" gpurun_out/r2d_dbg.log | head
BWT_DEBUG_SYNC=1 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/r2d_parity.log 2>&1; echo "parity rc=$?"; tail -30 gpurun_out/r2d_parity.log
