#!/bin/bash
# round 2, call 26: sw_double_kernel with DMNMX for the maxima, masked / light strip variants
python -m pytest tests/test_gpu_parity.py tests/test_gpu_packed.py tests/test_reference_golden.py -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2z_tests.log
python scripts/n_reads_probe.py 2>&1 | tail -1
