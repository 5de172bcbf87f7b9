#!/bin/bash
# debugging call: all gpu tests (no -x), then the failing test under memcheck
python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r2c_tests.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_packed.py -q -k "indel_reads" > gpurun_out/r2c_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -n "Invalid\|at 0x\|by thread\|Address\|ERROR SUMMARY\|kernel\|\.cu" gpurun_out/r2c_memcheck.log | head -40
