#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2e_tests.log
