#!/bin/bash
# round 2, call 27: the full GPU suite once more (reference-golden test now also holds the interpreted outputs) and smoke()
python -m pytest tests -m gpu -x -q > gpurun_out/r2fin2_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2fin2_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
