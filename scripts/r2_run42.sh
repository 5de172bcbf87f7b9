#!/bin/bash
# round 2, call 42: the golden inputs (3,369, 1,140 new pairs aimed at the banded table) through the C ABI, SW pairs in both routings
timeout 170 python -m pytest tests/test_reference_golden.py -m gpu -x -q 2>&1 | tail -3
