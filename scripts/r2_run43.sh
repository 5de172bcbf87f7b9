#!/bin/bash
# round 2, call 43: smoke() on the final tree
timeout 150 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
