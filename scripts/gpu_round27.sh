#!/bin/bash
# r01 round 27: last full GPU suite + smoke of the session (after the host-side parser / writer changes)
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
