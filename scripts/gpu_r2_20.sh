#!/bin/bash
# round 2, visit 20 (1 GPU): synccheck of visit 19's script again -- there the tool itself ran out of memory tracking the
# replayed graph launches of the second p123 solve (110 "Internal Sanitizer Error ... Unable to allocate enough memory",
# no barrier finding); here with plain stream launches (PF_GRAPH=0) and with the tool's default barrier table
mkdir -p gpurun_out
sed -n '/^cat > \/tmp\/san5.py/,/^PY$/p' scripts/gpu_r2_19.sh | sed '1d;$d' > /tmp/san5.py
PF_GRAPH=0 timeout 900 compute-sanitizer --tool synccheck --print-limit 10 python /tmp/san5.py > gpurun_out/sanitizer5_synccheck_nograph.log 2>&1
tail -3 gpurun_out/sanitizer5_synccheck_nograph.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 10 python /tmp/san5.py > gpurun_out/sanitizer5_synccheck_graph.log 2>&1
tail -3 gpurun_out/sanitizer5_synccheck_graph.log
