#!/bin/bash
# Runs bench.py for the in-tree build and every variants/<name>; prints one summary line per run.
# Usage (on the GPU box): tools/sweep.sh <workload> [steps]
WL=${1:-hall_260k_1080p}; STEPS=${2:-10}
run() {
  local name=$1 dir=$2
  MINOTERT_LIB_DIR=$dir timeout 300 python bench.py --workload $WL --steps $STEPS --warmup 3 --no-cpu-baseline $EXTRA_ARGS 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('%-10s %-16s value=%8.1f Mrays/s  ms/step=%.3f  primary_ms=%.3f trace_ms=%.3f e2e=%8.1f nodes/ray=%.2f tris/ray=%.2f wide=%d build_ms=%.1f sah=%.1f+%.1f' % ('$name','$WL',d['value'],d['ms_per_step'],d['kernels']['primary_ms_per_step'],d['kernels']['trace_ms_per_step'],d['e2e']['value'],r['nodes_per_ray'],r['tris_per_ray'],d['config']['wide_nodes'],d['config']['bvh_build_ms'],d['config']['sah_node_cost'],d['config']['sah_tri_cost']))" || echo "$name FAILED"
}
run base "$(pwd)/minotert_b200"
shopt -s nullglob
for d in variants/*/; do run "$(basename $d)" "$(pwd)/$d"; done
