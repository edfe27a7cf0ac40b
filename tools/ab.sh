#!/bin/bash
# A/B of mrt_set_option switches / library variants through bench.py on the GPU box: tools/ab.sh <tag> <bench args...>
tag=$1; shift
python bench.py --no-cpu-baseline --steps 20 --warmup 3 "$@" > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err || tail -5 gpurun_out/ab_$tag.err
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{tag}.json").read().strip().splitlines()[-1])
    k = d.get("kernels", {}); r = d.get("roofline", {}); p = d.get("pipelined") or {}
    print(f"{tag:28s} value {d['value']:8.1f} ms/step {d['ms_per_step']:.4f} primary {k.get('primary_ms_per_step', 0):.4f} trace {k.get('trace_ms_per_step', 0):.4f} "
          f"pipelined {p.get('value', 0):8.1f} e2e {d['e2e']['value']:8.1f} nodes/ray {r.get('nodes_per_ray', 0):.2f} launches {d['gpu_launches']}")
except Exception as e:
    print(tag, "FAILED", e)
PY
