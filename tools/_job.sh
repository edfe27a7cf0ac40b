for v in "" top32 top32b top16; do
  if [ -z "$v" ]; then unset MINOTERT_LIB_DIR; t=base; else export MINOTERT_LIB_DIR=variants/$v; t=$v; fi
  tools/ab.sh 10m_$t --no-extra-configs --workload scene_10m_4k --steps 5
  tools/ab.sh 1m_$t --no-extra-configs --workload scene_1m_1080p
  tools/ab.sh hall_$t --no-extra-configs
  python - $t <<'PY'
import json,sys
for w in ("10m","1m","hall"):
    d=json.loads(open(f"gpurun_out/ab_{w}_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("   ", w, "sah", round(d["details"]["sah_node_cost"],2), round(d["details"]["sah_tri_cost"],2), "build ms", round(d["details"]["bvh_build_ms"],3), "bounce nodes/ray", round(d["roofline"]["nodes_per_ray"],2), "all", round(d["roofline"]["nodes_per_ray_all_rays"],2))
PY
done
