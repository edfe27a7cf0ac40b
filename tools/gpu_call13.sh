set +e
for o in "" "--opt trace_ctas_per_sm=5" "--opt trace_ctas_per_sm=4" "--opt trace_ctas_per_sm=3" "--opt fused_shade=1" "--opt persistent_primary=1"; do
  echo "== $o"; timeout 60 python tools/bench_inflight.py --share 1 --frames 120 --contexts 3 $o 2>&1 | cut -c1-230
done
echo "== 1m"; timeout 60 python tools/bench_inflight.py --share 1 --frames 120 --contexts 3 --workload scene_1m_1080p 2>&1 | cut -c1-230
echo "== 1m ctas4"; timeout 60 python tools/bench_inflight.py --share 1 --frames 120 --contexts 3 --workload scene_1m_1080p --opt trace_ctas_per_sm=4 2>&1 | cut -c1-230
