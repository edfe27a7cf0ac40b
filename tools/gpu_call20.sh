set +e
for v in "" tm8 tm16 rf10 mb4; do
  if [ -z "$v" ]; then unset MINOTERT_LIB_DIR; else export MINOTERT_LIB_DIR=$PWD/variants/$v; fi
  echo "== ${v:-default}"; timeout 40 python tools/bench_inflight.py --share 1 --frames 150 --contexts 3 --opt trace_ctas_per_sm=3 2>&1 | tail -1 | cut -c1-190
done
