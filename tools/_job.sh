for v in tile4 tile16 tile32 ""; do
  if [ -z "$v" ]; then unset MINOTERT_LIB_DIR; t=base; else export MINOTERT_LIB_DIR=variants/$v; t=$v; fi
  python tools/check_option.py hall_260k 1920 1080 1 1 2>&1 | tail -1
  tools/ab.sh hall_$t --no-extra-configs; tools/ab.sh 1m_$t --no-extra-configs --workload scene_1m_1080p
  tools/ab.sh 10m_$t --no-extra-configs --workload scene_10m_4k --steps 4
done
