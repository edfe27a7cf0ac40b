python tools/check_option.py hall_260k 1921 1079 2 3 shade_tiles=0 2>&1 | tail -2
CHECK_FLAGS=24 python tools/check_option.py hall_260k 1280 720 2 2 shade_tiles=0 2>&1 | tail -1
for c in 1 0; do
  tools/ab.sh hall_t$c --no-extra-configs --opt shade_tiles=$c; tools/ab.sh 1m_t$c --no-extra-configs --workload scene_1m_1080p --opt shade_tiles=$c
  tools/ab.sh 10m_t$c --no-extra-configs --workload scene_10m_4k --steps 4 --opt shade_tiles=$c
done
