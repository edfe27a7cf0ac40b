python tools/check_option.py hall_260k 1920 1080 2 3 ray_split=50 2>&1 | tail -7
for c in 0 30 50 70 90; do
  tools/ab.sh hall_s$c --no-extra-configs --opt ray_split=$c; tools/ab.sh 1m_s$c --no-extra-configs --workload scene_1m_1080p --opt ray_split=$c
done
tools/ab.sh 10m_s0 --no-extra-configs --workload scene_10m_4k --steps 4
tools/ab.sh 10m_s50 --no-extra-configs --workload scene_10m_4k --steps 4 --opt ray_split=50
