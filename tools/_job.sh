for c in -1 38 44 58 30 100; do
  tools/ab.sh hall_c$c --no-extra-configs --opt trace_carveout=$c; tools/ab.sh 1m_c$c --no-extra-configs --workload scene_1m_1080p --opt trace_carveout=$c
done
tools/ab.sh 10m_c-1 --no-extra-configs --workload scene_10m_4k --steps 4 --opt trace_carveout=-1
tools/ab.sh 10m_c38 --no-extra-configs --workload scene_10m_4k --steps 4 --opt trace_carveout=38
tools/ab.sh 10m_c44 --no-extra-configs --workload scene_10m_4k --steps 4 --opt trace_carveout=44
