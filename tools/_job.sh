set +e
python tools/check_option.py hall_260k 1920 1080 1 2 primary_batched=1 2>&1 | tail -7
tools/ab.sh hall_base --no-extra-configs
tools/ab.sh hall_pb --no-extra-configs --opt primary_batched=1
tools/ab.sh 1m_base --no-extra-configs --workload scene_1m_1080p
tools/ab.sh 1m_pb --no-extra-configs --workload scene_1m_1080p --opt primary_batched=1
tools/ab.sh hall_b2c3 --no-extra-configs --opt bands=2 --opt trace_ctas_per_sm=3
tools/ab.sh hall_b3c2 --no-extra-configs --opt bands=3 --opt trace_ctas_per_sm=2
tools/ab.sh 10m_base --no-extra-configs --workload scene_10m_4k --steps 5
tools/ab.sh 10m_b2c3 --no-extra-configs --workload scene_10m_4k --steps 5 --opt bands=2 --opt trace_ctas_per_sm=3
