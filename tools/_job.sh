set +e
tools/profile.sh launches r2_hall
tools/profile.sh kernel r2_trace_hall "k_trace" 12 2
tools/profile.sh kernel r2_shade_hall "k_shade" 10 3
tools/profile.sh kernel r2_trace_1m "k_trace" 6 1 --workload scene_1m_1080p
tools/profile.sh kernel r2_trace_10m "k_trace" 24 3 --workload scene_10m_4k --steps 1
python tools/agg_launches.py gpurun_out/launches_r2_hall.csv > gpurun_out/launches_r2_hall_summary.txt; tail -30 gpurun_out/launches_r2_hall_summary.txt
rm -f gpurun_out/ncu_r2_trace_*.ncu-rep gpurun_out/ncu_r2_shade_hall.ncu-rep
python tools/bench_build.py --scenes hall_260k scene_1m > gpurun_out/r2_build.jsonl 2>/dev/null; grep -c . gpurun_out/r2_build.jsonl
