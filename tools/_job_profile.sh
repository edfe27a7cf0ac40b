# round-2 final profile captures (run on the GPU box): launch lists + ncu --set full of the kernels DESIGN.md quotes
set +e
tools/profile.sh launches r2_hall
tools/profile.sh launches r2_hall_nee --secondary-flags 8
tools/profile.sh kernel r2_trace_hall "k_trace" 12 2
tools/profile.sh kernel r2_primary_hall "k_mesh_primary" 3 1
tools/profile.sh kernel r2_shade_hall "k_shade" 10 3
tools/profile.sh kernel r2_nee_hall "k_trace|k_shade" 12 6 --secondary-flags 8
tools/profile.sh kernel r2_trace_1m "k_trace" 6 1 --workload scene_1m_1080p
tools/profile.sh kernel r2_trace_10m "k_trace" 24 3 --workload scene_10m_4k --steps 1
ls -la gpurun_out/*.txt gpurun_out/launches_r2_hall*.csv | tail -20
rm -f gpurun_out/*.ncu-rep
for w in "hall" "1m --workload scene_1m_1080p"; do set -- $w; t=$1; shift; tools/ab.sh ${t}_plain --no-extra-configs "$@"; tools/ab.sh ${t}_nee --no-extra-configs --secondary-flags 8 "$@"; tools/ab.sh ${t}_nee_athit --no-extra-configs --secondary-flags 24 "$@"; done
python tools/bench_animated.py > gpurun_out/r2_animated.json 2>/dev/null; tail -1 gpurun_out/r2_animated.json | cut -c1-600
python tools/bench_build.py --scenes hall_260k scene_1m > gpurun_out/r2_build.jsonl 2>/dev/null; grep -c . gpurun_out/r2_build.jsonl
