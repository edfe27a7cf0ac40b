set +e
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 500 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/pytest_all.log 2>&1; echo "gpu tests rc=$?"
tail -12 gpurun_out/pytest_all.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --workload scene_1m_1080p --no-cpu-baseline > gpurun_out/bench_v7_1m.json 2> gpurun_out/bench_v7_1m.err; echo "bench 1m rc=$?"
timeout 100 python tools/bench_inflight.py --share 1 --frames 100 > gpurun_out/inflight_share.jsonl 2>&1; echo "inflight rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_v7.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
cut -c1-600 gpurun_out/inflight_share.jsonl
