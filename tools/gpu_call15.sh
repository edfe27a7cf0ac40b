set +e
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_temporal.py -q -m gpu -k "shared or in_flight or temporal or renderer" > gpurun_out/pytest_share2.log 2>&1; echo "share tests rc=$?"
tail -15 gpurun_out/pytest_share2.log
timeout 200 python tools/bench_animated.py > gpurun_out/animated_v2.json 2> gpurun_out/animated_v2.err; echo "animated rc=$?"; cut -c1-700 gpurun_out/animated_v2.json; tail -2 gpurun_out/animated_v2.err
timeout 400 python bench.py --workload scene_10m_4k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v7_10m.json 2> gpurun_out/bench_v7_10m.err; echo "10m rc=$?"; tail -3 gpurun_out/bench_v7_10m.err
