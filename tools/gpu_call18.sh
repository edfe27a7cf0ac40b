set +e
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_denoise.py tests/test_gpu_spheres.py -q -m gpu -k "tonemap or denoise or render_cornell or renderer" > gpurun_out/pytest_tm.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/pytest_tm.log
timeout 200 python tools/bench_animated.py > gpurun_out/animated_v3.json 2> gpurun_out/animated_v3.err; echo "animated rc=$?"; cut -c1-600 gpurun_out/animated_v3.json; tail -2 gpurun_out/animated_v3.err
