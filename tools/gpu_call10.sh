set +e
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mesh.py -q -m gpu > gpurun_out/pytest_mesh.log 2>&1; echo "mesh tests rc=$?"
grep -E "passed|failed|FAILED|assert " gpurun_out/pytest_mesh.log | head -20
timeout 200 python tools/bench_build.py > gpurun_out/build_lt1024.jsonl 2> gpurun_out/build.err; echo "build bench rc=$?"
grep -v refit gpurun_out/build_lt1024.jsonl | grep '"device_loop": true' | cut -c1-200; tail -3 gpurun_out/build.err
echo lt512; MINOTERT_LIB_DIR=$PWD/variants/lt512 timeout 100 python tools/bench_build.py 2>/dev/null | grep '"device_loop": true' | cut -c1-200
