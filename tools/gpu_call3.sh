set +e
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mesh.py -q -m gpu -x > gpurun_out/pytest_mesh.log 2>&1; echo "mesh tests rc=$?"
tail -15 gpurun_out/pytest_mesh.log
timeout 200 python tools/bench_build.py > gpurun_out/build.jsonl 2> gpurun_out/build.err; echo "build bench rc=$?"
cat gpurun_out/build.jsonl; tail -3 gpurun_out/build.err
