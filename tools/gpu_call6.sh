set +e
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mesh.py -q -m gpu -k "device_side or soup" > gpurun_out/pytest_mesh.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|assert" gpurun_out/pytest_mesh.log | head -30
