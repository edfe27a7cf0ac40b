set +e
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mesh.py -q -m gpu -k "shared_scene or frames_in_flight or async_readback" > gpurun_out/pytest_inflight.log 2>&1; echo "inflight tests rc=$?"
tail -30 gpurun_out/pytest_inflight.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_inflight.json 2> gpurun_out/bench_inflight.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_inflight.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_inflight.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','pipelined','gpu_launches')})
PY
