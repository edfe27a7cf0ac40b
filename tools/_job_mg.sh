N=$1
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -q -x 2>&1 | tail -25; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 16 --warmup 3 > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err || tail -20 gpurun_out/r2_scale_$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2_scale_{n}.json") if l.startswith("{")][-1])
    print("N", d["n_gpus"], "value", round(d["value"], 1), "ms/frame", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "sha", d["details"]["framebuffer_sha256_16"], d["clocks"]["reasons"])
except Exception as e:
    print("FAILED", e)
PY
