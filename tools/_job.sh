timeout 900 python -m pytest tests/test_gpu_group.py -m gpu -q -x -k "frame" 2>&1 | tail -5
python bench.py --workload tiles_4k_progressive --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/tiles1_fif6.json 2> gpurun_out/tiles1_fif6.err || tail -20 gpurun_out/tiles1_fif6.err
python - <<'PY'
import json
for t in ("fif6",):
    try:
        d=json.loads([l for l in open(f"gpurun_out/tiles1_{t}.json") if l.startswith("{")][-1])
        print(t, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["details"]["framebuffer_sha256_16"], "roof", round(d["roofline"]["frac"],3))
    except Exception as e: print(t,"FAILED",e)
PY
