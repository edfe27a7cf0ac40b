set +e
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_v8.err
timeout 400 python bench.py --workload scene_10m_4k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v7_10m.json 2> gpurun_out/bench_v7_10m.err; echo "10m rc=$?"; tail -3 gpurun_out/bench_v7_10m.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_v8.json','gpurun_out/bench_v7_10m.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'pipe', d['pipelined'] and round(d['pipelined']['value'],1), 'frac', round(d['roofline']['frac'],3), 'build ms', d['config']['bvh_build_ms'])
    except Exception as e: print(f, 'ERR', e)
PY
