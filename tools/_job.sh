timeout 600 python -m pytest tests/test_gpu_spheres.py -x -q -m gpu 2>&1 | tail -1
ncu --set full --clock-control none -k regex:k_spheres -s 8 -c 2 -f -o gpurun_out/ncu_r2_spheres python bench.py --workload spheres_960x540 --no-extra-configs --no-cpu-baseline --steps 2 --warmup 3 --frames-in-flight 1 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/ncu_r2_spheres.ncu-rep > gpurun_out/ncu_r2_spheres.txt; grep -E "^##|duration|warp execution|issue-slot|warp instructions" gpurun_out/ncu_r2_spheres.txt
python bench.py --workload spheres_960x540 --no-extra-configs --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('final', d['value'], d['ms_per_step'], d['e2e']['value'])"
