set +e
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 300 python -m pytest tests/test_gpu_denoise.py -q -m gpu --durations=5 > gpurun_out/pytest_denoise.log 2>&1; echo "denoise tests rc=$?" 
timeout 150 python tools/bench_denoise.py --cpu > gpurun_out/denoise_base.json 2> gpurun_out/denoise_base.err; echo "dn base rc=$?"
for v in dn16 dn4 dnu4 dnu1; do MINOTERT_LIB_DIR=$PWD/variants/$v timeout 100 python tools/bench_denoise.py > gpurun_out/denoise_$v.json 2>/dev/null; done
timeout 100 python tools/bench_denoise.py --scene hall > gpurun_out/denoise_hall.json 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 420 python -m pytest tests -q -m gpu --deselect tests/test_gpu_denoise.py --durations=8 > gpurun_out/pytest_rest.log 2>&1; echo "rest tests rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_denoise -c 1 -f -o gpurun_out/denoise_full python tools/bench_denoise.py --frames 1 > gpurun_out/ncu_denoise.log 2>&1; echo "ncu rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v6.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
tail -3 gpurun_out/pytest_denoise.log gpurun_out/pytest_rest.log; cat gpurun_out/denoise_*.json | cut -c1-400
