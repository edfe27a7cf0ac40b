set +e
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_denoise.py -q -m gpu > gpurun_out/pytest_denoise.log 2>&1; echo "denoise tests rc=$?"
tail -5 gpurun_out/pytest_denoise.log
timeout 150 python tools/bench_denoise.py --cpu > gpurun_out/denoise_base.json 2> gpurun_out/denoise_base.err; echo "dn base rc=$?"
for v in dnu4 dnu1 dn4u1; do MINOTERT_LIB_DIR=$PWD/variants/$v timeout 100 python tools/bench_denoise.py > gpurun_out/denoise_$v.json 2>/dev/null; done
cat gpurun_out/denoise_*.json | cut -c1-330
