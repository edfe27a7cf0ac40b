set +e
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 150 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
tail -4 gpurun_out/san_memcheck_smoke.log
timeout 200 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_temporal.py tests/test_gpu_mesh.py -q -m gpu -x -k "moving_camera_bit_exact or (device_side_build and cornell) or shared_scene or degenerate" > gpurun_out/san_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"
tail -5 gpurun_out/san_memcheck_tests.log
timeout 150 $SAN --tool racecheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"
tail -4 gpurun_out/san_racecheck_smoke.log
