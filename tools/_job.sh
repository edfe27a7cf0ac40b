out=gpurun_out/sanitizer_r2b.txt
echo "compute-sanitizer $(compute-sanitizer --version | tail -1) on $(nvidia-smi -L | head -1)" > $out
run() { echo "== $1" >> $out; shift; timeout 1200 "$@" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninit|error" | head -20 >> $out; }
K1='(wide_refit or fused_sort or device_side) and (tiny or soup or small_terrain)'
run memcheck_build compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_mesh.py -m gpu -q -x -k "$K1"
run racecheck_build compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_mesh.py -m gpu -q -x -k "$K1"
run initcheck_build compute-sanitizer --tool initcheck python -m pytest tests/test_gpu_mesh.py -m gpu -q -x -k "$K1"
run synccheck_build compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_mesh.py -m gpu -q -x -k "$K1"
run memcheck_async compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_mesh.py -m gpu -q -x -k "async_mesh_updates or frames_in_flight"
run memcheck_denoise compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_denoise.py -m gpu -q -x
run racecheck_denoise compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_denoise.py -m gpu -q -x -k "default_params"
cat $out
