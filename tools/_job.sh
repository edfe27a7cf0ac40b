out=gpurun_out/sanitizer_r2c_spheres.txt
: > $out
run() { echo "== $1" >> $out; shift; timeout 1200 "$@" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninit|error" | head -12 >> $out; }
KS='batched and not size0'
run memcheck_spheres compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_spheres.py -m gpu -q -x -k "$KS"
run racecheck_spheres compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_spheres.py -m gpu -q -x -k "$KS"
run synccheck_spheres compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_spheres.py -m gpu -q -x -k "$KS"
run initcheck_spheres compute-sanitizer --tool initcheck python -m pytest tests/test_gpu_spheres.py -m gpu -q -x -k "$KS"
cat $out
