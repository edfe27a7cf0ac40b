out=gpurun_out/sanitizer_r2d_hall.txt
: > $out
run() { echo "== $1" >> $out; shift; timeout 900 "$@" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninit|error" | head -12 >> $out; }
run memcheck_build_hall compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_mesh.py -m gpu -q -x -k "device_side and hall_260k-ploc"
run racecheck_build_hall compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_mesh.py -m gpu -q -x -k "device_side and hall_260k-ploc"
run initcheck_build_hall compute-sanitizer --tool initcheck python -m pytest tests/test_gpu_mesh.py -m gpu -q -x -k "device_side and hall_260k-ploc"
cat $out
