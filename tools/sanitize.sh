#!/bin/bash
# compute-sanitizer passes over smoke() and the small GPU tests (run on the GPU box); output -> gpurun_out/sanitizer_r2.txt
out=gpurun_out/sanitizer_r2.txt
echo "compute-sanitizer $(compute-sanitizer --version | tail -1) on $(nvidia-smi -L | head -1)" > $out
run() { echo "== $1" >> $out; shift; timeout 900 "$@" 2>&1 | grep -E "passed|failed|smoke ok|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|error" | head -20 >> $out; }
SMOKE='import __graft_entry__ as g; g.smoke()'
SMALL='tests/test_gpu_nee.py tests/test_gpu_group.py tests/test_gpu_temporal.py'
KSEL='sky_extensions or nee_image or aerial or frames_in_flight_tiles or frame_sum or tiles_p2p'
run memcheck_smoke compute-sanitizer --tool memcheck python -c "$SMOKE"
run memcheck_tests compute-sanitizer --tool memcheck python -m pytest $SMALL -m gpu -q -x -k "$KSEL"
run racecheck_smoke compute-sanitizer --tool racecheck python -c "$SMOKE"
run racecheck_tests compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_nee.py -m gpu -q -x -k "sky_extensions or aerial"
run initcheck_smoke compute-sanitizer --tool initcheck python -c "$SMOKE"
run synccheck_smoke compute-sanitizer --tool synccheck python -c "$SMOKE"
cat $out
