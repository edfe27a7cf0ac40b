set +e
mkdir -p gpurun_out
timeout 100 python tools/bench_temporal.py > gpurun_out/temporal.json 2> gpurun_out/temporal.err; echo "temporal rc=$?"; cat gpurun_out/temporal.json; tail -2 gpurun_out/temporal.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_temporal -s 3 -c 1 -f -o gpurun_out/temporal_full python tools/bench_temporal.py --frames 5 > gpurun_out/ncu_temporal.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_temporal.log
