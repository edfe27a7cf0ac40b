set +e
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/build_launches.csv python tools/bench_build.py --scenes scene_1m --repeat 1 > gpurun_out/ncu_build.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/ncu_build.log
