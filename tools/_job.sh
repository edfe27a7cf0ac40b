set +e
( time timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) 2>&1 | tail -12
( time python bench.py > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err ) 2>&1 | tail -4
tail -3 gpurun_out/r2_bench_b.err
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2>&1 | tail -4
nproc
