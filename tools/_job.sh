for v in dnh2 ""; do
  if [ -z "$v" ]; then unset MINOTERT_LIB_DIR; t=base; else export MINOTERT_LIB_DIR=variants/$v; t=$v; fi
  echo "== $t"
  python tools/bench_denoise.py --scene hall 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:v for k,v in d.items() if 'ms' in k or 'taps' in k})"
  python tools/bench_denoise.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:v for k,v in d.items() if 'ms' in k})"
  timeout 600 python -m pytest tests/test_gpu_denoise.py tests/test_gpu_vs_reference.py -m gpu -q -x 2>&1 | tail -3
done
