timeout 900 python -m pytest tests/test_gpu_nee.py tests/test_gpu_frontend.py -m gpu -q -x 2>&1 | tail -4
tools/ab.sh hall_nee2 --no-extra-configs --secondary-flags 8
