timeout 900 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_nee.py tests/test_gpu_sizes.py -x -q -m gpu 2>&1 | tail -2
python tools/check_option.py hall_260k 1921 1079 2 3 prepared_rays=0 2>&1 | tail -2
for w in hall_260k_1080p scene_1m_1080p; do for v in 1 0 1 0; do tools/ab.sh prep${v}_$w --no-extra-configs --workload $w --opt prepared_rays=$v; done; done
tools/ab.sh prep1_10m --no-extra-configs --workload scene_10m_4k --steps 4 --opt prepared_rays=1
tools/ab.sh prep0_10m --no-extra-configs --workload scene_10m_4k --steps 4 --opt prepared_rays=0
