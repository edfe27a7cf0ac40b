timeout 300 python -m pytest tests/test_gpu_mesh.py -m gpu -q -x -k "path_kernel or ray_sort" 2>&1 | tail -4
for w in "hall" "1m --workload scene_1m_1080p" "10m --workload scene_10m_4k --steps 5"; do set -- $w; t=$1; shift; tools/ab.sh ${t}_pk1 --no-extra-configs "$@" --opt path_kernel=1; done
for v in 16 28; do MINOTERT_LIB_DIR=variants/shade$v tools/ab.sh hall_pk1_s$v --no-extra-configs --opt path_kernel=1; done
tools/ab.sh hall_sort2 --no-extra-configs --opt sort_rays=2
timeout 600 python tools/bench_slab.py scene_10m 2>/dev/null
