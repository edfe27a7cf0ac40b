timeout 600 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_sizes.py -m gpu -q -x -k "not 10m" 2>&1 | tail -4
for w in "hall" "1m --workload scene_1m_1080p" "10m --workload scene_10m_4k --steps 5"; do set -- $w; t=$1; shift; tools/ab.sh ${t}_ffma2 --no-extra-configs "$@"; MINOTERT_LIB_DIR=variants/noffma2 tools/ab.sh ${t}_scalar --no-extra-configs "$@"; done
