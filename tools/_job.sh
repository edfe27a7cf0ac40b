MINOTERT_LIB_DIR=variants/empty python tools/count_empty.py 2>&1 | tail -3
MINOTERT_LIB_DIR=variants/nearest timeout 600 python -m pytest tests/test_gpu_mesh.py -x -q -m gpu -k "brute_force or watertight or axis_parallel or render" 2>&1 | tail -2
for w in hall_260k_1080p scene_1m_1080p; do
  tools/ab.sh base_$w --no-extra-configs --workload $w --opt count_visits=1
  MINOTERT_LIB_DIR=variants/nearest tools/ab.sh nearest_$w --no-extra-configs --workload $w --opt count_visits=1
  tools/ab.sh base2_$w --no-extra-configs --workload $w
  MINOTERT_LIB_DIR=variants/nearest tools/ab.sh nearest2_$w --no-extra-configs --workload $w
done
