for v in variants/rf3 variants/rf4 variants/rf8; do for w in hall_260k_1080p scene_1m_1080p; do MINOTERT_LIB_DIR=$v tools/ab.sh $(basename $v)_$w --no-extra-configs --workload $w; done; done
for w in hall_260k_1080p scene_1m_1080p; do tools/ab.sh base_$w --no-extra-configs --workload $w; done
