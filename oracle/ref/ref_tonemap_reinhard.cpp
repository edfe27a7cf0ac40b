// oracle/_ref: src/gpu/tonemap/reinhard.comp (TEST INFRASTRUCTURE)
#define REF_TM_FN ref_tonemap_reinhard
#define REF_TM_FILE "tonemap/reinhard.comp"
#define REF_TM_NPARAMS 1
#include "ref_tonemap.inc"
