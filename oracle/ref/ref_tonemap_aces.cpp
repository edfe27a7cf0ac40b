// oracle/_ref: src/gpu/tonemap/aces.comp (TEST INFRASTRUCTURE)
#define REF_TM_FN ref_tonemap_aces
#define REF_TM_FILE "tonemap/aces.comp"
#define REF_TM_NPARAMS 0
#include "ref_tonemap.inc"
