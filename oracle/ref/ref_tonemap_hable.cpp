// oracle/_ref: src/gpu/tonemap/hable.comp (TEST INFRASTRUCTURE)
#define REF_TM_FN ref_tonemap_hable
#define REF_TM_FILE "tonemap/hable.comp"
#define REF_TM_NPARAMS 0
#include "ref_tonemap.inc"
