// oracle/_ref: src/gpu/tonemap/amd.comp (TEST INFRASTRUCTURE)
#define REF_TM_FN ref_tonemap_amd
#define REF_TM_FILE "tonemap/amd.comp"
#define REF_TM_NPARAMS 5
#include "ref_tonemap.inc"
