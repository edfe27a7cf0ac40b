// oracle/_ref: src/gpu/tonemap/linear.comp (TEST INFRASTRUCTURE)
#define REF_TM_FN ref_tonemap_linear
#define REF_TM_FILE "tonemap/linear.comp"
#define REF_TM_NPARAMS 0
#include "ref_tonemap.inc"
