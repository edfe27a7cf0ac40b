// oracle/_ref: src/gpu/tonemap/uchimura.comp (TEST INFRASTRUCTURE)
#define REF_TM_FN ref_tonemap_uchimura
#define REF_TM_FILE "tonemap/uchimura.comp"
#define REF_TM_NPARAMS 6
#include "ref_tonemap.inc"
