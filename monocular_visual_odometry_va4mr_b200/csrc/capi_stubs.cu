// Entry points declared in include/b200vo.h whose kernels are not written yet.  They fail
// loudly (B200VO_E_UNSUPPORTED) -- there is no CPU fallback behind any of them.
#include "internal.cuh"
#define STUB(ctx) return vo_set_err(ctx, B200VO_E_UNSUPPORTED, "%s: not implemented yet", __func__)
extern "C" int b200vo_find_essential_mat_ransac(b200vo_ctx* ctx, const float*, const float*, int, const double*, double, double, int, double*, uint8_t*, int*) { STUB(ctx); }
