// placeholder until K4/K5 land (same session): keeps every ABI symbol exported
#include "internal.h"
#define NOTYET(ctx) sg_fail(ctx, SLAMGPU_E_STATE, "%s: not implemented yet", __func__)
extern "C" int slamgpu_pyramid_create(slamgpu_ctx *ctx, slamgpu_map *, int32_t, slamgpu_pyramid **) { return NOTYET(ctx); }
extern "C" void slamgpu_pyramid_destroy(slamgpu_pyramid *) {}
extern "C" int slamgpu_pyramid_levels(slamgpu_pyramid *) { return NOTYET(nullptr); }
extern "C" int slamgpu_pyramid_level_info(slamgpu_pyramid *, int32_t, int32_t *, int32_t *, double *, int32_t *, int32_t *) { return NOTYET(nullptr); }
extern "C" int slamgpu_pyramid_build(slamgpu_pyramid *) { return NOTYET(nullptr); }
extern "C" int slamgpu_pyramid_level_download(slamgpu_pyramid *, int32_t, double *, double *) { return NOTYET(nullptr); }
extern "C" int slamgpu_pyramid_rescale(slamgpu_pyramid *, double) { return NOTYET(nullptr); }
extern "C" int slamgpu_pyramid_append_scan(slamgpu_pyramid *, slamgpu_scan *, const double *, double, int32_t, const slamgpu_estimator *, double, double, const double *, int64_t *) { return NOTYET(nullptr); }
extern "C" int slamgpu_score_windows(slamgpu_pyramid *, slamgpu_scan *const *, int32_t, const int32_t *, const double *, int64_t, const double *, const slamgpu_spe_params *, double *) { return NOTYET(nullptr); }
