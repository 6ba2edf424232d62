// placeholder until K2/K3 land (same session): keeps every ABI symbol exported
#include "internal.h"
#define NOTYET(ctx) sg_fail(ctx, SLAMGPU_E_STATE, "%s: not implemented yet", __func__)
extern "C" int slamgpu_map_reset_cell(slamgpu_map *m, int32_t, int32_t, const double *) { return NOTYET(m ? m->ctx : nullptr); }
extern "C" int slamgpu_map_update_cell(slamgpu_map *m, int32_t, int32_t, int32_t, double, double, double, double, double) { return NOTYET(m ? m->ctx : nullptr); }
extern "C" int slamgpu_raycast(slamgpu_ctx *ctx, slamgpu_map *, slamgpu_scan *, const double *, int64_t *, int32_t *, int64_t, int64_t *) { return NOTYET(ctx); }
extern "C" int slamgpu_append_scan(slamgpu_ctx *ctx, slamgpu_map *, slamgpu_scan *, const double *, double, int32_t, const slamgpu_estimator *, double, double, const double *, int64_t *) { return NOTYET(ctx); }
