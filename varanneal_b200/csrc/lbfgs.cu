// placeholder until the device-resident minimiser lands (next milestone)
#include "vab_ctx.h"
void lbfgs_destroy(vab_ctx*) {}
extern "C" {
int vab_minimize(vab_ctx* ctx, int32_t, double*, int64_t, double, const vab_lbfgs_opts*,
                 const double*, const double*, double*, double*, double*, int32_t*, int32_t*,
                 int32_t*) {
  return vab_fail(ctx, VAB_ERR_STATE, "minimiser not built yet");
}
int vab_anneal(vab_ctx* ctx, int32_t, double*, int64_t, double, const double*, int32_t,
               const vab_lbfgs_opts*, const double*, const double*, double*, double*, int32_t*,
               int32_t*, int32_t*) {
  return vab_fail(ctx, VAB_ERR_STATE, "minimiser not built yet");
}
}
