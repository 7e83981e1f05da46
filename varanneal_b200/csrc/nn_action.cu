// placeholder until the NN kernels land (next milestone)
#include "vab_ctx.h"
void nn_destroy(vab_ctx*) {}
long long nn_unknowns(const vab_ctx*) { return 0; }
int nn_eval(vab_ctx* ctx, int, const double*, long long, double, const int*, double*, double*,
            double*, double*, long long) {
  return vab_fail(ctx, VAB_ERR_STATE, "NN path not built yet");
}
extern "C" {
int vab_nn_problem_set(vab_ctx* ctx, int32_t, const int32_t*, int32_t, int32_t, int32_t,
                       const int32_t*, int32_t, const int32_t*, const double*, const double*,
                       int32_t, const int32_t*) {
  return vab_fail(ctx, VAB_ERR_STATE, "NN path not built yet");
}
int vab_nn_set_weights(vab_ctx* ctx, double, double, double) { return vab_fail(ctx, VAB_ERR_STATE, "NN path not built yet"); }
int vab_nn_set_fixed_params(vab_ctx* ctx, const double*, int64_t) { return vab_fail(ctx, VAB_ERR_STATE, "NN path not built yet"); }
int vab_nn_action_grad(vab_ctx* ctx, int32_t, const double*, int64_t, double, double*, double*,
                       double*, double*, int64_t) {
  return vab_fail(ctx, VAB_ERR_STATE, "NN path not built yet");
}
}
