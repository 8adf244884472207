/* Plain-C caller of libgpr_sm100a.so: proves include/gpr_sm100a.h is self-contained C99 and that the ABI can be
 * driven with nothing but pointers and sizes (no Python, no torch, no C++).
 *
 *   gcc -std=c99 -pedantic -Wall -Werror -I include tests/cabi/test_cabi.c -L <pkg> -lgpr_sm100a -lm -o test_cabi
 *
 * Known answer (the reference's own test, /root/reference/test/test_loss.jl:1-11): for a diagonal covariance
 * K = d I the NLML is 0.5 (y'y / d + n log d + n log 2 pi).  A SquaredExp() with a huge inverse length scale on
 * integer-spaced points is diagonal to machine precision: d = sigma^2 + eps + sigma_n^2, every gradient component
 * follows in closed form, the predictive mean at the training points is y sigma^2 / d (x !== md.x: no jitter in K*)
 * and its variance sigma^2 + sigma_n^2 - sigma^4 / d  (src/predict.jl:55-58,89-95).
 *
 * Exit code 0 = all checks passed, 77 = no usable sm_100 device (create failed with GPR_ERR_CUDA: there is no CPU
 * fallback), anything else = failure.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "gpr_sm100a.h"

#define N 300
#define CHECK(cond, msg)                                         \
  do {                                                           \
    if (!(cond)) { fprintf(stderr, "FAIL: %s (line %d)\n", msg, __LINE__); return 1; } \
  } while (0)

static double relerr(double a, double b) { return fabs(a - b) / fmax(fabs(b), 1e-300); }

int main(void) {
  const double PI = 3.14159265358979323846;
  gpr_ctx* ctx = NULL;
  gpr_model* model = NULL;
  int types[2] = {GPR_KERN_SE, GPR_KERN_NOISE};
  double x[N], y[N], mean[N], var[N], G[3], F = 0.0, ms[GPR_T_COUNT];
  double hp[3] = {0.8, 50.0, 0.3};
  const double eps = 1e-8;
  const double d = hp[0] * hp[0] + eps + hp[2] * hp[2];
  double yy = 0.0, Fref, g_sigma, g_noise;
  int64_t info = -1;
  int i, rc;

  CHECK(gpr_version() >= 100, "gpr_version");
  CHECK(gpr_dim_hp(types, 2, 1) == 3, "gpr_dim_hp");
  rc = gpr_ctx_create(0, &ctx);
  if (rc == GPR_ERR_CUDA) {
    printf("no usable sm_100 device: %s\n", gpr_last_error(NULL));
    return 77;
  }
  CHECK(rc == GPR_OK && ctx != NULL, "gpr_ctx_create");

  for (i = 0; i < N; ++i) {
    x[i] = (double)i;
    y[i] = sin(0.37 * i) + 0.01 * (double)((i * 7919) % 13);
    yy += y[i] * y[i];
  }
  rc = gpr_model_create(ctx, types, 2, 1, N, x, y, 1, 1, &model);
  CHECK(rc == GPR_OK && model != NULL, gpr_last_error(ctx));
  CHECK(gpr_model_create(ctx, types, 2, 1, N, x, y, 1, 2, &model) == GPR_ERR_ARG, "train_axis out of range must be rejected");
  rc = gpr_model_create(ctx, types, 2, 1, N, x, y, 1, 1, &model);
  CHECK(rc == GPR_OK, "second model");

  /* loss_grad! (src/cost.jl:50-58) */
  rc = gpr_nlml_grad(model, hp, 3, 0, eps, &F, G, &info);
  CHECK(rc == GPR_OK && info == 0, gpr_last_error(ctx));
  Fref = 0.5 * (yy / d + N * log(d) + N * log(2.0 * PI));
  CHECK(relerr(F, Fref) < 1e-13, "NLML known answer");
  /* dF/dsigma = -0.5 (alpha' dK alpha - tr(K^-1 dK)) with dK = (2/sigma)(sigma^2 + eps) I, alpha = y / d */
  g_sigma = -0.5 * (2.0 / hp[0]) * (hp[0] * hp[0] + eps) * (yy / (d * d) - N / d);
  g_noise = -0.5 * 2.0 * hp[2] * (yy / (d * d) - N / d);
  CHECK(relerr(G[0], g_sigma) < 1e-11, "gradient w.r.t. sigma");
  CHECK(relerr(G[2], g_noise) < 1e-11, "gradient w.r.t. sigma_n");
  CHECK(fabs(G[1]) < 1e-9 * fabs(g_sigma), "gradient w.r.t. the length scale of a diagonal K");
  CHECK(gpr_nlml_grad(model, hp, 2, 0, eps, &F, G, &info) == GPR_ERR_ARG, "Parameter size mismatch must be rejected");

  /* update_cache!(pc, md) + predict! with a Diagonal Sigma (src/predict.jl:29-34,51-62) */
  rc = gpr_update_cache(model, hp, 3, eps, 0, &info);
  CHECK(rc == GPR_OK, gpr_last_error(ctx));
  rc = gpr_predict(model, x, N, 0, mean, var, NULL);
  CHECK(rc == GPR_OK, gpr_last_error(ctx));
  for (i = 0; i < N; ++i) {
    CHECK(fabs(mean[i] - y[i] * hp[0] * hp[0] / d) < 1e-13, "predictive mean");
    CHECK(fabs(var[i] - (hp[0] * hp[0] + hp[2] * hp[2] - pow(hp[0], 4) / d)) < 1e-13, "predictive variance");
  }
  CHECK(gpr_timings(model, ms, GPR_T_COUNT) == GPR_OK && ms[GPR_T_POTRF] > 0.0, "gpr_timings");
  CHECK(gpr_ctx_launch_count(ctx) > 0, "launch count");

  /* not positive definite -> GPR_ERR_NOT_POSDEF with the LAPACK-style pivot index (src/cost.jl:77) */
  {
    double hp_bad[3] = {1.0, 0.0, 0.0};       /* all points identical after scaling by l = 0, no noise, no jitter */
    rc = gpr_update_cache(model, hp_bad, 3, 0.0, 0, &info);
    CHECK(rc == GPR_ERR_NOT_POSDEF && info >= 1 && info <= N, "PosDefException contract");
  }

  /* release in the "wrong" order (finalizers at exit): the context first, then its models -- must be harmless */
  CHECK(gpr_ctx_destroy(ctx) == GPR_OK, "gpr_ctx_destroy");
  CHECK(gpr_model_destroy(model) == GPR_OK, "gpr_model_destroy after its context");
  CHECK(gpr_ctx_destroy(ctx) == GPR_OK, "double gpr_ctx_destroy");
  printf("test_cabi ok: F = %.12f (closed form %.12f), G = [%.6e %.6e %.6e]\n", F, Fref, G[0], G[1], G[2]);
  return 0;
}
