/* Plain-C client of include/stz.h (no CUDA, no C++): proves the boundary is a C ABI.  Built and run by
 * tests/test_abi_cpu.py::test_plain_c_client_links_and_runs; needs no GPU (host-only entry points). */
#include <stdio.h>
#include <string.h>
#include "stz.h"

int main(void) {
  stz_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.n_style = 50; cfg.d_style = 512; cfg.d_model = 512; cfg.n_heads = 8; cfg.d_ff = 2048; cfg.n_layers = 8;
  cfg.d_text = 512; cfg.d_prompt = 512; cfg.d_time = 256; cfg.d_hid = 512; cfg.d_sty_tok = 128; cfg.n_sp_heads = 4;
  cfg.n_lstm = 4; cfg.max_dur = 50;
  cfg.sigma_data = 0.5f; cfg.sigma_max = 3.0f; cfg.sigma_min = 1e-4f; cfg.rho = 9.0f;
  if (stz_abi_version() != STZ_ABI_VERSION) { printf("abi mismatch\n"); return 1; }
  if (stz_weights_nfloats(&cfg) == 0 || stz_weight_offset(&cfg, "in.w") < 0) { printf("layout\n"); return 2; }
  double sigma[8], init[2];
  float coef[8 * 8];
  int e = stz_debug_plan(&cfg, 4, STZ_SAMPLER_TEACHER, 2.0f, sigma, coef, NULL, init);
  if (e != 8 || sigma[0] < 2.999 || sigma[0] > 3.001 || init[0] != sigma[0]) { printf("plan %d %f\n", e, sigma[0]); return 3; }
  if (stz_debug_plan(&cfg, 0, STZ_SAMPLER_STUDENT, 2.0f, NULL, NULL, NULL, NULL) != STZ_E_ARG) { printf("arg check\n"); return 4; }
  if (stz_sample_style(NULL, NULL, NULL, NULL, NULL, NULL, 1, 1, 1, 1, 1.0f, 0, NULL, NULL) != STZ_E_ARG) { printf("null handle\n"); return 5; }
  printf("ok %d evaluations, sigma_0 %.3f, %zu weight floats\n", e, sigma[0], stz_weights_nfloats(&cfg));
  return 0;
}
