/* Plain-C consumer of include/sindy_b200.h: proves that the header compiles as C (no C++ in the signatures), that
 * the shared library links with C linkage, and that argument validation works without a CUDA device.
 * Built and run by tests/test_abi_and_host.py::test_c_consumer_links_and_validates (gcc, no nvcc, no GPU). */
#include <stdio.h>
#include <string.h>

#include "sindy_b200.h"

#define CHECK(cond)                                                    \
  do {                                                                 \
    if (!(cond)) {                                                     \
      printf("FAILED line %d: %s (%s)\n", __LINE__, #cond, sb_last_error()); \
      return 1;                                                        \
    }                                                                  \
  } while (0)

int main(void) {
  sb_library lib = {3, 5, 0, 0};
  sb_library bad = {9, 2, 0, 0};
  int32_t expo[56 * 3];
  float dummy[4];
  double out[170];
  sb_fit_options opt;

  CHECK(sb_version() >= 100);
  CHECK(sb_library_size(&lib) == 56);
  CHECK(sb_library_size(&bad) == SB_ERR_UNSUPPORTED);
  CHECK(strstr(sb_last_error(), "dim=9") != NULL);
  CHECK(sb_library_exponents(&lib, expo) == SB_OK);
  CHECK(expo[4 * 3 + 0] == 2 && expo[4 * 3 + 1] == 0);            /* column 4 of d=3 is x*x */
  CHECK(expo[55 * 3 + 2] == 5);                                     /* last column is z^5 */
  CHECK(sb_workspace_bytes(&lib) > 0);
  CHECK(sb_train_step_out_len(&lib, SB_STEP_LOSS | SB_STEP_GRAD) == 2 + 3 * 56);
  CHECK(sb_peer_buffer_bytes(&lib, 8) == 2 * 8 * 170 * 16);
  CHECK(strcmp(sb_train_step_variant(&lib, SB_STEP_LOSS | SB_STEP_GRAD), "fused_tma<3,5>") == 0);
  CHECK(strcmp(sb_train_step_variant(&lib, SB_STEP_GRAM), "moments") == 0);
  /* validation happens before any CUDA call */
  CHECK(sb_forward(NULL, 4, &lib, dummy, dummy, NULL) == SB_ERR_INVALID);
  CHECK(sb_train_step(dummy, dummy, -1, &lib, dummy, 3u, out, dummy, 1024, NULL) == SB_ERR_INVALID);
  CHECK(sb_train_step(dummy, dummy, 4, &lib, dummy, 64u, out, dummy, 1024, NULL) == SB_ERR_INVALID);
  CHECK(sb_rollout(dummy, 4, &lib, dummy, 0.01, 10, 0, SB_RK4, SB_F32, 0, NULL, NULL, NULL, NULL) == SB_ERR_INVALID);
  memset(&opt, 0, sizeof(opt));
  opt.kind = 42;
  CHECK(sb_fit_step(dummy, dummy, 4, &lib, dummy, NULL, &opt, NULL, out, dummy, dummy, dummy, 1024, NULL, 1, 0, NULL,
                    0u, NULL) == SB_ERR_INVALID);
  printf("c abi ok: version %d, devices %d\n", sb_version(), sb_device_count());
  return 0;
}
