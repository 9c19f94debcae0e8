/* Compiled as C99 by tests/test_host_model.py::test_header_is_plain_c_and_links: include/b200sv.h must be usable from C (the
 * boundary a cgo / JNI / ctypes / pluginlib binding sees), and a context must be creatable and inspectable through it alone.
 * No CUDA device needed: B200SV_DEVICE_NONE builds the model and the lookup tables on the host and refuses to compute. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200sv.h"

static char* slurp(const char* path)
{
  FILE* f = fopen(path, "rb");
  long n;
  char* s;
  if (!f)
    return NULL;
  fseek(f, 0, SEEK_END);
  n = ftell(f);
  fseek(f, 0, SEEK_SET);
  s = (char*)malloc((size_t)n + 1);
  if (fread(s, 1, (size_t)n, f) != (size_t)n)
    n = 0;
  s[n] = 0;
  fclose(f);
  return s;
}

int main(int argc, char** argv)
{
  b200sv_field_cfg cfg;
  b200sv_ctx* ctx = NULL;
  b200sv_info info;
  uint8_t bit = 0;
  double q[7] = { 0, 0, 0, -1.5, 0, 1.5, 0 };
  char *urdf, *srdf;
  int rc;
  if (argc < 3)
    return 2;
  urdf = slurp(argv[1]);
  srdf = slurp(argv[2]);
  if (!urdf || !srdf)
    return 3;
  memset(&cfg, 0, sizeof(cfg));
  cfg.size[0] = cfg.size[1] = cfg.size[2] = 1.0;
  cfg.origin[2] = 0.5;
  cfg.resolution = 0.02;
  cfg.max_propagation_distance = 0.25;
  rc = b200sv_create_from_urdf(B200SV_DEVICE_NONE, urdf, srdf, "panda_arm", &cfg, &ctx);
  if (rc != B200SV_OK)
  {
    fprintf(stderr, "create: %d %s\n", rc, b200sv_last_error(NULL));
    return 4;
  }
  if (b200sv_get_info(ctx, &info) != B200SV_OK || info.n_group_vars != 7 || info.n_spheres != 58 || info.num_cells[0] != 50)
    return 5;
  /* a host-only context refuses every compute call: there is no CPU fallback behind the ABI */
  rc = b200sv_check_batch(ctx, q, 1, B200SV_CHECK_ALL, &bit);
  if (rc != B200SV_ERR_NO_DEVICE)
    return 6;
  /* ... and so does the switch to the reference's own propagation (it needs the device) */
  if (b200sv_field_set_propagation(ctx, B200SV_FIELD_WORLD, B200SV_PROPAGATION_REFERENCE) != B200SV_ERR_NO_DEVICE)
    return 7;
  rc = b200sv_check_batch(ctx, q, 1, B200SV_CHECK_ALL, &bit); /* the message printed below is this call's */
  printf("c abi ok: %d links, %d group variables, %d spheres, %d link pairs, field %d x %d x %d, max d2 %d; compute refused: %s\n",
         info.n_links, info.n_group_vars, info.n_spheres, info.n_link_pairs, info.num_cells[0], info.num_cells[1],
         info.num_cells[2], info.max_distance_sq, b200sv_last_error(ctx));
  b200sv_destroy(ctx);
  free(urdf);
  free(srdf);
  return 0;
}
