/* Plain C99 against include/b200sv.h: drives TWO devices (or one device twice when the box has a single GPU) through the
 * multi-device handle -- create, build the world field once and replicate it, check one sharded batch, compare with a
 * single-context check of the same states.  Built and run by tests/test_gpu_round2.py::test_c_program_drives_two_devices.
 *   usage: c_multi_check <urdf> <srdf> <n_devices_visible> */
#include <math.h>
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

static unsigned long long rng_state = 88172645463325252ull;
static double uniform01(void)
{
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return (double)(rng_state >> 11) / 9007199254740992.0;
}

int main(int argc, char** argv)
{
  enum { N = 200000 + 13, NP = 3000 };
  b200sv_field_cfg cfg;
  b200sv_multi* m = NULL;
  b200sv_ctx* one = NULL;
  b200sv_stats st;
  double lo[7], hi[7];
  double *q, *pts;
  uint8_t *bm_multi, *bm_one;
  int devices[2] = { 0, 1 };
  int rc, i, k;
  char *urdf, *srdf;
  if (argc < 4)
    return 2;
  urdf = slurp(argv[1]);
  srdf = slurp(argv[2]);
  if (!urdf || !srdf)
    return 3;
  if (atoi(argv[3]) < 2)
    devices[1] = 0;
  memset(&cfg, 0, sizeof(cfg));
  cfg.size[0] = cfg.size[1] = cfg.size[2] = 1.0;
  cfg.origin[2] = 0.5;
  cfg.resolution = 0.02;
  cfg.max_propagation_distance = 0.25;
  rc = b200sv_create_multi_from_urdf(devices, 2, urdf, srdf, "panda_arm", &cfg, &m);
  if (rc != B200SV_OK)
  {
    fprintf(stderr, "create_multi: %d %s\n", rc, b200sv_multi_last_error(NULL));
    return 4;
  }
  rc = b200sv_create_from_urdf(0, urdf, srdf, "panda_arm", &cfg, &one);
  if (rc != B200SV_OK)
    return 5;
  if (b200sv_multi_device_count(m) != 2 || !b200sv_multi_ctx(m, 1) || b200sv_multi_ctx(m, 2))
    return 6;
  pts = (double*)malloc(sizeof(double) * 3 * NP);
  for (i = 0; i < NP; ++i)
  { /* a shell of points around the arm */
    const double a = 6.283185307179586 * uniform01(), z = 0.1 + 0.8 * uniform01(), r = 0.3 + 0.1 * uniform01();
    pts[3 * i] = r * cos(a);
    pts[3 * i + 1] = r * sin(a);
    pts[3 * i + 2] = z;
  }
  if (b200sv_multi_field_add_points(m, B200SV_FIELD_WORLD, pts, NP) != B200SV_OK)
  {
    fprintf(stderr, "multi add_points: %s\n", b200sv_multi_last_error(m));
    return 7;
  }
  if (b200sv_field_add_points(one, B200SV_FIELD_WORLD, pts, NP) != B200SV_OK)
    return 8;
  if (b200sv_get_group_bounds(one, lo, hi) != B200SV_OK)
    return 9;
  q = (double*)malloc(sizeof(double) * 7 * N);
  for (i = 0; i < N; ++i)
    for (k = 0; k < 7; ++k)
      q[7 * i + k] = lo[k] + (hi[k] - lo[k]) * uniform01();
  bm_multi = (uint8_t*)calloc((N + 7) / 8 + 8, 1);
  bm_one = (uint8_t*)calloc((N + 7) / 8 + 8, 1);
  if (b200sv_check_batch_multi(m, q, N, B200SV_CHECK_ALL, bm_multi) != B200SV_OK)
  {
    fprintf(stderr, "check_batch_multi: %s\n", b200sv_multi_last_error(m));
    return 10;
  }
  if (b200sv_check_batch(one, q, N, B200SV_CHECK_ALL, bm_one) != B200SV_OK)
    return 11;
  if (memcmp(bm_multi, bm_one, (N + 7) / 8) != 0)
  {
    fprintf(stderr, "sharded bitmap differs from the single-context bitmap\n");
    return 12;
  }
  if (b200sv_multi_get_stats(m, &st) != B200SV_OK || st.n_states != (uint64_t)N || st.n_valid == 0 || st.n_valid == (uint64_t)N)
    return 13;
  printf("c multi ok: devices {%d, %d}, %d states, %llu valid, %llu decided in FP64, slowest device %.3f ms\n", devices[0],
         devices[1], (int)N, (unsigned long long)st.n_valid, (unsigned long long)st.n_rechecked, st.check_ms);
  b200sv_destroy_multi(m);
  b200sv_destroy(one);
  free(q);
  free(pts);
  free(bm_multi);
  free(bm_one);
  free(urdf);
  free(srdf);
  return 0;
}
