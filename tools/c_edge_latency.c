/* Latency of b200sv_check_motions from C: 1 / 8 / 64 edges per call on a C1-sized scene (what an OMPL MotionValidator pays per
 * checkMotion).  Build like tools/c_latency.c. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

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
static double now_us(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec * 1e6 + t.tv_nsec * 1e-3;
}
static int cmp(const void* a, const void* b) { return (*(const double*)a > *(const double*)b) - (*(const double*)a < *(const double*)b); }

int main(int argc, char** argv)
{
  b200sv_field_cfg cfg;
  b200sv_ctx* ctx = NULL;
  const int sizes[3] = { 1, 8, 64 };
  static double from[64 * 7], to[64 * 7], pts[3 * 2000], t[400], factor[7] = { 1, 1, 1, 1, 1, 1, 1 };
  static uint8_t bits[8], cont[7];
  static int32_t first[64];
  uint64_t checked = 0, total = 0;
  int i, k;
  char *urdf = slurp(argv[1]), *srdf = slurp(argv[2]);
  if (argc < 3 || !urdf || !srdf)
    return 2;
  memset(&cfg, 0, sizeof(cfg));
  cfg.size[0] = cfg.size[1] = cfg.size[2] = 1.0;
  cfg.origin[2] = 0.5;
  cfg.resolution = 0.02;
  cfg.max_propagation_distance = 0.25;
  if (b200sv_create_from_urdf(0, urdf, srdf, "panda_arm", &cfg, &ctx) != B200SV_OK)
    return 4;
  srand(1);
  for (i = 0; i < 2000; ++i)
  {
    pts[3 * i] = 0.4;
    pts[3 * i + 1] = rand() / (double)RAND_MAX - 0.5;
    pts[3 * i + 2] = rand() / (double)RAND_MAX;
  }
  b200sv_field_add_points(ctx, B200SV_FIELD_WORLD, pts, 2000);
  for (i = 0; i < 64 * 7; ++i)
  {
    from[i] = (rand() / (double)RAND_MAX - 0.5) * 3.0 - ((i % 7) == 3 ? 1.5 : 0.0);
    to[i] = from[i] + (rand() / (double)RAND_MAX - 0.5) * 0.6;  /* RRT-sized steps */
  }
  printf("{\"what\": \"b200sv_check_motions wall clock from C, median of 400 calls, max_segment_length 0.05\"");
  for (k = 0; k < 3; ++k)
  {
    for (i = 0; i < 50; ++i)
      b200sv_check_motions(ctx, from, to, (size_t)sizes[k], 0.05, cont, factor, 0, B200SV_CHECK_ALL, bits, first, &checked);
    total = 0;
    for (i = 0; i < 400; ++i)
    {
      const double t0 = now_us();
      if (b200sv_check_motions(ctx, from, to, (size_t)sizes[k], 0.05, cont, factor, 0, B200SV_CHECK_ALL, bits, first, &checked) !=
          B200SV_OK)
        return 5;
      t[i] = now_us() - t0;
      total += checked;
    }
    qsort(t, 400, sizeof(double), cmp);
    printf(", \"edges=%d\": {\"median_us\": %.1f, \"p90_us\": %.1f, \"states_per_call\": %.1f}", sizes[k], t[200], t[360],
           total / 400.0);
  }
  printf("}\n");
  b200sv_destroy(ctx);
  return 0;
}
