/* Latency of the C-ABI without any Python in the way: b200sv_check_batch with 1 / 32 / 1024 states on the C1 scene.
 *   gcc -O2 -std=gnu99 -Iinclude -o /tmp/c_latency tools/c_latency.c -Lmoveit2_b200 -lb200sv -Wl,-rpath,$PWD/moveit2_b200
 *   /tmp/c_latency moveit2_b200/resources/panda_spheres.urdf moveit2_b200/resources/panda.srdf */
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
  const int sizes[4] = { 1, 32, 1024, 4096 };
  static double q[4096 * 7], pts[3 * 2000], t[400];
  static uint8_t bits[4096 / 8];
  int i, k, rc;
  char *urdf = slurp(argv[1]), *srdf = slurp(argv[2]);
  if (argc < 3 || !urdf || !srdf)
    return 2;
  memset(&cfg, 0, sizeof(cfg));
  cfg.size[0] = cfg.size[1] = cfg.size[2] = 1.0;
  cfg.origin[2] = 0.5;
  cfg.resolution = 0.02;
  cfg.max_propagation_distance = 0.25;
  rc = b200sv_create_from_urdf(0, urdf, srdf, "panda_arm", &cfg, &ctx);
  if (rc != B200SV_OK)
  {
    fprintf(stderr, "create: %d %s\n", rc, b200sv_last_error(NULL));
    return 4;
  }
  srand(1);
  for (i = 0; i < 2000; ++i)
  {  /* a wall of points at x = 0.4 */
    pts[3 * i] = 0.4;
    pts[3 * i + 1] = rand() / (double)RAND_MAX - 0.5;
    pts[3 * i + 2] = rand() / (double)RAND_MAX;
  }
  b200sv_field_add_points(ctx, B200SV_FIELD_WORLD, pts, 2000);
  for (i = 0; i < 4096 * 7; ++i)
    q[i] = (rand() / (double)RAND_MAX - 0.5) * 3.0 - ((i % 7) == 3 ? 1.5 : 0.0);
  printf("{\"what\": \"b200sv_check_batch wall clock from C, pageable host buffers, median of 400 calls\"");
  for (k = 0; k < 4; ++k)
  {
    for (i = 0; i < 50; ++i)
      b200sv_check_batch(ctx, q, (size_t)sizes[k], B200SV_CHECK_ALL, bits);
    for (i = 0; i < 400; ++i)
    {
      const double t0 = now_us();
      if (b200sv_check_batch(ctx, q, (size_t)sizes[k], B200SV_CHECK_ALL, bits) != B200SV_OK)
        return 5;
      t[i] = now_us() - t0;
    }
    qsort(t, 400, sizeof(double), cmp);
    printf(", \"n=%d\": {\"median_us\": %.1f, \"p90_us\": %.1f}", sizes[k], t[200], t[360]);
  }
  {
    /* single states the way a planner issues them: a different state every call, back to back; then the same with a pause
     * that lets the resident kernel expire before every call; then the answers against ONE batch call over the same states */
    static double ts[2000];
    static uint8_t one[8], other[8], all[4096 / 8];
    int bad = 0;
    struct timespec pause = { 0, 1000000 };  /* 1 ms */
    for (i = 0; i < 2000; ++i)
    {
      const double t0 = now_us();
      if (b200sv_check_batch(ctx, q + 7 * (i % 4096), 1, B200SV_CHECK_ALL, one) != B200SV_OK)
        return 6;
      ts[i] = now_us() - t0;
    }
    qsort(ts, 2000, sizeof(double), cmp);
    printf(", \"n=1, a new state every call, back to back\": {\"median_us\": %.1f, \"p90_us\": %.1f, \"p99_us\": %.1f}", ts[1000],
           ts[1800], ts[1980]);
    for (i = 0; i < 200; ++i)
    {
      double t0;
      nanosleep(&pause, NULL);
      t0 = now_us();
      if (b200sv_check_batch(ctx, q + 7 * (i % 4096), 1, B200SV_CHECK_ALL, one) != B200SV_OK)
        return 6;
      ts[i] = now_us() - t0;
    }
    qsort(ts, 200, sizeof(double), cmp);
    printf(", \"n=1, 1 ms idle before every call\": {\"median_us\": %.1f, \"p90_us\": %.1f}", ts[100], ts[180]);
    if (b200sv_check_batch(ctx, q, 4096, B200SV_CHECK_ALL, all) != B200SV_OK)
      return 7;
    for (i = 0; i < 4096; ++i)
    {
      if (b200sv_check_batch(ctx, q + 7 * i, 1, B200SV_CHECK_ALL, one) != B200SV_OK)
        return 8;
      bad += (one[0] & 1) != ((all[i >> 3] >> (i & 7)) & 1);
      if ((i & 255) == 255)  /* other entry points in between: the resident kernel must step aside and come back */
        b200sv_check_batch(ctx, q, 64, B200SV_CHECK_ALL, other);
    }
    printf(", \"single_state_answers_differing_from_the_batch\": %d", bad);
  }
  printf("}\n");
  b200sv_destroy(ctx);
  return 0;
}
