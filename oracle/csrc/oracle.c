/*
 * ORACLE -- test infrastructure, NOT product code.
 *
 * Plain-C CPU restatement of MoveIt 2's distance-field state-validity path.  It exists to (a) check the CUDA
 * path's results and (b) serve as the timed CPU baseline in bench.py.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; moveit2_b200/ never does.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference/moveit_core).
 * The reference itself cannot be compiled in this image (Eigen, rclcpp, geometric_shapes, octomap, urdfdom
 * are absent -- SURVEY.md section 8c), so parity is pinned on the reference's own golden data instead
 * (tests/test_oracle_*.py): test_small.df / test_big.df occupancy, TestAddRemovePoints properties,
 * VoxelGrid sizing, OneRobot FK poses.
 *
 * Floating point: compiled with -ffp-contract=off so every product/sum rounds once, in the order written.
 * Eigen's fixed-size products (affine 3x4 * 4x4) accumulate k = 0..3 in order; that order is used here.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------------
 * VoxelGrid geometry: distance_field/include/moveit/distance_field/voxel_grid.hpp:364-396 (resize),
 * :405-408 (isCellValid), :423-426 (ref), :506-523 (getCellFromLocation), :526-529 (getLocationFromCell)
 * ---------------------------------------------------------------------------------------------------- */
typedef struct
{
  double size[3], origin[3], origin_minus[3], resolution, oo_resolution;
  int num_cells[3];
  int stride1, stride2;
  long num_cells_total;
} orc_grid;

void orc_grid_init(orc_grid* g, double sx, double sy, double sz, double res, double ox, double oy, double oz)
{
  g->size[0] = sx; g->size[1] = sy; g->size[2] = sz;
  g->origin[0] = ox; g->origin[1] = oy; g->origin[2] = oz;
  g->origin_minus[0] = ox - 0.5 * res;
  g->origin_minus[1] = oy - 0.5 * res;
  g->origin_minus[2] = oz - 0.5 * res;
  g->num_cells_total = 1;
  g->resolution = res;
  g->oo_resolution = 1.0 / res;
  for (int i = 0; i < 3; ++i)
  {
    g->num_cells[i] = (int)(g->size[i] * g->oo_resolution); /* voxel_grid.hpp:384 double -> int truncation */
    g->num_cells_total *= g->num_cells[i];
  }
  g->stride1 = g->num_cells[1] * g->num_cells[2];
  g->stride2 = g->num_cells[2];
}

static inline int grid_cell(const orc_grid* g, int dim, double loc)
{
  return (int)floor((loc - g->origin_minus[dim]) * g->oo_resolution); /* voxel_grid.hpp:522 */
}
static inline int grid_valid(const orc_grid* g, int x, int y, int z)
{
  return x >= 0 && x < g->num_cells[0] && y >= 0 && y < g->num_cells[1] && z >= 0 && z < g->num_cells[2];
}
static inline long grid_ref(const orc_grid* g, int x, int y, int z)
{
  return (long)x * g->stride1 + (long)y * g->stride2 + z;
}
int orc_grid_world_to_grid(const orc_grid* g, double wx, double wy, double wz, int* xyz)
{
  xyz[0] = grid_cell(g, 0, wx);
  xyz[1] = grid_cell(g, 1, wy);
  xyz[2] = grid_cell(g, 2, wz);
  return grid_valid(g, xyz[0], xyz[1], xyz[2]); /* voxel_grid.hpp:554-560 */
}
void orc_grid_grid_to_world(const orc_grid* g, int x, int y, int z, double* w)
{
  w[0] = g->origin[0] + g->resolution * (double)x; /* voxel_grid.hpp:526-529 */
  w[1] = g->origin[1] + g->resolution * (double)y;
  w[2] = g->origin[2] + g->resolution * (double)z;
}

/* ------------------------------------------------------------------------------------------------------
 * PropagationDistanceField (unsigned, and signed = propagate_negative_): distance_field/src/propagation_distance_field.cpp
 * The voxel record keeps the reference's 40-byte layout (propagation_distance_field.hpp:82-116).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct
{
  int distance_square;
  int negative_distance_square;
  int closest_point[3];
  int closest_negative_point[3];
  int update_direction;
  int negative_update_direction;
} orc_voxel;

typedef struct { int x, y, z; } ivec3;
typedef struct { ivec3* v; size_t n, cap; } ivec3_list;

static void list_push(ivec3_list* l, ivec3 p)
{
  if (l->n == l->cap)
  {
    l->cap = l->cap ? l->cap * 2 : 64;
    l->v = (ivec3*)realloc(l->v, l->cap * sizeof(ivec3));
  }
  l->v[l->n++] = p;
}

typedef struct
{
  orc_grid grid;
  orc_voxel* data;
  double max_distance;
  int max_distance_sq;
  double* sqrt_table;
  ivec3_list* bucket_queue; /* max_distance_sq + 1 buckets */
  ivec3_list* negative_bucket_queue;
  int propagate_negative; /* the constructor's propagate_negative_distances (:48-58) */
  ivec3 neighborhoods[2][27][26];
  int neighborhood_size[2][27];
  ivec3 direction_number_to_direction[27];
  int inv_twice_resolution; /* distance_field.hpp:614 declares this as int (reference quirk) */
} orc_pdf;

static inline int direction_number(int dx, int dy, int dz) { return (dx + 1) * 9 + (dy + 1) * 3 + dz + 1; } /* :584-587 */

static void pdf_init_neighborhoods(orc_pdf* f) /* :526-582 */
{
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dz = -1; dz <= 1; ++dz)
      {
        ivec3 p = { dx, dy, dz };
        f->direction_number_to_direction[direction_number(dx, dy, dz)] = p;
      }
  for (int n = 0; n < 2; ++n)
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz)
        {
          int dn = direction_number(dx, dy, dz);
          int cnt = 0;
          for (int tdx = -1; tdx <= 1; ++tdx)
            for (int tdy = -1; tdy <= 1; ++tdy)
              for (int tdz = -1; tdz <= 1; ++tdz)
              {
                if (tdx == 0 && tdy == 0 && tdz == 0)
                  continue;
                if (n >= 1)
                {
                  if ((abs(tdx) + abs(tdy) + abs(tdz)) != 1)
                    continue;
                  if (dx * tdx < 0 || dy * tdy < 0 || dz * tdz < 0)
                    continue;
                }
                ivec3 p = { tdx, tdy, tdz };
                f->neighborhoods[n][dn][cnt++] = p;
              }
          f->neighborhood_size[n][dn] = cnt;
        }
}

void orc_pdf_reset(orc_pdf* f) /* :506-524 and the PropDistanceFieldVoxel(int,int) ctor, hpp:589-598 */
{
  const orc_grid* g = &f->grid;
  for (int x = 0; x < g->num_cells[0]; ++x)
    for (int y = 0; y < g->num_cells[1]; ++y)
      for (int z = 0; z < g->num_cells[2]; ++z)
      {
        orc_voxel* v = &f->data[grid_ref(g, x, y, z)];
        v->distance_square = f->max_distance_sq;
        v->negative_distance_square = 0;
        v->closest_point[0] = v->closest_point[1] = v->closest_point[2] = -1; /* UNINITIALIZED */
        v->closest_negative_point[0] = x;
        v->closest_negative_point[1] = y;
        v->closest_negative_point[2] = z;
        v->update_direction = 0;
        v->negative_update_direction = 0;
      }
}

orc_pdf* orc_pdf_create_signed(double sx, double sy, double sz, double res, double ox, double oy, double oz,
                               double max_distance, int propagate_negative)
{
  orc_pdf* f = (orc_pdf*)calloc(1, sizeof(orc_pdf));
  f->propagate_negative = propagate_negative;
  orc_grid_init(&f->grid, sx, sy, sz, res, ox, oy, oz);
  f->max_distance = max_distance;
  f->inv_twice_resolution = (int)(1.0 / (2.0 * res)); /* distance_field.cpp:67 into an int member */
  /* initialize(): :86-104 */
  f->max_distance_sq = (int)(ceil(max_distance / res) * ceil(max_distance / res));
  f->data = (orc_voxel*)malloc((size_t)f->grid.num_cells_total * sizeof(orc_voxel));
  pdf_init_neighborhoods(f);
  f->bucket_queue = (ivec3_list*)calloc((size_t)f->max_distance_sq + 1, sizeof(ivec3_list));
  f->negative_bucket_queue = (ivec3_list*)calloc((size_t)f->max_distance_sq + 1, sizeof(ivec3_list));
  f->sqrt_table = (double*)malloc(((size_t)f->max_distance_sq + 1) * sizeof(double));
  for (int i = 0; i <= f->max_distance_sq; ++i)
    f->sqrt_table[i] = sqrt((double)i) * res;
  orc_pdf_reset(f);
  return f;
}

orc_pdf* orc_pdf_create(double sx, double sy, double sz, double res, double ox, double oy, double oz, double max_distance)
{
  return orc_pdf_create_signed(sx, sy, sz, res, ox, oy, oz, max_distance, 0);
}

void orc_pdf_destroy(orc_pdf* f)
{
  if (!f)
    return;
  for (int i = 0; i <= f->max_distance_sq; ++i)
  {
    free(f->bucket_queue[i].v);
    free(f->negative_bucket_queue[i].v);
  }
  free(f->bucket_queue);
  free(f->negative_bucket_queue);
  free(f->sqrt_table);
  free(f->data);
  free(f);
}

static void pdf_propagate_positive(orc_pdf* f) /* :390-444 */
{
  const orc_grid* g = &f->grid;
  for (int i = 0; i <= f->max_distance_sq; ++i)
  {
    ivec3_list* bucket = &f->bucket_queue[i];
    size_t end = bucket->n; /* list_end is captured before the loop in the reference */
    for (size_t it = 0; it < end; ++it)
    {
      ivec3 loc = bucket->v[it];
      orc_voxel* vptr = &f->data[grid_ref(g, loc.x, loc.y, loc.z)];
      int d = i;
      if (d > 1)
        d = 1;
      if (vptr->update_direction < 0 || vptr->update_direction > 26)
        continue;
      const ivec3* nb = f->neighborhoods[d][vptr->update_direction];
      int nn = f->neighborhood_size[d][vptr->update_direction];
      for (int k = 0; k < nn; ++k)
      {
        ivec3 diff = nb[k];
        ivec3 nloc = { loc.x + diff.x, loc.y + diff.y, loc.z + diff.z };
        if (!grid_valid(g, nloc.x, nloc.y, nloc.z))
          continue;
        orc_voxel* neighbor = &f->data[grid_ref(g, nloc.x, nloc.y, nloc.z)];
        int ex = vptr->closest_point[0] - nloc.x, ey = vptr->closest_point[1] - nloc.y,
            ez = vptr->closest_point[2] - nloc.z;
        int new_distance_sq = ex * ex + ey * ey + ez * ez;
        if (new_distance_sq > f->max_distance_sq)
          continue;
        if (new_distance_sq < neighbor->distance_square)
        {
          neighbor->distance_square = new_distance_sq;
          neighbor->closest_point[0] = vptr->closest_point[0];
          neighbor->closest_point[1] = vptr->closest_point[1];
          neighbor->closest_point[2] = vptr->closest_point[2];
          neighbor->update_direction = direction_number(diff.x, diff.y, diff.z);
          list_push(&f->bucket_queue[new_distance_sq], nloc);
          if (new_distance_sq == i)
          { /* cannot happen on an integer lattice (a unit step changes d^2 by an odd amount); keep the
               reference's iterator semantics if it ever did */
            bucket = &f->bucket_queue[i];
          }
        }
      }
    }
    f->bucket_queue[i].n = 0; /* clear() */
  }
}

static void pdf_propagate_negative(orc_pdf* f) /* :446-504 */
{
  const orc_grid* g = &f->grid;
  for (int i = 0; i <= f->max_distance_sq; ++i)
  {
    size_t end = f->negative_bucket_queue[i].n; /* list_end is captured before the loop in the reference */
    for (size_t it = 0; it < end; ++it)
    {
      ivec3 loc = f->negative_bucket_queue[i].v[it];
      orc_voxel* vptr = &f->data[grid_ref(g, loc.x, loc.y, loc.z)];
      int d = i;
      if (d > 1)
        d = 1;
      if (vptr->negative_update_direction < 0 || vptr->negative_update_direction > 26)
        continue;
      const ivec3* nb = f->neighborhoods[d][vptr->negative_update_direction];
      int nn = f->neighborhood_size[d][vptr->negative_update_direction];
      for (int k = 0; k < nn; ++k)
      {
        ivec3 diff = nb[k];
        ivec3 nloc = { loc.x + diff.x, loc.y + diff.y, loc.z + diff.z };
        if (!grid_valid(g, nloc.x, nloc.y, nloc.z))
          continue;
        orc_voxel* neighbor = &f->data[grid_ref(g, nloc.x, nloc.y, nloc.z)];
        int ex = vptr->closest_negative_point[0] - nloc.x, ey = vptr->closest_negative_point[1] - nloc.y,
            ez = vptr->closest_negative_point[2] - nloc.z;
        int new_distance_sq = ex * ex + ey * ey + ez * ez;
        if (new_distance_sq > f->max_distance_sq)
          continue;
        if (new_distance_sq < neighbor->negative_distance_square)
        {
          neighbor->negative_distance_square = new_distance_sq;
          neighbor->closest_negative_point[0] = vptr->closest_negative_point[0];
          neighbor->closest_negative_point[1] = vptr->closest_negative_point[1];
          neighbor->closest_negative_point[2] = vptr->closest_negative_point[2];
          neighbor->negative_update_direction = direction_number(diff.x, diff.y, diff.z);
          list_push(&f->negative_bucket_queue[new_distance_sq], nloc);
        }
      }
    }
    f->negative_bucket_queue[i].n = 0; /* clear() */
  }
}

static void pdf_add_new_obstacle_voxels(orc_pdf* f, const ivec3* pts, size_t n) /* :221-301 */
{
  const orc_grid* g = &f->grid;
  int initial_update_direction = direction_number(0, 0, 0);
  ivec3_list negative_stack = { 0, 0, 0 };
  for (size_t i = 0; i < n; ++i)
  {
    orc_voxel* v = &f->data[grid_ref(g, pts[i].x, pts[i].y, pts[i].z)];
    v->distance_square = 0;
    v->closest_point[0] = pts[i].x;
    v->closest_point[1] = pts[i].y;
    v->closest_point[2] = pts[i].z;
    v->update_direction = initial_update_direction;
    list_push(&f->bucket_queue[0], pts[i]);
    if (f->propagate_negative)
    {
      v->negative_distance_square = f->max_distance_sq;
      v->closest_negative_point[0] = v->closest_negative_point[1] = v->closest_negative_point[2] = -1; /* UNINITIALIZED */
      list_push(&negative_stack, pts[i]);
    }
  }
  pdf_propagate_positive(f);
  if (f->propagate_negative)
  {
    while (negative_stack.n) /* :255-298 */
    {
      ivec3 loc = negative_stack.v[--negative_stack.n];
      for (int neighbor = 0; neighbor < 27; ++neighbor)
      {
        ivec3 diff = f->direction_number_to_direction[neighbor];
        ivec3 nloc = { loc.x + diff.x, loc.y + diff.y, loc.z + diff.z };
        if (!grid_valid(g, nloc.x, nloc.y, nloc.z))
          continue;
        orc_voxel* nvoxel = &f->data[grid_ref(g, nloc.x, nloc.y, nloc.z)];
        int* close_point = nvoxel->closest_negative_point;
        if (!grid_valid(g, close_point[0], close_point[1], close_point[2]))
        {
          close_point[0] = nloc.x;
          close_point[1] = nloc.y;
          close_point[2] = nloc.z;
        }
        orc_voxel* closest_point_voxel = &f->data[grid_ref(g, close_point[0], close_point[1], close_point[2])];
        if (closest_point_voxel->negative_distance_square != 0)
        { /* our closest non-obstacle cell has become an obstacle */
          if (nvoxel->negative_distance_square != f->max_distance_sq)
          {
            nvoxel->negative_distance_square = f->max_distance_sq;
            nvoxel->closest_negative_point[0] = nvoxel->closest_negative_point[1] = nvoxel->closest_negative_point[2] = -1;
            list_push(&negative_stack, nloc);
          }
        }
        else
        { /* this cell still has a valid non-obstacle cell: propagate from it */
          nvoxel->negative_update_direction = initial_update_direction;
          list_push(&f->negative_bucket_queue[0], nloc);
        }
      }
    }
    pdf_propagate_negative(f);
  }
  free(negative_stack.v);
}

static void pdf_remove_obstacle_voxels(orc_pdf* f, const ivec3* pts, size_t n) /* :303-388 */
{
  const orc_grid* g = &f->grid;
  ivec3_list stack = { 0, 0, 0 };
  int initial_update_direction = direction_number(0, 0, 0);
  for (size_t i = 0; i < n; ++i)
  {
    orc_voxel* v = &f->data[grid_ref(g, pts[i].x, pts[i].y, pts[i].z)];
    v->distance_square = f->max_distance_sq;
    v->closest_point[0] = pts[i].x;
    v->closest_point[1] = pts[i].y;
    v->closest_point[2] = pts[i].z;
    v->update_direction = initial_update_direction;
    list_push(&stack, pts[i]);
    if (f->propagate_negative)
    {
      v->negative_distance_square = 0;
      v->closest_negative_point[0] = pts[i].x;
      v->closest_negative_point[1] = pts[i].y;
      v->closest_negative_point[2] = pts[i].z;
      v->negative_update_direction = initial_update_direction;
      list_push(&f->negative_bucket_queue[0], pts[i]);
    }
  }
  while (stack.n)
  {
    ivec3 loc = stack.v[--stack.n];
    for (int neighbor = 0; neighbor < 27; ++neighbor)
    {
      ivec3 diff = f->direction_number_to_direction[neighbor];
      ivec3 nloc = { loc.x + diff.x, loc.y + diff.y, loc.z + diff.z };
      if (!grid_valid(g, nloc.x, nloc.y, nloc.z))
        continue;
      orc_voxel* nvoxel = &f->data[grid_ref(g, nloc.x, nloc.y, nloc.z)];
      int* close_point = nvoxel->closest_point;
      if (!grid_valid(g, close_point[0], close_point[1], close_point[2]))
      {
        close_point[0] = nloc.x;
        close_point[1] = nloc.y;
        close_point[2] = nloc.z;
      }
      orc_voxel* closest_point_voxel = &f->data[grid_ref(g, close_point[0], close_point[1], close_point[2])];
      if (closest_point_voxel->distance_square != 0)
      {
        if (nvoxel->distance_square != f->max_distance_sq)
        {
          nvoxel->distance_square = f->max_distance_sq;
          nvoxel->closest_point[0] = nloc.x;
          nvoxel->closest_point[1] = nloc.y;
          nvoxel->closest_point[2] = nloc.z;
          nvoxel->update_direction = initial_update_direction;
          list_push(&stack, nloc);
        }
      }
      else
      {
        nvoxel->update_direction = initial_update_direction;
        list_push(&f->bucket_queue[0], nloc);
      }
    }
  }
  free(stack.v);
  pdf_propagate_positive(f);
  if (f->propagate_negative)
    pdf_propagate_negative(f);
}

void orc_pdf_add_points(orc_pdf* f, const double* xyz, size_t n) /* addPointsToField :177-196 */
{
  ivec3* vox = (ivec3*)malloc((n ? n : 1) * sizeof(ivec3));
  size_t m = 0;
  for (size_t i = 0; i < n; ++i)
  {
    int c[3];
    if (orc_grid_world_to_grid(&f->grid, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], c))
      if (f->data[grid_ref(&f->grid, c[0], c[1], c[2])].distance_square > 0)
      {
        ivec3 p = { c[0], c[1], c[2] };
        vox[m++] = p;
      }
  }
  pdf_add_new_obstacle_voxels(f, vox, m);
  free(vox);
}

void orc_pdf_add_voxels(orc_pdf* f, const int32_t* ijk, size_t n) /* readFromStream tail :732-758 */
{
  ivec3* vox = (ivec3*)malloc((n ? n : 1) * sizeof(ivec3));
  for (size_t i = 0; i < n; ++i)
  {
    ivec3 p = { ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2] };
    vox[i] = p;
  }
  pdf_add_new_obstacle_voxels(f, vox, n);
  free(vox);
}

void orc_pdf_remove_points(orc_pdf* f, const double* xyz, size_t n) /* removePointsFromField :198-219 */
{
  ivec3* vox = (ivec3*)malloc((n ? n : 1) * sizeof(ivec3));
  size_t m = 0;
  for (size_t i = 0; i < n; ++i)
  {
    int c[3];
    if (orc_grid_world_to_grid(&f->grid, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], c))
    {
      ivec3 p = { c[0], c[1], c[2] };
      vox[m++] = p;
    }
  }
  pdf_remove_obstacle_voxels(f, vox, m);
  free(vox);
}

/* CompareEigenVector3i orders by z, then y, then x (propagation_distance_field.hpp:60-75) */
static int cmp_zyx(const void* a, const void* b)
{
  const ivec3 *p = (const ivec3*)a, *q = (const ivec3*)b;
  if (p->z != q->z) return p->z < q->z ? -1 : 1;
  if (p->y != q->y) return p->y < q->y ? -1 : 1;
  if (p->x != q->x) return p->x < q->x ? -1 : 1;
  return 0;
}
static size_t to_sorted_set(const orc_grid* g, const double* xyz, size_t n, ivec3* out)
{
  size_t m = 0;
  for (size_t i = 0; i < n; ++i)
  {
    int c[3];
    if (orc_grid_world_to_grid(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], c))
    {
      ivec3 p = { c[0], c[1], c[2] };
      out[m++] = p;
    }
  }
  qsort(out, m, sizeof(ivec3), cmp_zyx);
  size_t u = 0;
  for (size_t i = 0; i < m; ++i)
    if (u == 0 || cmp_zyx(&out[u - 1], &out[i]) != 0)
      out[u++] = out[i];
  return u;
}
static size_t set_difference(const ivec3* a, size_t na, const ivec3* b, size_t nb, ivec3* out)
{
  size_t i = 0, j = 0, m = 0;
  while (i < na)
  {
    if (j == nb) { out[m++] = a[i++]; continue; }
    int c = cmp_zyx(&a[i], &b[j]);
    if (c < 0) out[m++] = a[i++];
    else if (c > 0) ++j;
    else { ++i; ++j; }
  }
  return m;
}

void orc_pdf_update_points(orc_pdf* f, const double* old_xyz, size_t n_old, const double* new_xyz, size_t n_new)
{ /* updatePointsInField :130-175 */
  ivec3* so = (ivec3*)malloc((n_old + 1) * sizeof(ivec3));
  ivec3* sn = (ivec3*)malloc((n_new + 1) * sizeof(ivec3));
  size_t uo = to_sorted_set(&f->grid, old_xyz, n_old, so);
  size_t un = to_sorted_set(&f->grid, new_xyz, n_new, sn);
  ivec3* old_not_new = (ivec3*)malloc((uo + 1) * sizeof(ivec3));
  ivec3* new_not_old = (ivec3*)malloc((un + 1) * sizeof(ivec3));
  size_t a = set_difference(so, uo, sn, un, old_not_new);
  size_t b = set_difference(sn, un, so, uo, new_not_old);
  size_t c = 0;
  for (size_t i = 0; i < b; ++i)
    if (f->data[grid_ref(&f->grid, new_not_old[i].x, new_not_old[i].y, new_not_old[i].z)].distance_square != 0)
      new_not_old[c++] = new_not_old[i];
  pdf_remove_obstacle_voxels(f, old_not_new, a);
  pdf_add_new_obstacle_voxels(f, new_not_old, c);
  free(so); free(sn); free(old_not_new); free(new_not_old);
}

void orc_pdf_dims(const orc_pdf* f, int32_t* n_xyz, int32_t* max_d2)
{
  n_xyz[0] = f->grid.num_cells[0];
  n_xyz[1] = f->grid.num_cells[1];
  n_xyz[2] = f->grid.num_cells[2];
  *max_d2 = f->max_distance_sq;
}
void orc_pdf_get_d2(const orc_pdf* f, int32_t* out)
{
  for (long i = 0; i < f->grid.num_cells_total; ++i)
    out[i] = f->data[i].distance_square;
}
void orc_pdf_get_neg_d2(const orc_pdf* f, int32_t* out)
{
  for (long i = 0; i < f->grid.num_cells_total; ++i)
    out[i] = f->data[i].negative_distance_square;
}
const orc_grid* orc_pdf_grid(const orc_pdf* f) { return &f->grid; }

static inline double pdf_voxel_distance(const orc_pdf* f, const orc_voxel* v)
{
  return f->sqrt_table[v->distance_square] - f->sqrt_table[v->negative_distance_square]; /* hpp:604-607 */
}
double orc_pdf_get_cell_distance(const orc_pdf* f, int x, int y, int z)
{
  return pdf_voxel_distance(f, &f->data[grid_ref(&f->grid, x, y, z)]);
}
/* PropagationDistanceField::getDistance(double,double,double): VoxelGrid::operator() returns the default
 * object (max_d2, 0) when out of range (voxel_grid.hpp:453-461, propagation_distance_field.cpp:594-597). */
double orc_pdf_get_distance(const orc_pdf* f, double x, double y, double z)
{
  int c[3];
  if (!orc_grid_world_to_grid(&f->grid, x, y, z, c))
    return f->sqrt_table[f->max_distance_sq] - f->sqrt_table[0];
  return pdf_voxel_distance(f, &f->data[grid_ref(&f->grid, c[0], c[1], c[2])]);
}

/* DistanceField::getDistanceGradient: distance_field/src/distance_field.cpp:73-97 */
double orc_pdf_get_distance_gradient(const orc_pdf* f, double x, double y, double z, double* grad, int* in_bounds)
{
  const orc_grid* g = &f->grid;
  int gx = grid_cell(g, 0, x), gy = grid_cell(g, 1, y), gz = grid_cell(g, 2, z);
  if (gx < 1 || gy < 1 || gz < 1 || gx >= g->num_cells[0] - 1 || gy >= g->num_cells[1] - 1 || gz >= g->num_cells[2] - 1)
  {
    grad[0] = grad[1] = grad[2] = 0.0;
    *in_bounds = 0;
    return f->max_distance; /* getUninitializedDistance() */
  }
#define D(a, b, c) pdf_voxel_distance(f, &f->data[grid_ref(g, (a), (b), (c))])
  grad[0] = (D(gx + 1, gy, gz) - D(gx - 1, gy, gz)) * f->inv_twice_resolution;
  grad[1] = (D(gx, gy + 1, gz) - D(gx, gy - 1, gz)) * f->inv_twice_resolution;
  grad[2] = (D(gx, gy, gz + 1) - D(gx, gy, gz - 1)) * f->inv_twice_resolution;
  *in_bounds = 1;
  return D(gx, gy, gz);
#undef D
}

/* ------------------------------------------------------------------------------------------------------
 * Forward kinematics: robot_state/src/robot_state.cpp:793-848 + robot_model/src/{revolute,prismatic,planar,floating,fixed}_joint_model.cpp
 * Matrices are 4x4 column-major doubles, as Eigen::Isometry3d stores them.
 * ---------------------------------------------------------------------------------------------------- */
enum { ORC_FIXED = 0, ORC_REVOLUTE = 1, ORC_PRISMATIC = 2, ORC_PLANAR = 3, ORC_FLOATING = 4 };

typedef struct
{
  int32_t n_links, n_vars;
  const int32_t* parent;             /* [n_links] parent link index, -1 for the root link */
  const int32_t* joint_type;         /* [n_links] type of the link's parent joint */
  const double* axis;                /* [n_links*3] */
  const double* origin;              /* [n_links*16] joint origin transform, column-major */
  const int32_t* origin_is_identity; /* [n_links] */
  const int32_t* first_var;          /* [n_links] first variable of the parent joint in the full state */
  int32_t n_group_vars;
  const int32_t* group_var_index;    /* [n_group_vars] JointModelGroup::getVariableIndexList() */
  int32_t n_mimic;
  const int32_t* mimic_dst;
  const int32_t* mimic_src;
  const double* mimic_factor;
  const double* mimic_offset;
} orc_robot;

#define M(m, r, c) (m)[(c) * 4 + (r)]

/* affine(3x4) * matrix(4x4) -> affine; Eigen accumulates k = 0..3 in order */
static void mul_affine(const double* a, const double* b, double* out)
{
  double t[16];
  for (int c = 0; c < 4; ++c)
  {
    for (int r = 0; r < 3; ++r)
    {
      double s = M(a, r, 0) * M(b, 0, c);
      s += M(a, r, 1) * M(b, 1, c);
      s += M(a, r, 2) * M(b, 2, c);
      s += M(a, r, 3) * M(b, 3, c);
      M(t, r, c) = s;
    }
    M(t, 3, c) = (c == 3) ? 1.0 : 0.0;
  }
  memcpy(out, t, sizeof(t));
}

static void set_identity(double* m)
{
  memset(m, 0, 16 * sizeof(double));
  m[0] = m[5] = m[10] = m[15] = 1.0;
}

static void quat_to_rot(double x, double y, double z, double w, double* m) /* Eigen toRotationMatrix */
{
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  M(m, 0, 0) = 1.0 - (tyy + tzz); M(m, 0, 1) = txy - twz; M(m, 0, 2) = txz + twy;
  M(m, 1, 0) = txy + twz; M(m, 1, 1) = 1.0 - (txx + tzz); M(m, 1, 2) = tyz - twx;
  M(m, 2, 0) = txz - twy; M(m, 2, 1) = tyz + twx; M(m, 2, 2) = 1.0 - (txx + tyy);
}

static void joint_transform(int type, const double* axis, const double* v, double* d)
{
  set_identity(d);
  switch (type)
  {
    case ORC_REVOLUTE: /* revolute_joint_model.cpp:250-287; x2_ etc. precomputed at :73-82 */
    {
      const double x2 = axis[0] * axis[0], y2 = axis[1] * axis[1], z2 = axis[2] * axis[2];
      const double xy = axis[0] * axis[1], xz = axis[0] * axis[2], yz = axis[1] * axis[2];
      const double c = cos(v[0]), s = sin(v[0]);
      const double t = 1.0 - c;
      const double txy = t * xy, txz = t * xz, tyz = t * yz;
      const double zs = axis[2] * s, ys = axis[1] * s, xs = axis[0] * s;
      d[0] = t * x2 + c; d[1] = txy + zs; d[2] = txz - ys;
      d[4] = txy - zs; d[5] = t * y2 + c; d[6] = tyz + xs;
      d[8] = txz + ys; d[9] = tyz - xs; d[10] = t * z2 + c;
      break;
    }
    case ORC_PRISMATIC: /* prismatic_joint_model.cpp:124-146 */
      d[12] = axis[0] * v[0]; d[13] = axis[1] * v[0]; d[14] = axis[2] * v[0];
      break;
    case ORC_PLANAR: /* planar_joint_model.cpp:345-349; Eigen AngleAxis(theta, UnitZ).toRotationMatrix() */
    {
      const double c = cos(v[2]), s = sin(v[2]);
      M(d, 0, 0) = c; M(d, 0, 1) = -s; M(d, 1, 0) = s; M(d, 1, 1) = c;
      M(d, 2, 2) = (1.0 - c) + c;
      M(d, 0, 3) = v[0]; M(d, 1, 3) = v[1];
      break;
    }
    case ORC_FLOATING: /* floating_joint_model.cpp:231-236; Quaterniond(w,x,y,z).normalized() */
    {
      const double n = sqrt(v[3] * v[3] + v[4] * v[4] + v[5] * v[5] + v[6] * v[6]);
      quat_to_rot(v[3] / n, v[4] / n, v[5] / n, v[6] / n, d);
      M(d, 0, 3) = v[0]; M(d, 1, 3) = v[1]; M(d, 2, 3) = v[2];
      break;
    }
    default: break; /* fixed_joint_model.cpp:95-98 */
  }
}

/* setJointGroupPositions + updateMimicJoints(group): robot_state.cpp:571-584, 210-220 */
static void set_group_positions(const orc_robot* r, double* positions, const double* gstate)
{
  for (int i = 0; i < r->n_group_vars; ++i)
    positions[r->group_var_index[i]] = gstate[i];
  for (int i = 0; i < r->n_mimic; ++i)
    positions[r->mimic_dst[i]] = r->mimic_factor[i] * positions[r->mimic_src[i]] + r->mimic_offset[i];
}

/* updateLinkTransformsInternal: robot_state.cpp:793-840.  out: n_links 4x4 column-major */
static void fk_links(const orc_robot* r, const double* positions, double* out)
{
  double jt[16], tmp[16];
  for (int l = 0; l < r->n_links; ++l)
  {
    double* T = out + 16 * l;
    const double* origin = r->origin + 16 * l;
    const int type = r->joint_type[l];
    if (r->parent[l] >= 0)
    {
      const double* P = out + 16 * r->parent[l];
      if (type == ORC_FIXED)
        mul_affine(P, origin, T);
      else
      {
        joint_transform(type, r->axis + 3 * l, positions + r->first_var[l], jt);
        if (r->origin_is_identity[l])
          mul_affine(P, jt, T);
        else
        {
          mul_affine(P, origin, tmp);
          mul_affine(tmp, jt, T);
        }
      }
    }
    else
    {
      if (type == ORC_FIXED)
        set_identity(T);
      else
      {
        joint_transform(type, r->axis + 3 * l, positions + r->first_var[l], jt);
        if (r->origin_is_identity[l])
          memcpy(T, jt, sizeof(jt));
        else
          mul_affine(origin, jt, T);
      }
    }
  }
}

void orc_fk_batch(const orc_robot* r, const double* base_positions, const double* q, size_t n, double* out)
{
  double* pos = (double*)malloc((size_t)r->n_vars * sizeof(double));
  for (size_t i = 0; i < n; ++i)
  {
    memcpy(pos, base_positions, (size_t)r->n_vars * sizeof(double));
    set_group_positions(r, pos, q + i * (size_t)r->n_group_vars);
    fk_links(r, pos, out + i * (size_t)r->n_links * 16);
  }
  free(pos);
}

/* ------------------------------------------------------------------------------------------------------
 * Collision checks: collision_distance_field/src/collision_env_distance_field.cpp and
 * collision_distance_field_types.cpp
 * ---------------------------------------------------------------------------------------------------- */
typedef struct
{
  int32_t n_glinks;                /* dfce->link_names_ = group->getUpdatedLinkModelNames() (:735) */
  const int32_t* glink_index;      /* model link index of each */
  const int32_t* has_geometry;     /* dfce->link_has_geometry_ (:770-837) */
  const int32_t* self_enabled;     /* dfce->self_collision_enabled_ */
  const uint8_t* intra_enabled;    /* dfce->intra_group_collision_enabled_, [n_glinks*n_glinks], only j>i read */
  const int32_t* sphere_start;     /* [n_glinks+1] */
  const double* sphere_rel;        /* CollisionSphere::relative_vec_ */
  const double* sphere_radius;     /* CollisionSphere::radius_ */
  const double* bs_center;         /* BodyDecomposition::relative_bounding_sphere_.center [n_glinks*3] */
  const double* bs_radius;         /* ... .radius */
  const int32_t* point_start;      /* [n_glinks+1] interior collision points (posed on every check by the reference) */
  const double* point_rel;
  double max_propagation_distance; /* max_propogation_distance_ */
  double collision_tolerance;      /* collision_tolerance_ */
  const int32_t* body_id;          /* NULL, or [n_glinks]: -1 for links; the entries that are the SHAPES of one attached body
                                      share an id (the reference keeps such a body as ONE entry of several decompositions) */
} orc_collision_model;

enum { ORC_CHECK_WORLD = 1, ORC_CHECK_SELF = 2, ORC_CHECK_INTRA = 4 };

static inline void transform_point(const double* T, const double* p, double* o)
{ /* Isometry3d * Vector3d: linear * p + translation, k in order */
  for (int r = 0; r < 3; ++r)
  {
    double s = M(T, r, 0) * p[0];
    s += M(T, r, 1) * p[1];
    s += M(T, r, 2) * p[2];
    o[r] = s + M(T, r, 3);
  }
}

/* The link-level cull of getIntraGroupCollisions (collision_env_distance_field.cpp:452-507).  Two links: one
 * doBoundingSpheresIntersect (types.cpp:408-418, reference quirk kept: SQUARED centre distance against the plain radius
 * sum).  An attached body is tested shape by shape and passes AS A WHOLE as soon as any of its shapes' bounding spheres
 * passes (:461-505: all_ok turns false on the first hit; every sphere of the body is then tested).  Entries here are per
 * shape, so for an entry of a several-shape body the test runs over all entries with its body_id. */
static int bounds_pass(const orc_collision_model* cm, const double* bs_centers, int i, int j)
{
  const double* ci = bs_centers + 3 * i;
  const double* cj = bs_centers + 3 * j;
  const double ex = ci[0] - cj[0], ey = ci[1] - cj[1], ez = ci[2] - cj[2];
  return (ex * ex + ey * ey + ez * ez) < (cm->bs_radius[i] + cm->bs_radius[j]);
}
static int bounds_meet(const orc_collision_model* cm, const double* bs_centers, int i, int j)
{
  if (bounds_pass(cm, bs_centers, i, j))
    return 1;
  if (!cm->body_id)
    return 0;
  const int bi = cm->body_id[i], bj = cm->body_id[j];
  if (bi < 0 && bj < 0)
    return 0;
  for (int u = 0; u < cm->n_glinks; ++u)
  {
    if (!(u == i || (bi >= 0 && cm->body_id[u] == bi)) || !cm->has_geometry[u])
      continue;
    for (int v = 0; v < cm->n_glinks; ++v)
      if ((v == j || (bj >= 0 && cm->body_id[v] == bj)) && cm->has_geometry[v] && bounds_pass(cm, bs_centers, u, v))
        return 1;
  }
  return 0;
}

/* getCollisionSphereCollision (bool overload): collision_distance_field_types.cpp:211-236 */
static int sphere_list_collides(const orc_pdf* f, const double* centers, const double* radii, int n,
                                double maximum_value, double tolerance, int faithful)
{
  for (int i = 0; i < n; ++i)
  {
    const double* p = centers + 3 * i;
    double dist;
    if (faithful)
    {
      double grad[3];
      int in_bounds = 1;
      dist = orc_pdf_get_distance_gradient(f, p[0], p[1], p[2], grad, &in_bounds);
      if (!in_bounds && sqrt(grad[0] * grad[0] + grad[1] * grad[1] + grad[2] * grad[2]) > 0)
        return 1;
    }
    else
    { /* lean: centre voxel only -- the gradient does not influence the boolean (it is zero when out of bounds) */
      const orc_grid* g = &f->grid;
      int gx = grid_cell(g, 0, p[0]), gy = grid_cell(g, 1, p[1]), gz = grid_cell(g, 2, p[2]);
      if (gx < 1 || gy < 1 || gz < 1 || gx >= g->num_cells[0] - 1 || gy >= g->num_cells[1] - 1 ||
          gz >= g->num_cells[2] - 1)
        dist = f->max_distance;
      else
        dist = pdf_voxel_distance(f, &f->data[grid_ref(g, gx, gy, gz)]);
    }
    if ((maximum_value > dist) && (radii[i] - dist > tolerance))
      return 1;
  }
  return 0;
}

typedef struct
{
  double* positions;
  double* link_T;     /* n_links * 16 */
  double* centers;    /* n_spheres * 3 */
  double* bs_centers; /* n_glinks * 3 */
  double* points;     /* posed interior points (faithful mode) */
} orc_scratch;

/* One state through PlanningScene::checkCollision order: world (checkRobotCollision), then self field + intra
 * group (checkSelfCollision) -- planning_scene/src/planning_scene.cpp:436-455.  Returns 1 if in collision. */
static int check_state(const orc_robot* r, const orc_collision_model* cm, const orc_pdf* world, const orc_pdf* self,
                       const double* base_positions, const double* q, int flags, int faithful, orc_scratch* s)
{
  memcpy(s->positions, base_positions, (size_t)r->n_vars * sizeof(double));
  set_group_positions(r, s->positions, q);
  fk_links(r, s->positions, s->link_T);
  /* PosedBodySphereDecomposition::updatePose: types.cpp:388-406 (posed from the LINK frame,
   * collision_env_distance_field.cpp:1111,1222) */
  for (int i = 0; i < cm->n_glinks; ++i)
  {
    if (!cm->has_geometry[i])
      continue;
    const double* T = s->link_T + 16 * cm->glink_index[i];
    transform_point(T, cm->bs_center + 3 * i, s->bs_centers + 3 * i);
    for (int k = cm->sphere_start[i]; k < cm->sphere_start[i + 1]; ++k)
      transform_point(T, cm->sphere_rel + 3 * k, s->centers + 3 * k);
    if (faithful && cm->point_start)
      for (int k = cm->point_start[i]; k < cm->point_start[i + 1]; ++k)
        transform_point(T, cm->point_rel + 3 * k, s->points + 3 * k);
  }
  const double maxv = cm->max_propagation_distance, tol = cm->collision_tolerance;
  if ((flags & ORC_CHECK_WORLD) && world)
  { /* getEnvironmentCollisions: collision_env_distance_field.cpp:1561-1643 (ACM not consulted) */
    for (int i = 0; i < cm->n_glinks; ++i)
    {
      if (!cm->has_geometry[i])
        continue;
      int a = cm->sphere_start[i], n = cm->sphere_start[i + 1] - a;
      if (sphere_list_collides(world, s->centers + 3 * a, cm->sphere_radius + a, n, maxv, tol, faithful))
        return 1;
    }
  }
  if ((flags & ORC_CHECK_SELF) && self)
  { /* getSelfCollisions: :274-353 */
    for (int i = 0; i < cm->n_glinks; ++i)
    {
      if (!cm->has_geometry[i] || !cm->self_enabled[i])
        continue;
      int a = cm->sphere_start[i], n = cm->sphere_start[i + 1] - a;
      if (sphere_list_collides(self, s->centers + 3 * a, cm->sphere_radius + a, n, maxv, tol, faithful))
        return 1;
    }
  }
  if (flags & ORC_CHECK_INTRA)
  { /* getIntraGroupCollisions: :430-642 with doBoundingSpheresIntersect types.cpp:408-418 */
    const int L = cm->n_glinks;
    for (int i = 0; i < L; ++i)
      for (int j = i + 1; j < L; ++j)
      {
        if (!cm->has_geometry[i] || !cm->has_geometry[j])
          continue;
        if (!cm->intra_enabled[(size_t)i * L + j])
          continue;
        if (!bounds_meet(cm, s->bs_centers, i, j))
          continue;
        for (int k = cm->sphere_start[i]; k < cm->sphere_start[i + 1]; ++k)
          for (int l = cm->sphere_start[j]; l < cm->sphere_start[j + 1]; ++l)
          {
            const double gx = s->centers[3 * k] - s->centers[3 * l];
            const double gy = s->centers[3 * k + 1] - s->centers[3 * l + 1];
            const double gz = s->centers[3 * k + 2] - s->centers[3 * l + 2];
            const double dist = sqrt(gx * gx + gy * gy + gz * gz);
            if (dist < cm->sphere_radius[k] + cm->sphere_radius[l])
              return 1;
          }
      }
  }
  return 0;
}

static void scratch_alloc(orc_scratch* s, const orc_robot* r, const orc_collision_model* cm)
{
  int ns = cm->sphere_start[cm->n_glinks];
  int np = cm->point_start ? cm->point_start[cm->n_glinks] : 0;
  s->positions = (double*)malloc((size_t)r->n_vars * sizeof(double));
  s->link_T = (double*)malloc((size_t)r->n_links * 16 * sizeof(double));
  s->centers = (double*)malloc((size_t)(ns + 1) * 3 * sizeof(double));
  s->bs_centers = (double*)malloc((size_t)(cm->n_glinks + 1) * 3 * sizeof(double));
  s->points = (double*)malloc((size_t)(np + 1) * 3 * sizeof(double));
}
static void scratch_free(orc_scratch* s)
{
  free(s->positions); free(s->link_T); free(s->centers); free(s->bs_centers); free(s->points);
}

/* Batch of N group-state vectors -> collision byte per state (1 = in collision).  threads <= 0: all cores. */
int orc_check_batch(const orc_robot* r, const orc_collision_model* cm, const orc_pdf* world, const orc_pdf* self,
                    const double* base_positions, const double* q, size_t n, int flags, int faithful, int threads,
                    uint8_t* collision)
{
  int used = 1;
#ifdef _OPENMP
  if (threads <= 0)
    threads = omp_get_max_threads();
  used = threads;
#pragma omp parallel num_threads(threads)
#endif
  {
    orc_scratch s;
    scratch_alloc(&s, r, cm);
#ifdef _OPENMP
#pragma omp for schedule(static, 256)
#endif
    for (long i = 0; i < (long)n; ++i)
    {
      if (faithful)
      { /* the reference heap-allocates a fresh GroupStateRepresentation per check (collision_common_distance_field.hpp:61-82) */
        orc_scratch t;
        scratch_alloc(&t, r, cm);
        collision[i] = (uint8_t)check_state(r, cm, world, self, base_positions, q + (size_t)i * r->n_group_vars, flags, 1, &t);
        scratch_free(&t);
      }
      else
        collision[i] = (uint8_t)check_state(r, cm, world, self, base_positions, q + (size_t)i * r->n_group_vars, flags, 0, &s);
    }
    scratch_free(&s);
  }
  return used;
}

/* Posed sphere centres of one state (for tests of sphere posing / clearance bucketing).
 * min_clearance: min over checked spheres/fields of (dist - radius), +inf if nothing in range. */
void orc_state_spheres(const orc_robot* r, const orc_collision_model* cm, const double* base_positions,
                       const double* q, double* centers_out)
{
  orc_scratch s;
  scratch_alloc(&s, r, cm);
  memcpy(s.positions, base_positions, (size_t)r->n_vars * sizeof(double));
  set_group_positions(r, s.positions, q);
  fk_links(r, s.positions, s.link_T);
  for (int i = 0; i < cm->n_glinks; ++i)
  {
    if (!cm->has_geometry[i])
      continue;
    const double* T = s.link_T + 16 * cm->glink_index[i];
    for (int k = cm->sphere_start[i]; k < cm->sphere_start[i + 1]; ++k)
      transform_point(T, cm->sphere_rel + 3 * k, centers_out + 3 * k);
  }
  scratch_free(&s);
}

/* Signed margin of one state: the smallest (dist - radius) over world/self sphere reads and
 * (centre distance - radius sum) over enabled, un-culled intra pairs.  Used only to bucket GPU/oracle
 * mismatches by "reference clearance within 1e-4 m of zero" as north_star requires. */
double orc_state_margin(const orc_robot* r, const orc_collision_model* cm, const orc_pdf* world, const orc_pdf* self,
                        const double* base_positions, const double* q, int flags)
{
  orc_scratch s;
  scratch_alloc(&s, r, cm);
  memcpy(s.positions, base_positions, (size_t)r->n_vars * sizeof(double));
  set_group_positions(r, s.positions, q);
  fk_links(r, s.positions, s.link_T);
  double best = INFINITY;
  const int L = cm->n_glinks;
  for (int i = 0; i < L; ++i)
  {
    if (!cm->has_geometry[i])
      continue;
    const double* T = s.link_T + 16 * cm->glink_index[i];
    transform_point(T, cm->bs_center + 3 * i, s.bs_centers + 3 * i);
    for (int k = cm->sphere_start[i]; k < cm->sphere_start[i + 1]; ++k)
    {
      double* c = s.centers + 3 * k;
      transform_point(T, cm->sphere_rel + 3 * k, c);
      for (int w = 0; w < 2; ++w)
      {
        const orc_pdf* f = w == 0 ? world : self;
        if (!f || !(flags & (w == 0 ? ORC_CHECK_WORLD : ORC_CHECK_SELF)))
          continue;
        if (w == 1 && !cm->self_enabled[i])
          continue;
        double grad[3];
        int inb;
        double dist = orc_pdf_get_distance_gradient(f, c[0], c[1], c[2], grad, &inb);
        if (cm->max_propagation_distance > dist)
        {
          double m = dist - cm->sphere_radius[k] + cm->collision_tolerance;
          if (m < best)
            best = m;
        }
      }
    }
  }
  if (flags & ORC_CHECK_INTRA)
    for (int i = 0; i < L; ++i)
      for (int j = i + 1; j < L; ++j)
      {
        if (!cm->has_geometry[i] || !cm->has_geometry[j] || !cm->intra_enabled[(size_t)i * L + j])
          continue;
        if (!bounds_meet(cm, s.bs_centers, i, j))
          continue;
        for (int k = cm->sphere_start[i]; k < cm->sphere_start[i + 1]; ++k)
          for (int l = cm->sphere_start[j]; l < cm->sphere_start[j + 1]; ++l)
          {
            const double gx = s.centers[3 * k] - s.centers[3 * l];
            const double gy = s.centers[3 * k + 1] - s.centers[3 * l + 1];
            const double gz = s.centers[3 * k + 2] - s.centers[3 * l + 2];
            double m = sqrt(gx * gx + gy * gy + gz * gz) - (cm->sphere_radius[k] + cm->sphere_radius[l]);
            if (m < best)
              best = m;
          }
      }
  scratch_free(&s);
  return best;
}

/* ------------------------------------------------------------------------------------------------------
 * CollisionEnvDistanceField::getCollisionGradients (collision_env_distance_field.cpp:1517-1538) for n states:
 *   getSelfProximityGradients        :355-428   -- the dfce distance_field_ part (type SELF).  The per-link
 *                                                  PosedDistanceFields are consulted only when dfce->acm_.getSize() > 0
 *                                                  (:386), i.e. never in CHOMP's optimisation loop, which passes
 *                                                  acm = nullptr (chomp_optimizer.cpp:890); they are NOT restated here.
 *   getIntraGroupProximityGradients  :644-709   -- every sphere pair of every enabled link pair (type INTRA)
 *   getEnvironmentProximityGradients :1645-1681 -- world field (type ENVIRONMENT)
 * with getCollisionSphereGradients (collision_distance_field_types.cpp:149-209; subtract_radii = false,
 * maximum_value = max_propogation_distance_, stop_at_first_collision = false).
 * Outputs per state: distances [n_spheres] (DBL_MAX where nothing was recorded), gradients [n_spheres*3] (0 where
 * nothing was recorded; the reference leaves them uninitialised), types [n_spheres] (CollisionType: NONE 0, SELF 1,
 * INTRA 2, ENVIRONMENT 3), closest [n_glinks] (GradientInfo::closest_distance).
 * ---------------------------------------------------------------------------------------------------- */
static void sphere_gradients(const orc_pdf* f, const orc_collision_model* cm, int i, const double* centers, int type,
                             double* dist_out, double* grad_out, uint8_t* type_out, double* closest)
{
  for (int k = cm->sphere_start[i]; k < cm->sphere_start[i + 1]; ++k)
  {
    double grad[3];
    int inb;
    const double* c = centers + 3 * k;
    double dist = orc_pdf_get_distance_gradient(f, c[0], c[1], c[2], grad, &inb);
    /* (!in_bounds && grad.norm() > EPSILON) cannot happen: the gradient is zero when out of bounds */
    if (dist < cm->max_propagation_distance)
    {
      if (dist < closest[i])
        closest[i] = dist;
      if (dist < dist_out[k])
      {
        type_out[k] = (uint8_t)type;
        dist_out[k] = dist;
        grad_out[3 * k] = grad[0];
        grad_out[3 * k + 1] = grad[1];
        grad_out[3 * k + 2] = grad[2];
      }
    }
  }
}

/* PosedDistanceField::getCollisionSphereGradients (collision_distance_field_types.cpp:87-147) on the PosedDistanceField of link
 * j for the spheres of entry i, with PosedDistanceField::getDistanceGradient (collision_distance_field_types.hpp:163-176)
 * EXACTLY as written there:
 *   rel_pos = pose_.linear().transpose() * p        -- the pose's translation is NOT subtracted
 *   grad    = pose_ * (gx, gy, gz)                  -- an isometry times a vector: the translation IS added
 * so an out-of-bounds query returns the link's translation as its "gradient", grad.norm() > 0 (types.cpp:104) is true for
 * every link away from the world origin, and the loop over the entry's spheres ends there ("return true"), without looking
 * at the spheres after it.  Tj: link j's global transform, 4x4 column-major. */
static void sphere_gradients_posed(const orc_pdf* f, const double* Tj, const orc_collision_model* cm, int i,
                                   const double* centers, double* dist_out, double* grad_out, uint8_t* type_out, double* closest)
{
  for (int k = cm->sphere_start[i]; k < cm->sphere_start[i + 1]; ++k)
  {
    const double* c = centers + 3 * k;
    double rel[3], gl[3], grad[3];
    int inb;
    for (int a = 0; a < 3; ++a) /* (R^T p)_a = R(0,a) p0 + R(1,a) p1 + R(2,a) p2 */
      rel[a] = (M(Tj, 0, a) * c[0] + M(Tj, 1, a) * c[1]) + M(Tj, 2, a) * c[2];
    double dist = orc_pdf_get_distance_gradient(f, rel[0], rel[1], rel[2], gl, &inb);
    for (int a = 0; a < 3; ++a) /* linear() * g + translation() */
      grad[a] = ((M(Tj, a, 0) * gl[0] + M(Tj, a, 1) * gl[1]) + M(Tj, a, 2) * gl[2]) + M(Tj, a, 3);
    if (!inb && sqrt((grad[0] * grad[0] + grad[1] * grad[1]) + grad[2] * grad[2]) > 0.0)
      return; /* "out of bounds": return true */
    if (dist < cm->max_propagation_distance)
    {
      if (dist < closest[i])
        closest[i] = dist;
      if (dist < dist_out[k])
      {
        type_out[k] = 1; /* SELF */
        dist_out[k] = dist;
        grad_out[3 * k] = grad[0];
        grad_out[3 * k + 1] = grad[1];
        grad_out[3 * k + 2] = grad[2];
      }
    }
  }
}

void orc_collision_gradients_lf(const orc_robot* r, const orc_collision_model* cm, const orc_pdf* world, const orc_pdf* self,
                                const double* base_positions, const double* q, size_t n, double* distances,
                                double* gradients, uint8_t* types, double* closest, const orc_pdf* const* link_fields,
                                const uint8_t* lf_enabled);

void orc_collision_gradients(const orc_robot* r, const orc_collision_model* cm, const orc_pdf* world, const orc_pdf* self,
                             const double* base_positions, const double* q, size_t n, double* distances, double* gradients,
                             uint8_t* types, double* closest)
{
  orc_collision_gradients_lf(r, cm, world, self, base_positions, q, n, distances, gradients, types, closest, NULL, NULL);
}

/* link_fields [n_glinks] (NULL entries: no geometry) and lf_enabled [n_glinks * n_glinks]: the per-link PosedDistanceFields
 * of the GroupStateRepresentation (collision_env_distance_field.cpp:1190-1197) and, for entry i, which links j they are
 * consulted for -- getSelfProximityGradients :386-411: only when the cache entry's ACM is not empty, j's name differs from
 * i's, and the ACM has no entry (i, j) other than NEVER.  Both NULL: the ACM is empty (:386 false). */
void orc_collision_gradients_lf(const orc_robot* r, const orc_collision_model* cm, const orc_pdf* world, const orc_pdf* self,
                                const double* base_positions, const double* q, size_t n, double* distances,
                                double* gradients, uint8_t* types, double* closest, const orc_pdf* const* link_fields,
                                const uint8_t* lf_enabled)
{
  const int L = cm->n_glinks, S = cm->sphere_start[L];
  orc_scratch s;
  scratch_alloc(&s, r, cm);
  for (size_t st = 0; st < n; ++st)
  {
    double* D = distances + st * (size_t)S;
    double* G = gradients + st * (size_t)S * 3;
    uint8_t* Ty = types + st * (size_t)S;
    double* Cl = closest + st * (size_t)L;
    memcpy(s.positions, base_positions, (size_t)r->n_vars * sizeof(double));
    set_group_positions(r, s.positions, q + st * (size_t)r->n_group_vars);
    fk_links(r, s.positions, s.link_T);
    for (int i = 0; i < L; ++i)
    {
      Cl[i] = DBL_MAX;
      const double* T = s.link_T + 16 * cm->glink_index[i];
      for (int k = cm->sphere_start[i]; k < cm->sphere_start[i + 1]; ++k)
      {
        transform_point(T, cm->sphere_rel + 3 * k, s.centers + 3 * k);
        D[k] = DBL_MAX;
        Ty[k] = 0;
        G[3 * k] = G[3 * k + 1] = G[3 * k + 2] = 0.0;
      }
    }
    for (int i = 0; i < L; ++i)
      if (cm->has_geometry[i] && cm->self_enabled[i])
      { /* getSelfProximityGradients :355-428: the other links' posed fields first, then the cache entry's self field */
        if (link_fields && lf_enabled)
          for (int j = 0; j < L; ++j)
            if (lf_enabled[(size_t)i * L + j] && link_fields[j])
              sphere_gradients_posed(link_fields[j], s.link_T + 16 * cm->glink_index[j], cm, i, s.centers, D, G, Ty, Cl);
        if (self)
          sphere_gradients(self, cm, i, s.centers, 1, D, G, Ty, Cl);
      }
    for (int i = 0; i < L; ++i)
      for (int j = i + 1; j < L; ++j)
      {
        if (!cm->has_geometry[i] || !cm->has_geometry[j] || !cm->intra_enabled[(size_t)i * L + j])
          continue;
        for (int k = cm->sphere_start[i]; k < cm->sphere_start[i + 1]; ++k)
          for (int l = cm->sphere_start[j]; l < cm->sphere_start[j + 1]; ++l)
          {
            const double gx = s.centers[3 * k] - s.centers[3 * l];
            const double gy = s.centers[3 * k + 1] - s.centers[3 * l + 1];
            const double gz = s.centers[3 * k + 2] - s.centers[3 * l + 2];
            const double dist = sqrt(gx * gx + gy * gy + gz * gz);
            if (dist < D[k])
            {
              D[k] = dist;
              G[3 * k] = gx; G[3 * k + 1] = gy; G[3 * k + 2] = gz;
              Ty[k] = 2;
            }
            if (dist < D[l])
            {
              D[l] = dist;
              G[3 * l] = -gx; G[3 * l + 1] = -gy; G[3 * l + 2] = -gz;
              Ty[l] = 2;
            }
          }
      }
    for (int i = 0; i < L; ++i)
      if (cm->has_geometry[i] && world)
        sphere_gradients(world, cm, i, s.centers, 3, D, G, Ty, Cl);
  }
  scratch_free(&s);
}

/* ------------------------------------------------------------------------------------------------------
 * CollisionEnvDistanceField::checkCollision with req.contacts = true (collision_env_distance_field.cpp:1389-1411):
 * getSelfCollisions (:274-353) -> getIntraGroupCollisions (:430-642) -> getEnvironmentCollisions (:1561-1643), each
 * stopping the sequence once res.contact_count >= req.max_contacts; the contact-listing overload of
 * getCollisionSphereCollision (collision_distance_field_types.cpp:238-271).
 * A contact: pos = centre of the colliding sphere (of body 1), body_1 = dfce entry index, body_2 = the other entry
 * (intra), -1 = "self", -2 = "environment"; sphere_1 / sphere_2 = indices into the entry's sphere list (-1: none).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct
{
  int32_t body_1, body_2, sphere_1, sphere_2;
  double pos[3];
} orc_contact;

static int field_contacts(const orc_pdf* f, const orc_collision_model* cm, int i, const double* centers, unsigned num_coll,
                          int* colls, int* n_colls)
{ /* getCollisionSphereCollision(..., num_coll, colls): types.cpp:238-271 */
  *n_colls = 0;
  const double maxv = cm->max_propagation_distance, tol = cm->collision_tolerance;
  for (int k = cm->sphere_start[i]; k < cm->sphere_start[i + 1]; ++k)
  {
    double grad[3];
    int inb;
    const double* c = centers + 3 * k;
    const double dist = orc_pdf_get_distance_gradient(f, c[0], c[1], c[2], grad, &inb);
    if (maxv > dist && (cm->sphere_radius[k] - dist > tol))
    {
      if (num_coll == 0)
        return 1;
      colls[(*n_colls)++] = k - cm->sphere_start[i];
      if ((unsigned)*n_colls >= num_coll)
        return 1;
    }
  }
  return *n_colls > 0;
}

void orc_check_contacts(const orc_robot* r, const orc_collision_model* cm, const orc_pdf* world, const orc_pdf* self,
                        const double* base_positions, const double* q, size_t n, int flags, unsigned max_contacts,
                        unsigned max_contacts_per_pair, uint8_t* collision, int32_t* contact_count, orc_contact* contacts)
{
  const int L = cm->n_glinks, S = cm->sphere_start[L];
  orc_scratch s;
  scratch_alloc(&s, r, cm);
  int* colls = (int*)malloc((size_t)(S + 1) * sizeof(int));
  for (size_t st = 0; st < n; ++st)
  {
    orc_contact* out = contacts + st * (size_t)max_contacts;
    unsigned count = 0;
    int coll_any = 0, done = 0;
    memcpy(s.positions, base_positions, (size_t)r->n_vars * sizeof(double));
    set_group_positions(r, s.positions, q + st * (size_t)r->n_group_vars);
    fk_links(r, s.positions, s.link_T);
    for (int i = 0; i < L; ++i)
    {
      const double* T = s.link_T + 16 * cm->glink_index[i];
      transform_point(T, cm->bs_center + 3 * i, s.bs_centers + 3 * i);
      for (int k = cm->sphere_start[i]; k < cm->sphere_start[i + 1]; ++k)
        transform_point(T, cm->sphere_rel + 3 * k, s.centers + 3 * k);
    }
    for (int pass = 0; pass < 3 && !done; ++pass)
    {
      if (pass == 0 || pass == 2)
      { /* getSelfCollisions / getEnvironmentCollisions */
        const orc_pdf* f = pass == 0 ? self : world;
        if (!f || !(flags & (pass == 0 ? ORC_CHECK_SELF : ORC_CHECK_WORLD)))
          continue;
        for (int i = 0; i < L && !done; ++i)
        {
          if (!cm->has_geometry[i] || (pass == 0 && !cm->self_enabled[i]))
            continue;
          const unsigned left = max_contacts - count;
          int nc = 0;
          const int coll = field_contacts(f, cm, i, s.centers, max_contacts_per_pair < left ? max_contacts_per_pair : left,
                                          colls, &nc);
          if (coll)
          {
            coll_any = 1;
            for (int c = 0; c < nc; ++c)
            {
              orc_contact* o = out + count++;
              o->body_1 = i;
              o->body_2 = pass == 0 ? -1 : -2;
              o->sphere_1 = colls[c];
              o->sphere_2 = -1;
              memcpy(o->pos, s.centers + 3 * (cm->sphere_start[i] + colls[c]), 3 * sizeof(double));
            }
            if (count >= max_contacts)
              done = 1;
          }
        }
        if (count >= max_contacts)
          done = 1;
      }
      else
      { /* getIntraGroupCollisions */
        if (!(flags & ORC_CHECK_INTRA))
          continue;
        for (int i = 0; i < L && !done; ++i)
          for (int j = i + 1; j < L && !done; ++j)
          {
            if (!cm->has_geometry[i] || !cm->has_geometry[j] || !cm->intra_enabled[(size_t)i * L + j])
              continue;
            if (!bounds_meet(cm, s.bs_centers, i, j))
              continue;
            int num_pair = 0; /* res.contacts has no entry for this pair yet: every state starts from an empty result */
            for (int k = cm->sphere_start[i]; k < cm->sphere_start[i + 1] && num_pair < (int)max_contacts_per_pair && !done; ++k)
              for (int l = cm->sphere_start[j]; l < cm->sphere_start[j + 1] && num_pair < (int)max_contacts_per_pair; ++l)
              {
                const double gx = s.centers[3 * k] - s.centers[3 * l];
                const double gy = s.centers[3 * k + 1] - s.centers[3 * l + 1];
                const double gz = s.centers[3 * k + 2] - s.centers[3 * l + 2];
                const double dist = sqrt(gx * gx + gy * gy + gz * gz);
                if (dist < cm->sphere_radius[k] + cm->sphere_radius[l])
                {
                  coll_any = 1;
                  orc_contact* o = out + count++;
                  o->body_1 = i;
                  o->body_2 = j;
                  o->sphere_1 = k - cm->sphere_start[i];
                  o->sphere_2 = l - cm->sphere_start[j];
                  memcpy(o->pos, s.centers + 3 * k, 3 * sizeof(double));
                  num_pair++;
                  if (count >= max_contacts)
                  {
                    done = 1;
                    break;
                  }
                }
              }
          }
      }
    }
    collision[st] = (uint8_t)coll_any;
    contact_count[st] = (int32_t)count;
  }
  free(colls);
  scratch_free(&s);
}

int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
