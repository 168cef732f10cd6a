// kernels_lattice.cuh -- neighbour-table construction on the device: lattice%nncal + lattice%remd
// (lattice.f90:3035-3123, 2823-2907) without the reference's O(kk^2) pair loop.
//
// nncal's result for site i is simply "every j /= i with |r_j - r_i|^2 < ct^2 (minimum image under PBC), ascending j"
// (the pair loop appends smaller indices while row i is processed and larger ones afterwards).  So:
//   1. periodic images of the sites that fall within ct of the bounding box are materialised as ghost points;
//   2. all points are binned into a uniform grid of cell size >= ct (counting sort, integer work);
//   3. one thread per site scans its 27 cells, keeps the candidates that pass the reference's cut-off test
//      (`mapa`: r2 >= ct^2 -> no), sorts them ascending and removes duplicate images;
//   4. remd: the neighbour vectors of each site are matched (|d|^2 < 1e-4) against the vector set of the
//      representative atom of its bravais type; slot k of the output is the neighbour whose vector equals the
//      representative's k-th vector, 0 when the site has no such neighbour (cluster edge).
// Distances are evaluated with explicit round-to-nearest operations in the reference's order of operations.
#pragma once
#include "common.cuh"

struct LatPbc {
  int use, b[3], n[3];
  double a[9];  // lattice%a, column-major (3,3)
  double alat;
};
struct LatGrid {
  double org[3], inv_cs;
  int dim[3];
};

// t_l = ((x n1) a(l,1)) alat, accumulated as in f_wrap_coord_diff: ((odiff + t1) + t2) + t3
__device__ __forceinline__ void lat_shift(const LatPbc &p, const double *od, int x, int y, int z, double *os) {
#pragma unroll
  for (int l = 0; l < 3; l++) {
    double v = od[l];
    v = __dadd_rn(v, __dmul_rn(__dmul_rn((double)(x * p.n[0]), p.a[l]), p.alat));
    v = __dadd_rn(v, __dmul_rn(__dmul_rn((double)(y * p.n[1]), p.a[l + 3]), p.alat));
    v = __dadd_rn(v, __dmul_rn(__dmul_rn((double)(z * p.n[2]), p.a[l + 6]), p.alat));
    os[l] = v;
  }
}
__device__ __forceinline__ double lat_r2(const double *d) {
  return __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2]));
}
// minimum image of crd(:,j) - crd(:,i) (f_wrap_coord_diff, lattice.f90:2975-3018)
__device__ inline void lat_wrap_diff(const LatPbc &p, const double *crd, int i, int j, double *cd) {
  double od[3], md[3];
#pragma unroll
  for (int l = 0; l < 3; l++) { od[l] = __dsub_rn(crd[l + 3 * j], crd[l + 3 * i]); md[l] = od[l]; }
  const int xm = p.b[0] ? 1 : 0, ym = p.b[1] ? 1 : 0, zm = p.b[2] ? 1 : 0;
  for (int z = -zm; z <= zm; z++)
    for (int y = -ym; y <= ym; y++)
      for (int x = -xm; x <= xm; x++) {
        double os[3];
        lat_shift(p, od, x, y, z, os);
        if (sqrt(lat_r2(os)) < sqrt(lat_r2(md))) { md[0] = os[0]; md[1] = os[1]; md[2] = os[2]; }
      }
  cd[0] = md[0]; cd[1] = md[1]; cd[2] = md[2];
}
__device__ __forceinline__ int lat_cell(const LatGrid &g, const double *p) {
  int c[3];
#pragma unroll
  for (int l = 0; l < 3; l++) c[l] = min(max((int)floor((p[l] - g.org[l]) * g.inv_cs), 0), g.dim[l] - 1);
  return (c[2] * g.dim[1] + c[1]) * g.dim[0] + c[0];
}

// pass 0: count (out == null) or write the ghost images (index, shift code) that land inside the padded bounding box
__global__ void k_lat_ghosts(const double *crd, int kk, LatPbc p, LatGrid g, double lo0, double lo1, double lo2, double hi0,
                             double hi1, double hi2, int *counter, int32_t *gidx, double *gpos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kk) return;
  const int xm = p.b[0] ? 1 : 0, ym = p.b[1] ? 1 : 0, zm = p.b[2] ? 1 : 0;
  const double zero[3] = {0.0, 0.0, 0.0};
  for (int z = -zm; z <= zm; z++)
    for (int y = -ym; y <= ym; y++)
      for (int x = -xm; x <= xm; x++) {
        if (!x && !y && !z) continue;
        double sh[3], q[3];
        lat_shift(p, zero, x, y, z, sh);
        for (int l = 0; l < 3; l++) q[l] = crd[l + 3 * i] + sh[l];
        if (q[0] < lo0 || q[1] < lo1 || q[2] < lo2 || q[0] > hi0 || q[1] > hi1 || q[2] > hi2) continue;
        const int slot = atomicAdd(counter, 1);
        if (gidx) { gidx[slot] = i; gpos[3 * slot] = q[0]; gpos[3 * slot + 1] = q[1]; gpos[3 * slot + 2] = q[2]; }
      }
}

// points 0..kk-1 are the sites, kk.. are ghosts
__device__ __forceinline__ void lat_point(const double *crd, const double *gpos, int kk, int q, double *p) {
  const double *s = q < kk ? crd + 3 * q : gpos + 3 * (q - kk);
  p[0] = s[0]; p[1] = s[1]; p[2] = s[2];
}
__global__ void k_lat_cell_count(const double *crd, const double *gpos, int kk, int npts, LatGrid g, int *cell_cnt) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npts) return;
  double p[3];
  lat_point(crd, gpos, kk, q, p);
  atomicAdd(cell_cnt + lat_cell(g, p), 1);
}
// exclusive scan of cell counts: single CTA, chunked (integer work; ncells is at most a few million)
__global__ void k_lat_scan(const int *cnt, int *start, int n) {
  __shared__ int warp_sum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < n ? cnt[i] : 0;
    int s = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if ((threadIdx.x & 31) >= o) s += t; }
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = threadIdx.x < (blockDim.x >> 5) ? warp_sum[threadIdx.x] : 0;
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += t; }
      warp_sum[threadIdx.x] = w;
    }
    __syncthreads();
    const int wprev = (threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0;
    if (i < n) start[i] = carry + wprev + s - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += wprev + s;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[n] = carry;
}
__global__ void k_lat_cell_fill(const double *crd, const double *gpos, int kk, int npts, LatGrid g, const int *start,
                                int *cursor, int32_t *cell_pts) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npts) return;
  double p[3];
  lat_point(crd, gpos, kk, q, p);
  const int c = lat_cell(g, p);
  cell_pts[start[c] + atomicAdd(cursor + c, 1)] = q;
}

// nncal: rows == null: count only (cnt[i] = nn(i,1) = neighbours + 1, nnmax via atomicMax); else fill rows[i + kk*k]
// (k = 0.. ascending neighbour index, 1-based site numbers), duplicates (several images of one site) removed.
__global__ void k_lat_nncal(const double *crd, const double *gpos, const int32_t *gidx, int kk, LatPbc p, LatGrid g,
                            const int *start, const int32_t *cell_pts, double ct, int *cnt, int *nnmax, int32_t *rows,
                            int rowcap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kk) return;
  const double ctm = __ddiv_rn(__dadd_rn(ct, ct), 2.0), ctsm = __dmul_rn(ctm, ctm);  // mapa, lattice.f90:2956-2973
  double pi[3] = {crd[3 * i], crd[3 * i + 1], crd[3 * i + 2]};
  int c0[3];
  for (int l = 0; l < 3; l++) c0[l] = min(max((int)floor((pi[l] - g.org[l]) * g.inv_cs), 0), g.dim[l] - 1);
  int n = 0;
  for (int dz = -1; dz <= 1; dz++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        const int cx = c0[0] + dx, cy = c0[1] + dy, cz = c0[2] + dz;
        if (cx < 0 || cy < 0 || cz < 0 || cx >= g.dim[0] || cy >= g.dim[1] || cz >= g.dim[2]) continue;
        const int c = (cz * g.dim[1] + cy) * g.dim[0] + cx;
        for (int e = start[c]; e < start[c + 1]; e++) {
          const int q = cell_pts[e];
          const int j = q < kk ? q : gidx[q - kk];
          if (j == i) continue;  // the pair loop never pairs a site with itself (or its own image)
          double d[3], r2;
          if (p.use) {
            lat_wrap_diff(p, crd, i, j, d);  // the reference tests the MINIMUM image, whichever image was binned here
            r2 = lat_r2(d);
          } else {
            r2 = 0.0;
            for (int l = 0; l < 3; l++) { d[l] = __dsub_rn(pi[l], crd[l + 3 * j]); r2 = __dadd_rn(r2, __dmul_rn(d[l], d[l])); }
          }
          if (r2 >= ctsm) continue;
          if (rows) {
            // insertion into the ascending row, skipping duplicates
            int k = n;
            bool dup = false;
            for (int t = 0; t < n; t++) if (rows[i + (size_t)kk * t] == j + 1) { dup = true; break; }
            if (dup) continue;
            while (k > 0 && rows[i + (size_t)kk * (k - 1)] > j + 1) { rows[i + (size_t)kk * k] = rows[i + (size_t)kk * (k - 1)]; k--; }
            rows[i + (size_t)kk * k] = j + 1;
            n++;
          } else {
            n++;  // may over-count duplicate images; the fill pass gives the exact count
          }
        }
      }
  if (rows) cnt[i] = n + 1;
  atomicMax(nnmax, n + 1);
  (void)rowcap;
}

// remd, part 1: set(:, t, j) for the representative of type t (1-based types, slot j = 2..cnt)
__global__ void k_lat_set(const double *crd, int kk, LatPbc p, const int32_t *iu, int ntot, const int *cnt, const int32_t *rows,
                          int nmcols, double *set) {
  const int t = blockIdx.x, la = iu[t] - 1;
  for (int j = 2 + threadIdx.x; j <= cnt[la]; j += blockDim.x) {
    const int jj = rows[la + (size_t)kk * (j - 2)] - 1;
    double d[3];
    if (p.use) lat_wrap_diff(p, crd, la, jj, d);
    else for (int l = 0; l < 3; l++) d[l] = __dsub_rn(crd[l + 3 * la], crd[l + 3 * jj]);
    for (int l = 0; l < 3; l++) set[l + 3 * ((size_t)t * nmcols + j)] = d[l];
  }
}
// remd, part 2: nn(i, 1) = imax of the representative, nn(i, k) = neighbour whose vector equals set(:, n, k)
__global__ void k_lat_remd(const double *crd, int kk, LatPbc p, const int32_t *no, const int32_t *iu, int ntot, const int *cnt,
                           const int32_t *rows, int nmcols, const double *set, int ncols, int32_t *nn, int *err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kk) return;
  const int n = no[i];
  int ino = -1;
  for (int lk = 0; lk < ntot; lk++) if (no[iu[lk] - 1] == n) { ino = iu[lk] - 1; break; }
  if (ino < 0 || n < 1 || n > ntot) { atomicMax(err, 3); return; }
  const int imax = cnt[ino];
  nn[i] = imax;
  for (int k = 2; k <= imax; k++) nn[i + (size_t)kk * (k - 1)] = 0;
  const double eps = (double)0.0001f;  // `eps = .0001`: a default-real (single precision) literal in the reference
  for (int j = 2; j <= cnt[i]; j++) {
    const int jj = rows[i + (size_t)kk * (j - 2)] - 1;
    double ret[3];
    if (p.use) lat_wrap_diff(p, crd, i, jj, ret);
    else for (int l = 0; l < 3; l++) ret[l] = __dsub_rn(crd[l + 3 * i], crd[l + 3 * jj]);
    int k = 0;
    for (int ii = 2; ii <= imax; ii++) {
      const double *s = set + 3 * ((size_t)(n - 1) * nmcols + ii);
      const double a1 = __dsub_rn(ret[0], s[0]), a2 = __dsub_rn(ret[1], s[1]), a3 = __dsub_rn(ret[2], s[2]);
      const double aaa = __dadd_rn(__dadd_rn(__dmul_rn(a1, a1), __dmul_rn(a2, a2)), __dmul_rn(a3, a3));
      if (aaa < eps) { k = ii; break; }
    }
    if (!k) { atomicMax(err, 2); return; }  // " VECTOR   NOT FOUND " -> stop in the reference
    nn[i + (size_t)kk * (k - 1)] = jj + 1;
  }
  (void)ncols;
}
