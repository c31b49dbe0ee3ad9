// lvo_knn.cuh — exact K-nearest-neighbour search on a uniform grid; replaces pcl::KdTreeFLANN::setInputCloud +
// nearestKSearch (reference src/laserMapping.cpp:558-559,582,648 with K = 5 and the d5^2 < 1.0 gate of :584,652;
// src/laserOdometry.cpp:386,470,640-641 with K = 1 and the d^2 < 25 gate of :389,473).
//
// Build (all problems of all lanes in one launch sequence; a "problem" is one cloud to be searched):
//   k_grid_bbox -> k_grid_setup (cell size, origin, dims, table offsets) -> k_grid_zero -> k_grid_count
//   (rank = atomicAdd per cell) -> scan (cell start table) -> k_grid_fill (points + their original index, sorted by
//   cell; x-fastest cell order so that the cells x-1..x+1 of a row are ONE contiguous range of points).
// Search (device function, one warp per query): rings of cells of growing Chebyshev radius; a ring r can only
// contain points with |d| > (r-1) * cell along some axis, so the search stops as soon as the K-th best squared
// distance is < (r * cell)^2, and never goes beyond the ring that covers the gate radius.  Squared distances use
// FLANN's L2_Simple accumulation ((dx*dx + dy*dy) + dz*dz) in float without FMA; candidates are ranked by
// (distance, original index), which is the oracle's fixed tie-break, so the result does not depend on the order
// of the points inside a cell (the atomics above are free to race).
// Cell sizes are powers of two (>= the requested size), so floor(x / cell) is exact and the ring bound is safe in
// floating point (rounding is monotone).
#pragma once
#include "lvo_internal.h"
#include "lvo_prims.cuh"
#include <float.h>
#include <limits.h>

struct GridProblem {   // device-resident descriptor
  const float4* pts;   // source cloud
  const int* d_n;      // its size (device)
  float want_cell;     // requested cell size (power of two)
  // filled by k_grid_setup
  float cell, inv_cell;
  int org[3], dim[3];
  int ncells;
  unsigned table_off;  // offset of this problem's cells in the shared table
  int n;               // snapshot of *d_n
  int bb_mn[3], bb_mx[3];  // ordered-int encoded bbox
};

struct GridSet {
  GridProblem* prob;        // [nprob]
  int nprob;
  int cells_cap_per_problem;
  unsigned* table;          // [nprob * cells_cap + 1] counts -> starts
  int* d_table_len;         // total cells + 1
  int* rank;                // [pts_cap_total] rank of a point inside its cell, laid out per problem at pt_off
  unsigned* pt_off;         // [nprob + 1] offsets of each problem's points in the sorted arrays
  float4* sorted_pts;       // [pts_cap_total]
  int* sorted_id;           // [pts_cap_total]
  int pts_cap_per_problem;  // bound for grid sizing
  LvoScanScratch scan;
};

struct GridView {  // what a search needs
  const float4* pts; const int* ids; const unsigned* cell_start;
  float cell, inv_cell; int org[3], dim[3];
};
__device__ __forceinline__ GridView grid_view(const GridSet& g, int p) {
  const GridProblem& q = g.prob[p];
  GridView v;
  v.pts = g.sorted_pts; v.ids = g.sorted_id; v.cell_start = g.table + q.table_off;
  v.cell = q.cell; v.inv_cell = q.inv_cell;
  for (int c = 0; c < 3; ++c) { v.org[c] = q.org[c]; v.dim[c] = q.dim[c]; }
  return v;
}

__device__ __forceinline__ int f2ord_k(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f_k(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ int cell_coord(float x, float inv_cell) {
  float f = floorf(x * inv_cell);
  f = fminf(fmaxf(f, -1.0e8f), 1.0e8f);
  return (int)f;
}

__global__ void k_grid_reset(GridSet g) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < g.nprob) {
    for (int c = 0; c < 3; ++c) { g.prob[p].bb_mn[c] = INT_MAX; g.prob[p].bb_mx[c] = INT_MIN; }
    g.prob[p].n = *g.prob[p].d_n;
  }
}
__global__ void k_grid_bbox(GridSet g) {
  const int p = blockIdx.y;
  GridProblem& q = g.prob[p];
  const int n = *q.d_n;
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 v = q.pts[i];
    const int a = f2ord_k(v.x), b = f2ord_k(v.y), c = f2ord_k(v.z);
    mn[0] = min(mn[0], a); mx[0] = max(mx[0], a);
    mn[1] = min(mn[1], b); mx[1] = max(mx[1], b);
    mn[2] = min(mn[2], c); mx[2] = max(mx[2], c);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) { mn[c] = __reduce_min_sync(0xffffffffu, mn[c]); mx[c] = __reduce_max_sync(0xffffffffu, mx[c]); }
  if ((threadIdx.x & 31) == 0 && mn[0] != INT_MAX)
    for (int c = 0; c < 3; ++c) { atomicMin(&q.bb_mn[c], mn[c]); atomicMax(&q.bb_mx[c], mx[c]); }
}
// one thread: cell size (doubling until the table fits), origin, dims, offsets
__global__ void k_grid_setup(GridSet g) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned toff = 0, poff = 0;
  for (int p = 0; p < g.nprob; ++p) {
    GridProblem& q = g.prob[p];
    q.table_off = toff;
    g.pt_off[p] = poff;
    float cell = q.want_cell;
    int org[3] = {0, 0, 0}, dim[3] = {0, 0, 0};
    long long nc = 0;
    if (q.n > 0) {
      for (int it = 0; it < 40; ++it) {
        const float inv = 1.0f / cell;
        nc = 1;
        for (int c = 0; c < 3; ++c) {
          const int lo = cell_coord(ord2f_k(q.bb_mn[c]), inv), hi = cell_coord(ord2f_k(q.bb_mx[c]), inv);
          org[c] = lo; dim[c] = hi - lo + 1;
          nc *= (long long)dim[c];
        }
        if (nc <= (long long)g.cells_cap_per_problem) break;
        cell *= 2.0f;
      }
    }
    q.cell = cell; q.inv_cell = 1.0f / cell;
    for (int c = 0; c < 3; ++c) { q.org[c] = org[c]; q.dim[c] = dim[c]; }
    q.ncells = (int)nc;
    toff += (unsigned)nc;
    poff += (unsigned)q.n;
  }
  g.pt_off[g.nprob] = poff;
  *g.d_table_len = (int)toff + 1;
}
__global__ void k_grid_zero(GridSet g) {
  const int len = *g.d_table_len;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) g.table[i] = 0;
}
__device__ __forceinline__ int grid_cell_of(const GridProblem& q, float4 v) {
  const int cx = cell_coord(v.x, q.inv_cell) - q.org[0], cy = cell_coord(v.y, q.inv_cell) - q.org[1], cz = cell_coord(v.z, q.inv_cell) - q.org[2];
  return (cz * q.dim[1] + cy) * q.dim[0] + cx;
}
__global__ void k_grid_count(GridSet g) {
  const int p = blockIdx.y;
  const GridProblem& q = g.prob[p];
  const int n = q.n;
  const unsigned po = g.pt_off[p];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    g.rank[po + i] = (int)atomicAdd(&g.table[q.table_off + grid_cell_of(q, q.pts[i])], 1u);
}
__global__ void k_grid_fill(GridSet g) {
  const int p = blockIdx.y;
  const GridProblem& q = g.prob[p];
  const int n = q.n;
  const unsigned po = g.pt_off[p];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 v = q.pts[i];
    const unsigned dst = g.table[q.table_off + grid_cell_of(q, v)] + (unsigned)g.rank[po + i];
    g.sorted_pts[dst] = v;
    g.sorted_id[dst] = i;
  }
}

static inline void lvo_grid_build(cudaStream_t st, const GridSet& g, long long* launches) {
  const int gx = max(1, min(lvo_div_up(g.pts_cap_per_problem, 256), 296));
  dim3 gp(gx, g.nprob);
  k_grid_reset<<<lvo_div_up(g.nprob, 64), 64, 0, st>>>(g);
  k_grid_bbox<<<gp, 256, 0, st>>>(g);
  k_grid_setup<<<1, 32, 0, st>>>(g);
  k_grid_zero<<<1184, 256, 0, st>>>(g);
  k_grid_count<<<gp, 256, 0, st>>>(g);
  lvo_scan_exclusive(st, g.table, g.d_table_len, g.nprob * g.cells_cap_per_problem + 1, nullptr, g.scan, launches);
  k_grid_fill<<<gp, 256, 0, st>>>(g);
  if (launches) *launches += 6;
}

// ---------------------------------------------------------------------------------------------------------------
// warp-cooperative exact K-NN
// ---------------------------------------------------------------------------------------------------------------
template <int K>
struct TopK {
  float d[K]; int id[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int k = 0; k < K; ++k) { d[k] = FLT_MAX; id[k] = INT_MAX; }
  }
  __device__ __forceinline__ static bool less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }
  __device__ __forceinline__ void insert(float dd, int ii) {
    if (!less(dd, ii, d[K - 1], id[K - 1])) return;
    d[K - 1] = dd; id[K - 1] = ii;
#pragma unroll
    for (int k = K - 1; k > 0; --k) {
      if (less(d[k], id[k], d[k - 1], id[k - 1])) {
        float td = d[k]; d[k] = d[k - 1]; d[k - 1] = td;
        int ti = id[k]; id[k] = id[k - 1]; id[k - 1] = ti;
      }
    }
  }
};

// scan the points of cells [c0, c1] of one x-row (contiguous), lanes strided
template <int K>
__device__ __forceinline__ void scan_row(const GridView& g, int cz, int cy, int x0, int x1, float qx, float qy, float qz, TopK<K>& tk, unsigned ln) {
  if (cz < 0 || cz >= g.dim[2] || cy < 0 || cy >= g.dim[1]) return;
  x0 = max(x0, 0); x1 = min(x1, g.dim[0] - 1);
  if (x0 > x1) return;
  const int rowbase = (cz * g.dim[1] + cy) * g.dim[0];
  const unsigned b = g.cell_start[rowbase + x0], e = g.cell_start[rowbase + x1 + 1];
  for (unsigned t = b + ln; t < e; t += 32) {
    const float4 p = g.pts[t];
    tk.insert(sqdist3(p, qx, qy, qz), g.ids[t]);
  }
}

// Merge the per-lane candidate lists; afterwards every lane holds the warp-wide best K, sorted by (d, id).
template <int K>
__device__ __forceinline__ void warp_merge(TopK<K>& tk) {
  float rd[K]; int ri[K];
  int ptr = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float cd = FLT_MAX; int ci = INT_MAX;
#pragma unroll
    for (int j = 0; j < K; ++j) if (j == ptr) { cd = tk.d[j]; ci = tk.id[j]; }
    float bd = cd; int bi = ci;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (TopK<K>::less(od, oi, bd, bi)) { bd = od; bi = oi; }
    }
    rd[k] = bd; ri[k] = bi;
    if (bi == ci && bd == cd && ci != INT_MAX) ptr++;  // ids are unique, so exactly one lane advances
  }
#pragma unroll
  for (int k = 0; k < K; ++k) { tk.d[k] = rd[k]; tk.id[k] = ri[k]; }
}

// Exact K nearest points of (qx,qy,qz) with d^2 < max_sq.  Returns true (in all lanes) iff K such points exist;
// then tk holds them sorted by (d, id).  Must be called by all 32 lanes.
template <int K>
__device__ __forceinline__ bool warp_knn(const GridView& g, float qx, float qy, float qz, float max_sq, TopK<K>& tk) {
  const unsigned ln = threadIdx.x & 31;
  tk.init();
  if (g.dim[0] <= 0) return false;
  const int cx = cell_coord(qx, g.inv_cell) - g.org[0], cy = cell_coord(qy, g.inv_cell) - g.org[1], cz = cell_coord(qz, g.inv_cell) - g.org[2];
  // number of rings that can hold a point inside the gate: points of ring r have d > (r-1)*cell
  int R = (int)ceilf(sqrtf(max_sq) * g.inv_cell);
  if (R < 1) R = 1;
  // rings 0 and 1 together: 9 rows of 3 cells
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy) scan_row<K>(g, cz + dz, cy + dy, cx - 1, cx + 1, qx, qy, qz, tk, ln);
  warp_merge<K>(tk);
  for (int r = 2; r <= R; ++r) {
    const float bound = (float)(r - 1) * g.cell;
    if (tk.d[K - 1] < bound * bound) break;  // nothing outside ring r-1 can beat or tie the K-th best
    for (int dz = -r; dz <= r; ++dz)
      for (int dy = -r; dy <= r; ++dy) {
        if (dz == -r || dz == r || dy == -r || dy == r) scan_row<K>(g, cz + dz, cy + dy, cx - r, cx + r, qx, qy, qz, tk, ln);
        else { scan_row<K>(g, cz + dz, cy + dy, cx - r, cx - r, qx, qy, qz, tk, ln); scan_row<K>(g, cz + dz, cy + dy, cx + r, cx + r, qx, qy, qz, tk, ln); }
      }
    warp_merge<K>(tk);
  }
  return tk.id[K - 1] != INT_MAX && (double)tk.d[K - 1] < (double)max_sq;
}

// Stand-alone operator (lvo_knn): one warp per query, problem 0 of the grid set.
template <int K>
__global__ void k_knn_queries(GridSet gs, const float4* __restrict__ q, int nq, float max_sq, int* __restrict__ ind, float* __restrict__ sq) {
  const GridView g = grid_view(gs, 0);
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const unsigned ln = threadIdx.x & 31;
  for (int i = wid; i < nq; i += nw) {
    const float4 p = q[i];
    TopK<K> tk;
    const bool ok = warp_knn<K>(g, p.x, p.y, p.z, max_sq, tk);
    if (ln == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) { ind[i * K + k] = ok ? tk.id[k] : -1; sq[i * K + k] = ok ? tk.d[k] : INFINITY; }
    }
  }
}
