// lvo_knn.cuh — exact K-nearest-neighbour search on a uniform grid; replaces pcl::KdTreeFLANN::setInputCloud +
// nearestKSearch (reference src/laserMapping.cpp:558-559,582,648 with K = 5 and the d5^2 < 1.0 gate of :584,652;
// src/laserOdometry.cpp:386,470,640-641 with K = 1 and the d^2 < 25 gate of :389,473).
//
// Build (all problems of all lanes in one launch sequence; a "problem" is one cloud to be searched):
//   k_grid_bbox -> k_grid_setup (cell size, origin, dims, table offsets) -> k_grid_zero -> k_grid_count
//   (rank = atomicAdd per cell) -> scan (cell start table) -> k_grid_fill (points + their original index, sorted by
//   cell; x-fastest cell order so that the cells x-1..x+1 of a row are ONE contiguous range of points).
// Search: rings of cells of growing Chebyshev radius; a ring r can only contain points with |d| > (r-1) * cell along
// some axis, so the search stops as soon as the K-th best squared distance is < ((r-1) * cell)^2, and never goes beyond
// the ring that covers the gate radius.  Three forms share this rule: one warp per query (warp_knn, stand-alone
// operator and depth association), one thread per query (thread_knn, the map search: most queries, fewest instructions)
// and TW-lane tiles (tile_* primitives, the scan-to-scan association).  Squared distances use
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

#define LVO_AZ_BUCKETS 512                 // azimuth buckets of the (ring, azimuth) grid: 0.703 degrees each
#define LVO_AZ_RINGS (LVO_MAX_RINGS + 2)   // ring ids are clamped to [0, 65]

struct GridProblem {   // device-resident descriptor
  const float4* pts;   // source cloud
  const int* d_n;      // its size (device)
  float want_cell;     // requested cell size in x and y (power of two)
  float want_cell_z;   // requested cell size in z (power of two; 0 = same as want_cell)
  int mode;            // 0: uniform xyz grid; 1: (ring, azimuth) grid — cell = ring * LVO_AZ_BUCKETS + azimuth bucket
  int bbox_from;       // >= 0: this problem searches the same cloud as problem `bbox_from` (offset back from its own index is
                       //       not used; absolute index) and copies its bounding box instead of recomputing it
  int cells_cap;       // > 0: this problem's share of the cell table (the cell size doubles until the grid fits); 0 = the set's default
  float clamp_xy;      // > 0: the table only covers |x|, |y| <= clamp_xy; points outside go to the border cells (their true
                       //      distance is larger than the cell suggests, so every ring bound stays valid)
  // filled by k_grid_setup
  float cell, inv_cell;      // x / y cell size (the smallest one: all search bounds use it)
  float inv_cell_z;
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
  long long table_cap_total; // cells allocated for the whole set (sum of the per-problem shares)
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
  float cell, inv_cell, inv_cell_z; int org[3], dim[3];   // cell = the smallest cell edge (x / y); z cells may be taller
};
__device__ __forceinline__ GridView grid_view(const GridSet& g, int p) {
  const GridProblem& q = g.prob[p];
  GridView v;
  v.pts = g.sorted_pts; v.ids = g.sorted_id; v.cell_start = g.table + q.table_off;
  v.cell = q.cell; v.inv_cell = q.inv_cell; v.inv_cell_z = q.inv_cell_z;
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
  if (q.mode == 1 || q.bbox_from >= 0) return;  // fixed dimensions / box copied from the problem that shares the cloud
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
// one block: per problem (one thread each) cell size (doubling until the table fits), origin, dims; then a block
// scan gives the table / point offsets
__global__ void __launch_bounds__(256) k_grid_setup(GridSet g) {
  __shared__ unsigned sm[33];
  unsigned toff = 0, poff = 0;
  for (int base = 0; base < g.nprob; base += blockDim.x) {
    const int p = base + threadIdx.x;
    unsigned my_cells = 0, my_pts = 0;
    if (p < g.nprob) {
      GridProblem& q = g.prob[p];
      if (q.bbox_from >= 0) for (int c = 0; c < 3; ++c) { q.bb_mn[c] = g.prob[q.bbox_from].bb_mn[c]; q.bb_mx[c] = g.prob[q.bbox_from].bb_mx[c]; }
      float cell = q.want_cell, cellz = q.want_cell_z > 0.f ? q.want_cell_z : q.want_cell;
      int org[3] = {0, 0, 0}, dim[3] = {0, 0, 0};
      long long nc = 0;
      if (q.mode == 1) {
        org[0] = org[1] = org[2] = 0; dim[0] = LVO_AZ_BUCKETS; dim[1] = LVO_AZ_RINGS; dim[2] = 1;
        nc = (long long)LVO_AZ_BUCKETS * LVO_AZ_RINGS;
      } else if (q.n > 0) {
        for (int it = 0; it < 40; ++it) {
          nc = 1;
          for (int c = 0; c < 3; ++c) {
            const float inv = 1.0f / (c == 2 ? cellz : cell);
            float fmn = ord2f_k(q.bb_mn[c]), fmx = ord2f_k(q.bb_mx[c]);
            if (q.clamp_xy > 0.f && c < 2) { fmn = fminf(fmaxf(fmn, -q.clamp_xy), q.clamp_xy); fmx = fminf(fmaxf(fmx, -q.clamp_xy), q.clamp_xy); }
            const int lo = cell_coord(fmn, inv), hi = cell_coord(fmx, inv);
            org[c] = lo; dim[c] = hi - lo + 1;
            nc *= (long long)dim[c];
          }
          if (nc <= (long long)(q.cells_cap > 0 ? q.cells_cap : g.cells_cap_per_problem)) break;
          cell *= 2.0f;
          if (cellz < cell) cellz = cell;
        }
      }
      q.cell = cell; q.inv_cell = 1.0f / cell; q.inv_cell_z = 1.0f / cellz;
      for (int c = 0; c < 3; ++c) { q.org[c] = org[c]; q.dim[c] = dim[c]; }
      q.ncells = (int)nc;
      my_cells = (unsigned)nc; my_pts = (unsigned)q.n;
    }
    unsigned tc, tp;
    const unsigned ec = block_excl_scan(my_cells, sm, &tc);
    const unsigned ep = block_excl_scan(my_pts, sm, &tp);
    if (p < g.nprob) { g.prob[p].table_off = toff + ec; g.pt_off[p] = poff + ep; }
    toff += tc; poff += tp;
  }
  if (threadIdx.x == 0) { g.pt_off[g.nprob] = poff; *g.d_table_len = (int)toff + 1; }
}
__global__ void k_grid_zero(GridSet g) {
  const int len = *g.d_table_len;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) g.table[i] = 0;
}
// azimuth bucket of a direction; any consistent monotone function of the angle works (the searches add margins)
__device__ __forceinline__ int az_bucket(float x, float y) {
  const float phi = atan2f(y, x) + 3.14159274f;
  int b = (int)floorf(phi * ((float)LVO_AZ_BUCKETS / 6.28318548f));
  return min(max(b, 0), LVO_AZ_BUCKETS - 1);
}
__device__ __forceinline__ int ring_clamped(float intensity) { return min(max(int(intensity), 0), LVO_AZ_RINGS - 1); }
__device__ __forceinline__ int grid_cell_of(const GridProblem& q, float4 v) {
  if (q.mode == 1) return ring_clamped(v.w) * LVO_AZ_BUCKETS + az_bucket(v.x, v.y);
  int cx = cell_coord(v.x, q.inv_cell) - q.org[0], cy = cell_coord(v.y, q.inv_cell) - q.org[1], cz = cell_coord(v.z, q.inv_cell_z) - q.org[2];
  cx = min(max(cx, 0), q.dim[0] - 1); cy = min(max(cy, 0), q.dim[1] - 1); cz = min(max(cz, 0), q.dim[2] - 1);
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
    g.sorted_pts[dst] = make_float4(v.x, v.y, v.z, __int_as_float(i));  // w carries the original index: one 16-byte load per candidate
  }
}

static inline void lvo_grid_build(cudaStream_t st, const GridSet& g, long long* launches) {
  // grid-stride kernels: a few thousand threads per problem are enough; capacity-sized grids would launch ~300 k idle blocks
  const int gx = max(1, min(lvo_div_up(g.pts_cap_per_problem, 256), max(24, 2368 / g.nprob)));
  dim3 gp(gx, g.nprob);
  k_grid_reset<<<lvo_div_up(g.nprob, 64), 64, 0, st>>>(g);
  k_grid_bbox<<<gp, 256, 0, st>>>(g);
  k_grid_setup<<<1, 256, 0, st>>>(g);
  k_grid_zero<<<1184, 256, 0, st>>>(g);
  k_grid_count<<<gp, 256, 0, st>>>(g);
  const long long table_cap = (g.table_cap_total > 0 ? g.table_cap_total : (long long)g.nprob * g.cells_cap_per_problem) + 1;
  lvo_scan_exclusive(st, g.table, g.d_table_len, (int)(table_cap < 0x7fffffffLL ? table_cap : 0x7fffffffLL), nullptr, g.scan, launches);
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
  // empty list that only accepts candidates with d < bound, or d == bound (they rank before the INT_MAX placeholders): for searches whose
  // result is void anyway unless K candidates lie inside `bound`, most candidates of a cell block are then rejected by ONE comparison
  __device__ __forceinline__ void init(float bound) {
#pragma unroll
    for (int k = 0; k < K; ++k) { d[k] = bound; id[k] = INT_MAX; }
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
    tk.insert(sqdist3(p, qx, qy, qz), __float_as_int(p.w));
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
  const int cx = cell_coord(qx, g.inv_cell) - g.org[0], cy = cell_coord(qy, g.inv_cell) - g.org[1], cz = cell_coord(qz, g.inv_cell_z) - g.org[2];
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

// ---------------------------------------------------------------------------------------------------------------
// thread-per-query exact K-NN (throughput form): every thread walks the 27 (or more) cells of its own query.
// Per query this issues ~9x fewer warp instructions than the warp-cooperative form (no cross-lane merge, full lane
// utilisation); neighbouring threads hold neighbouring queries (the stack clouds are voxel-sorted), so their
// candidate rows share L1 lines.  Same ring-expansion rule and the same (distance, index) ranking => same result.
// ---------------------------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void thread_scan_row(const GridView& g, int cz, int cy, int x0, int x1, float qx, float qy, float qz, TopK<K>& tk) {
  if (cz < 0 || cz >= g.dim[2] || cy < 0 || cy >= g.dim[1]) return;
  x0 = max(x0, 0); x1 = min(x1, g.dim[0] - 1);
  if (x0 > x1) return;
  const int rowbase = (cz * g.dim[1] + cy) * g.dim[0];
  const unsigned b = __ldg(g.cell_start + rowbase + x0), e = __ldg(g.cell_start + rowbase + x1 + 1);
  for (unsigned t = b; t < e; ++t) {
    const float4 p = __ldg(g.pts + t);
    tk.insert(sqdist3(p, qx, qy, qz), __float_as_int(p.w));
  }
}
template <int K>
__device__ __forceinline__ bool thread_knn(const GridView& g, float qx, float qy, float qz, float max_sq, TopK<K>& tk) {
  tk.init(max_sq);   // a candidate at or beyond the gate can only end in a void result (the test at the end), so it need not be ranked
  if (g.dim[0] <= 0) return false;
  const int cx = cell_coord(qx, g.inv_cell) - g.org[0], cy = cell_coord(qy, g.inv_cell) - g.org[1], cz = cell_coord(qz, g.inv_cell_z) - g.org[2];
  int R = (int)ceilf(sqrtf(max_sq) * g.inv_cell);
  if (R < 1) R = 1;
  // rings 0 and 1: fetch the 9 row ranges first (independent loads), then stream the candidates
  unsigned rb[9], re[9];
  const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int z = cz + k / 3 - 1, y = cy + k % 3 - 1;
    rb[k] = re[k] = 0;
    if (z >= 0 && z < g.dim[2] && y >= 0 && y < g.dim[1] && x0 <= x1) {
      const int rowbase = (z * g.dim[1] + y) * g.dim[0];
      rb[k] = __ldg(g.cell_start + rowbase + x0); re[k] = __ldg(g.cell_start + rowbase + x1 + 1);
    }
  }
  // candidates in batches of four: four independent 16-byte loads in flight per thread instead of one; insertion order is unchanged,
  // hence the same result.  ONE copy of the row body (unroll 1): nine inlined copies made the kernel 7200 SASS instructions (115 KB) and
  // instruction-fetch bound (ncu: stalled_no_instruction 20 of 27 warp-cycles per issue, profiles/r2a_map_knn_fit_stalls.txt).
#pragma unroll 1
  for (int k = 0; k < 9; ++k) {
    unsigned t = rb[k];
    const unsigned e = re[k];
    for (; t + 4 <= e; t += 4) {
      const float4 p0 = __ldg(g.pts + t), p1 = __ldg(g.pts + t + 1), p2 = __ldg(g.pts + t + 2), p3 = __ldg(g.pts + t + 3);
      tk.insert(sqdist3(p0, qx, qy, qz), __float_as_int(p0.w));
      tk.insert(sqdist3(p1, qx, qy, qz), __float_as_int(p1.w));
      tk.insert(sqdist3(p2, qx, qy, qz), __float_as_int(p2.w));
      tk.insert(sqdist3(p3, qx, qy, qz), __float_as_int(p3.w));
    }
    if (t < e) {   // 1..3 left: load them together as well
      const float4 p0 = __ldg(g.pts + t);
      const float4 p1 = t + 1 < e ? __ldg(g.pts + t + 1) : p0;
      const float4 p2 = t + 2 < e ? __ldg(g.pts + t + 2) : p0;
      tk.insert(sqdist3(p0, qx, qy, qz), __float_as_int(p0.w));
      if (t + 1 < e) tk.insert(sqdist3(p1, qx, qy, qz), __float_as_int(p1.w));
      if (t + 2 < e) tk.insert(sqdist3(p2, qx, qy, qz), __float_as_int(p2.w));
    }
  }
  for (int r = 2; r <= R; ++r) {
    const float bound = (float)(r - 1) * g.cell;
    if (tk.d[K - 1] < bound * bound) break;
    for (int dz = -r; dz <= r; ++dz)
      for (int dy = -r; dy <= r; ++dy) {
        if (dz == -r || dz == r || dy == -r || dy == r) thread_scan_row<K>(g, cz + dz, cy + dy, cx - r, cx + r, qx, qy, qz, tk);
        else { thread_scan_row<K>(g, cz + dz, cy + dy, cx - r, cx - r, qx, qy, qz, tk); thread_scan_row<K>(g, cz + dz, cy + dy, cx + r, cx + r, qx, qy, qz, tk); }
      }
  }
  return tk.id[K - 1] != INT_MAX && (double)tk.d[K - 1] < (double)max_sq;
}

// ---------------------------------------------------------------------------------------------------------------
// Tiles: TW lanes per query.  Every primitive below is executed by all 32 lanes (full-mask shuffles with width TW); a
// tile that has nothing to do passes empty ranges.  This shares all per-query overhead (bounds, prefix
// sums, reductions, control flow) between four queries, which is what the latency-bound association kernels need.
// ---------------------------------------------------------------------------------------------------------------
// TW = lanes per query (8: four queries per warp; 32: one query per warp, for the few expensive ones).
template <int TW> __device__ __forceinline__ unsigned tile_lane() { return threadIdx.x & (TW - 1); }
template <int TW>
__device__ __forceinline__ unsigned tile_incl_scan(unsigned v) {
  const unsigned tl = tile_lane<TW>();
#pragma unroll
  for (int o = 1; o < TW; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, v, o, TW);
    if (tl >= (unsigned)o) v += t;
  }
  return v;
}
// Lane j < NR of the tile holds range [b, e) (possibly empty); f(p, r): candidate point p (w = original index) of range r.
// Software-pipelined: the candidate of step k+1 is fetched before step k is processed, so two loads per lane are in flight.
template <int TW, int NR, class F>
__device__ __forceinline__ void tile_scan_ranges(const float4* __restrict__ pts, unsigned b, unsigned e, F&& f) {
  const unsigned tl = tile_lane<TW>();
  const unsigned len = e > b ? e - b : 0u;
  const unsigned incl = tile_incl_scan<TW>(len);
  const unsigned total = __shfl_sync(0xffffffffu, incl, TW - 1, TW);
  auto locate = [&](unsigned k, int& r) -> unsigned {
    r = 0;
#pragma unroll
    for (int j = 0; j < NR - 1; ++j) r += (k >= __shfl_sync(0xffffffffu, incl, j, TW)) ? 1 : 0;
    const unsigned rb = __shfl_sync(0xffffffffu, b, r, TW), ri = __shfl_sync(0xffffffffu, incl, r, TW), rl = __shfl_sync(0xffffffffu, len, r, TW);
    return rb + (k - (ri - rl));
  };
  if (!__any_sync(0xffffffffu, total > 0)) return;
  int r_cur, r_nxt = 0;
  unsigned t = locate(tl, r_cur);
  bool v_cur = tl < total;
  float4 p_cur = make_float4(0.f, 0.f, 0.f, 0.f), p_nxt = p_cur;
  if (v_cur) p_cur = __ldg(pts + t);
  for (unsigned k0 = 0; __any_sync(0xffffffffu, k0 < total); k0 += TW) {
    const unsigned kn = k0 + TW + tl;
    const bool v_nxt = kn < total;
    if (__any_sync(0xffffffffu, k0 + TW < total)) {
      const unsigned tn = locate(kn, r_nxt);
      if (v_nxt) p_nxt = __ldg(pts + tn);
    }
    if (v_cur) f(p_cur, r_cur);
    p_cur = p_nxt; r_cur = r_nxt; v_cur = v_nxt;
  }
}
// (d, key) lexicographic minimum over the tile, payload j
template <int TW>
__device__ __forceinline__ void tile_min3(float& d, int& key, int& j) {
#pragma unroll
  for (int o = TW / 2; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, d, o, TW);
    const int ok = __shfl_xor_sync(0xffffffffu, key, o, TW), oj = __shfl_xor_sync(0xffffffffu, j, o, TW);
    if (od < d || (od == d && ok < key)) { d = od; key = ok; j = oj; }
  }
}

// bounds of the x-row (cz, cy, x0..x1) of a grid, empty when outside
__device__ __forceinline__ void row_bounds(const GridView& g, int cz, int cy, int x0, int x1, unsigned& b, unsigned& e) {
  b = e = 0;
  if (cz < 0 || cz >= g.dim[2] || cy < 0 || cy >= g.dim[1]) return;
  x0 = max(x0, 0); x1 = min(x1, g.dim[0] - 1);
  if (x0 > x1) return;
  const int rowbase = (cz * g.dim[1] + cy) * g.dim[0];
  b = __ldg(g.cell_start + rowbase + x0); e = __ldg(g.cell_start + rowbase + x1 + 1);
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

// Throughput form: S independent (map, query set) problems in ONE launch, blockIdx.y = problem.
// One thread per query; q_off[p] is the offset of problem p's queries in the concatenated query / output arrays.
__global__ void __launch_bounds__(128) k_knn5_batch(GridSet gs, const float4* __restrict__ q, const unsigned* __restrict__ q_off, float max_sq,
                                                    int* __restrict__ ind, float* __restrict__ sq) {
  const int p = blockIdx.y;
  const GridView g = grid_view(gs, p);
  const unsigned b = q_off[p], n = q_off[p + 1] - b;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 v = __ldg(q + b + i);
    TopK<5> tk;
    const bool ok = thread_knn<5>(g, v.x, v.y, v.z, max_sq, tk);
    int* o = ind + (size_t)(b + i) * 5;
    float* d = sq + (size_t)(b + i) * 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) { o[k] = ok ? tk.id[k] : -1; d[k] = ok ? tk.d[k] : INFINITY; }
  }
}
