// lvo_knn_tile.cuh — the 5-NN map search of lvo_scan_to_map with the candidate cells STAGED IN SHARED MEMORY
// (reference src/laserMapping.cpp:582-584 / :648-652: kdtree->nearestKSearch(pointSel, 5, ...) + the d5^2 < 1.0 gate).
//
// The thread-per-query search of lvo_knn.cuh (thread_knn) walks the 27 cells of every query with dependent 16-byte loads from
// L2: ncu showed it latency-bound (long scoreboard 9 warps per issue, issue slots 29 % busy).  Here:
//
//   k_map_qsort      once per frame and (lane, type): the queries (the voxel-filtered stack) sorted by the map-grid cell of their
//                    position under the frame's initial pose — register bitonic sort of (cell key, query index), up to 8192
//                    queries per CTA (larger stacks keep their order: the search below is exact for ANY order, only slower).
//   k_map_knn_tile   a warp takes 32 consecutive queries of that order.  Every outer iteration it groups them by the (z, y) cell row of
//                    their CURRENT position (queries that drifted across a cell border simply form their own small group), and per
//                    group stages the candidate points into the warp's shared-memory buffer: the cell-start entries of the 9
//                    neighbouring rows, then 9 bulk asynchronous copies (cp.async.bulk, completion on an mbarrier) of the point
//                    ranges x-1 .. x+1 of the whole x-run of the group — rows of the cell-sorted cloud are contiguous, so every
//                    copy is one coalesced stream of float4.  Each lane then scans the 27 cells of its own query out of shared
//                    memory.  The copies of one warp overlap the scans of the other warps of the SM.
//
// Exactness: the same candidate cells, the same float distance expression (FLANN L2_Simple order, no FMA) and the same
// (distance, original index) ranking as thread_knn, hence the same index sets bit for bit.
#pragma once
#include "lvo_knn.cuh"

#define LVO_KT_WARPS 8          // warps per CTA
#define LVO_KT_CAP 512          // candidate points a warp can stage at once (8 KB)
#define LVO_KT_RUN 60           // cells of an x-run; the staged range is the run plus one cell on either side
#define LVO_KT_CS (LVO_KT_RUN + 4)
#define LVO_QSORT_MAX 8192      // queries per (lane, type) that k_map_qsort orders
#define LVO_KT_BULK_MIN 64      // rows of at least this many points (1 KB) go through cp.async.bulk; shorter ones are loaded by the lanes
                                // (a bulk copy costs ~0.2 us of the SM's copy unit whatever its size: measured while tuning this kernel)
#define LVO_KT_MAX_GROUPS 5     // a chunk whose 32 queries fall into more cell rows than this is searched thread-per-query instead

// ---- mbarrier / bulk-copy PTX (sm_90+; SASS: SYNCS.* and UBLKCP) ---------------------------------------------------------------
__device__ __forceinline__ unsigned kt_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kt_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void kt_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void kt_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(kt_smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(kt_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool kt_mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(kt_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

struct __align__(16) KtWarpSmem {
  float4 pts[LVO_KT_CAP];            // staged candidates, row after row
  unsigned cs[9][LVO_KT_CS];         // cell-start entries of the 9 rows for x in [lo, hi + 1]
  unsigned long long bar;            // completion barrier of this warp's bulk copies
  unsigned long long pad;
};

// The search problem of one (lane, type) as the tiled kernels see it.
struct KtProblem {
  GridView g;
  const float4* queries;   // stack points (sensor frame)
  const int* order;        // sorted query order, or null = identity
  int n;
  int* out;                // [n][5] index sets (-1 rows when d5^2 >= max_sq)
};

// One warp: exact 5-NN of up to 32 queries (one per lane; `have` false = no query in this lane).  sel = the query point in the
// map frame.  Results into tk (valid == true iff 5 neighbours lie inside the gate).  All 32 lanes must call it together.
__device__ __forceinline__ void kt_warp_search(const GridView& g, KtWarpSmem& sm, unsigned& phase, bool have, float qx, float qy, float qz, float max_sq,
                                               TopK<5>& tk, bool& valid) {
  const unsigned ln = threadIdx.x & 31;
  tk.init();
  valid = false;
  int cx = 0, cy = 0, cz = 0;
  if (have) {
    cx = cell_coord(qx, g.inv_cell) - g.org[0]; cy = cell_coord(qy, g.inv_cell) - g.org[1]; cz = cell_coord(qz, g.inv_cell_z) - g.org[2];
    // no cell of the 3 x 3 x 3 block inside the grid: nothing to search
    if (g.dim[0] <= 0 || cx < -1 || cx > g.dim[0] || cy < -1 || cy > g.dim[1] || cz < -1 || cz > g.dim[2]) have = false;
  }
  const int rowkey = have ? (cz + 1) * (g.dim[1] + 2) + (cy + 1) : -1;
  unsigned todo = __ballot_sync(0xffffffffu, have);
  {
    // Scattered queries (far-field arcs of the sweep: a few queries per cell row) would be staged one or two at a time, each group paying
    // two dependent L2 round trips; there the 32 lanes walking their own cells concurrently are faster.
    const unsigned same = __match_any_sync(0xffffffffu, rowkey);
    const unsigned heads = __ballot_sync(0xffffffffu, have && (int)ln == __ffs(same) - 1);
    if (__popc(heads) > LVO_KT_MAX_GROUPS) {
      if (have) valid = thread_knn<5>(g, qx, qy, qz, max_sq, tk);
      return;
    }
  }
  while (todo) {
    // ---- the group: lanes whose query lies in the cell row of the first lane still to do, x-run capped at LVO_KT_RUN cells
    const int leader = __ffs(todo) - 1;
    const int lk = __shfl_sync(0xffffffffu, rowkey, leader);
    bool in = ((todo >> ln) & 1u) && rowkey == lk;
    const int xa = __reduce_min_sync(0xffffffffu, in ? cx : INT_MAX);
    in = in && cx <= xa + LVO_KT_RUN - 1;
    const int xb = __reduce_max_sync(0xffffffffu, in ? cx : INT_MIN);
    const int gy = lk % (g.dim[1] + 2) - 1, gz = lk / (g.dim[1] + 2) - 1;   // the group's row
    const int lo = max(xa - 1, 0), hi = min(xb + 1, g.dim[0] - 1);          // staged cells (hi < lo: the run lies outside the grid)
    const int ncs = hi - lo + 2;                                             // cell-start entries per row
    // ---- cell-start slices of the 9 rows (rows outside the grid are empty)
    if (hi >= lo) {
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const int z = gz + r / 3 - 1, y = gy + r % 3 - 1;
        const bool row_ok = z >= 0 && z < g.dim[2] && y >= 0 && y < g.dim[1];
        const unsigned* src = g.cell_start + (z * g.dim[1] + y) * g.dim[0] + lo;
        for (int x = (int)ln; x < ncs; x += 32) sm.cs[r][x] = row_ok ? __ldg(src + x) : 0u;
      }
    }
    __syncwarp();
    // ---- how much of the run fits the buffer: points of all 9 rows up to the right edge of each lane's own block
    int jl = 0, jr = 0;   // my block inside the staged range: entries [jl, jr] of cs
    unsigned need = 0;
    if (in && hi >= lo) {
      jl = max(cx - 1, 0) - lo; jr = min(cx + 1, g.dim[0] - 1) + 1 - lo;
#pragma unroll
      for (int r = 0; r < 9; ++r) need += sm.cs[r][jr] - sm.cs[r][0];
    }
    const bool fits = in && need <= LVO_KT_CAP;
    const unsigned fit_mask = __ballot_sync(0xffffffffu, fits);
    const unsigned in_mask = __ballot_sync(0xffffffffu, in);
    if (hi < lo) { todo &= ~in_mask; continue; }   // the whole run is outside the grid in x: no candidates
    if (fit_mask == 0) {
      // not even the left-most block fits (cells with hundreds of points): that query scans global memory directly
      const int xl = __reduce_min_sync(0xffffffffu, in ? cx : INT_MAX);
      const bool me = in && cx == xl;
      if (me) {
#pragma unroll
        for (int r = 0; r < 9; ++r) {
          const int z = gz + r / 3 - 1, y = gy + r % 3 - 1;
          if (z < 0 || z >= g.dim[2] || y < 0 || y >= g.dim[1]) continue;
          for (unsigned t = sm.cs[r][jl]; t < sm.cs[r][jr]; ++t) {
            const float4 p = __ldg(g.pts + t);
            tk.insert(sqdist3(p, qx, qy, qz), __float_as_int(p.w));
          }
        }
      }
      todo &= ~__ballot_sync(0xffffffffu, me);
      __syncwarp();
      continue;
    }
    // the lanes that fit form a prefix in x (need grows with cx); stage up to the right-most of them
    const int jmax = __reduce_max_sync(0xffffffffu, fits ? jr : 0);
    unsigned off[9];
    unsigned total = 0;
#pragma unroll
    for (int r = 0; r < 9; ++r) { off[r] = total; total += sm.cs[r][jmax] - sm.cs[r][0]; }
    if (total > 0) {
      unsigned bulk_bytes = 0;
#pragma unroll
      for (int r = 0; r < 9; ++r) { const unsigned cnt = sm.cs[r][jmax] - sm.cs[r][0]; if (cnt >= LVO_KT_BULK_MIN) bulk_bytes += cnt * 16u; }
      if (bulk_bytes && ln == 0) kt_mbar_expect_tx(&sm.bar, bulk_bytes);
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const unsigned b = sm.cs[r][0], cnt = sm.cs[r][jmax] - b;
        if (cnt >= LVO_KT_BULK_MIN) { if (ln == 0) kt_bulk_g2s(&sm.pts[off[r]], g.pts + b, cnt * 16u, &sm.bar); }
        else for (unsigned t = ln; t < cnt; t += 32) sm.pts[off[r] + t] = __ldg(g.pts + b + t);
      }
      if (bulk_bytes) { while (!kt_mbar_try_wait(&sm.bar, phase)) {} phase ^= 1u; }
      __syncwarp();
    }
    // ---- every lane of the group scans the 27 cells of its own query out of shared memory
    if (fits) {
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const unsigned base = sm.cs[r][0];
        const unsigned b = off[r] + (sm.cs[r][jl] - base), e = off[r] + (sm.cs[r][jr] - base);
        for (unsigned t = b; t < e; ++t) {
          const float4 p = sm.pts[t];
          tk.insert(sqdist3(p, qx, qy, qz), __float_as_int(p.w));
        }
      }
    }
    todo &= ~fit_mask;
    __syncwarp();   // all reads of pts / cs are done before the next group overwrites them
  }
  valid = tk.id[4] != INT_MAX && (double)tk.d[4] < (double)max_sq;
}

// Shared-memory set-up common to the tiled kernels: one barrier per warp.
__device__ __forceinline__ KtWarpSmem* kt_setup(unsigned char* smem_raw) {
  KtWarpSmem* all = reinterpret_cast<KtWarpSmem*>(smem_raw);
  const unsigned ln = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (ln == 0) kt_mbar_init(&all[w].bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  return &all[w];
}
#define LVO_KT_SMEM_BYTES (sizeof(KtWarpSmem) * LVO_KT_WARPS)
