// lvo_voxel.cuh — segmented, sort-based voxel-grid downsample: pcl::VoxelGrid<PointXYZI>::applyFilter restated
// (PCL 1.8.0 voxel_grid.hpp; reference call sites src/laserMapping.cpp:543-549 and :793-799).  One invocation
// filters many independent clouds ("segments": corner/surf stacks of every lane, or every valid map cube of every
// lane) at once:
//
//   k_vx_reset / k_vx_bbox     getMinMax3D per segment (ordered-int float atomics, warp-aggregated)
//   k_vx_prepare               min_b / div_b per segment, the INT_MAX overflow rule ("leaf size too small" -> the
//                              cloud passes through unchanged), number of significant key bits
//   k_vx_keys                  key = segment << vbits | voxel idx, value = item index
//   lvo_sort_pairs             stable LSD radix sort => (idx, input index) order = the oracle's tie-break contract
//   k_vx_heads + scan          one output per run of equal keys
//   k_vx_centroid              centroid of x, y, z, intensity accumulated sequentially in sorted order in float
//                              and divided by the count (PCL CentroidPoint), which makes the result bit-exact
//   scan                       per-segment output offsets
//
// A segment with leaf == 0 is a *pseudo segment*: items are only grouped by `aux` (stable), never merged — used for
// points inserted into map cubes outside the 5x5x3 neighbourhood, which the reference does not re-filter.
#pragma once
#include "lvo_internal.h"
#include "lvo_prims.cuh"
#include <float.h>
#include <limits.h>

struct VoxelEngine {
  int cap_items, cap_segs;
  // inputs (filled by the caller's gather kernel)
  float4* in_pts;     // [cap_items]
  int* in_seg;        // [cap_items]
  unsigned* in_aux;   // [cap_items] grouping key for pseudo segments (ignored otherwise)
  int* d_n;           // number of items
  float* seg_leaf;    // [cap_segs]
  int* d_nsegs;       // number of segments in use
  // per-segment work
  int* seg_mn;        // [3][cap_segs] ordered-int encoded float min
  int* seg_mx;        // [3][cap_segs]
  int* seg_minb;      // [3][cap_segs]
  int* seg_div;       // [2][cap_segs]
  int* seg_mode;      // 0 voxel, 1 passthrough (overflow), 2 pseudo
  int* d_vbits;       // bits of the voxel-index field
  int* d_bits;        // total significant key bits
  LvoSortBufs sort;
  unsigned* outpos;   // [cap_items]
  // outputs
  float4* out_pts;    // [cap_items], grouped by segment in ascending order
  unsigned* out_aux;  // [cap_items] aux of the (first) source item
  unsigned* seg_out_cnt;    // [cap_segs]
  unsigned* seg_out_start;  // [cap_segs + 1] exclusive offsets
  unsigned* d_n_out;
  LvoScanScratch scan;
};

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ int bitlen64(unsigned long long v) { return v ? 64 - __clzll((long long)v) : 0; }

__global__ void k_vx_reset(VoxelEngine e) {
  const int nsegs = *e.d_nsegs;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nsegs; s += gridDim.x * blockDim.x) {
    for (int c = 0; c < 3; ++c) { e.seg_mn[c * e.cap_segs + s] = INT_MAX; e.seg_mx[c * e.cap_segs + s] = INT_MIN; }
    e.seg_out_cnt[s] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { *e.d_vbits = 0; *e.d_n_out = 0; }
}

__global__ void k_vx_bbox(VoxelEngine e) {
  const int n = *e.d_n;
  const int stride = gridDim.x * blockDim.x;
  for (int base = blockIdx.x * blockDim.x; base < n; base += stride) {
    const int i = base + threadIdx.x;
    const bool valid = i < n;
    int seg = -1;
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    if (valid) {
      seg = e.in_seg[i];
      const float4 p = e.in_pts[i];
      mn[0] = mx[0] = f2ord(p.x); mn[1] = mx[1] = f2ord(p.y); mn[2] = mx[2] = f2ord(p.z);
    }
    const unsigned m = __match_any_sync(0xffffffffu, seg);
    if (m == 0xffffffffu) {  // whole warp in one segment: reduce first
      if (seg >= 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { mn[c] = __reduce_min_sync(0xffffffffu, mn[c]); mx[c] = __reduce_max_sync(0xffffffffu, mx[c]); }
        if ((threadIdx.x & 31) == 0)
          for (int c = 0; c < 3; ++c) { atomicMin(&e.seg_mn[c * e.cap_segs + seg], mn[c]); atomicMax(&e.seg_mx[c * e.cap_segs + seg], mx[c]); }
      }
    } else if (valid) {
      for (int c = 0; c < 3; ++c) { atomicMin(&e.seg_mn[c * e.cap_segs + seg], mn[c]); atomicMax(&e.seg_mx[c * e.cap_segs + seg], mx[c]); }
    }
  }
}

__global__ void k_vx_prepare(VoxelEngine e) {
  const int nsegs = *e.d_nsegs;
  const int n = *e.d_n;
  int vb = 0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nsegs; s += gridDim.x * blockDim.x) {
    const float leaf = e.seg_leaf[s];
    int mode = 0, bits = 0;
    if (leaf == 0.f) {
      mode = 2; bits = 13;  // aux = cube index < 4851 < 2^13
    } else if (e.seg_mn[s] != INT_MAX) {
      const float inv = 1.0f / leaf;
      float mn[3], mx[3];
      for (int c = 0; c < 3; ++c) { mn[c] = ord2f(e.seg_mn[c * e.cap_segs + s]); mx[c] = ord2f(e.seg_mx[c * e.cap_segs + s]); }
      const long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1, dz = (long long)((mx[2] - mn[2]) * inv) + 1;
      if (dx * dy * dz > (long long)INT_MAX) {
        mode = 1; bits = bitlen64((unsigned long long)(n > 0 ? n - 1 : 0));
      } else {
        int minb[3], maxb[3];
        for (int c = 0; c < 3; ++c) { minb[c] = (int)floorf(mn[c] * inv); maxb[c] = (int)floorf(mx[c] * inv); e.seg_minb[c * e.cap_segs + s] = minb[c]; }
        const int d0 = maxb[0] - minb[0] + 1, d1 = maxb[1] - minb[1] + 1, d2 = maxb[2] - minb[2] + 1;
        e.seg_div[s] = d0; e.seg_div[e.cap_segs + s] = d1;
        unsigned long long prod = (unsigned long long)d0 * (unsigned long long)d1 * (unsigned long long)d2;
        bits = bitlen64(prod - 1);
        if (bits > 32) bits = 32;
      }
    }
    e.seg_mode[s] = mode;
    vb = max(vb, bits);
  }
  vb = __reduce_max_sync(0xffffffffu, vb);
  if ((threadIdx.x & 31) == 0 && vb > 0) atomicMax(e.d_vbits, vb);
}

__global__ void k_vx_keys(VoxelEngine e) {
  const int n = *e.d_n;
  const int vb = *e.d_vbits;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int nsegs = *e.d_nsegs;
    *e.d_bits = vb + bitlen64((unsigned long long)(nsegs > 0 ? nsegs - 1 : 0));
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = e.in_seg[i];
    const int mode = e.seg_mode[s];
    unsigned idx;
    if (mode == 2) idx = e.in_aux[i];
    else if (mode == 1) idx = (unsigned)i;
    else {
      const float inv = 1.0f / e.seg_leaf[s];
      const float4 p = e.in_pts[i];
      const int ijk0 = (int)(floorf(p.x * inv) - (float)e.seg_minb[s]);
      const int ijk1 = (int)(floorf(p.y * inv) - (float)e.seg_minb[e.cap_segs + s]);
      const int ijk2 = (int)(floorf(p.z * inv) - (float)e.seg_minb[2 * e.cap_segs + s]);
      const int d0 = e.seg_div[s], d1 = e.seg_div[e.cap_segs + s];
      idx = (unsigned)(ijk0 + ijk1 * d0 + ijk2 * d0 * d1);
    }
    e.sort.keys[0][i] = ((unsigned long long)(unsigned)s << vb) | (unsigned long long)idx;
    e.sort.vals[0][i] = (unsigned)i;
  }
}

__global__ void k_vx_heads(VoxelEngine e) {
  const int n = *e.d_n;
  const int vb = *e.d_vbits;
  const int p = lvo_sort_passes(*e.d_bits) & 1;
  const unsigned long long* keys = e.sort.keys[p];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const unsigned long long k = keys[t];
    bool head = (t == 0) || (k != keys[t - 1]);
    if (!head && e.seg_mode[(int)(k >> vb)] == 2) head = true;
    e.outpos[t] = head ? 1u : 0u;
  }
}

// One thread per sorted item: every lane gathers its own point (independent loads), then the head of each run adds the points
// of the following lanes IN ORDER through warp shuffles — the same sequential float sum as a serial walk, without its chain of
// dependent loads.  A run that reaches the end of the warp's 32 items is continued from memory by its head.
__global__ void k_vx_centroid(VoxelEngine e) {
  const int n = *e.d_n;
  const int vb = *e.d_vbits;
  const int p = lvo_sort_passes(*e.d_bits) & 1;
  const unsigned long long* keys = e.sort.keys[p];
  const unsigned* vals = e.sort.vals[p];
  const unsigned lane = threadIdx.x & 31;
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n; base += gridDim.x * blockDim.x) {
    const int t = base + (int)lane;
    const bool in = t < n;
    const unsigned long long k = in ? keys[t] : 0ull;
    const unsigned v = in ? vals[t] : 0u;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in) q = e.in_pts[v];
    unsigned long long kprev = __shfl_up_sync(0xffffffffu, k, 1);
    if (lane == 0 && in && t > 0) kprev = keys[t - 1];
    const int seg = (int)(k >> vb);
    const bool pseudo = in && e.seg_mode[seg] == 2;
    const bool head = in && (t == 0 || k != kprev || pseudo);
    const unsigned stops = __ballot_sync(0xffffffffu, head || !in);
    const unsigned above = stops & ~((2u << lane) - 1u);
    const int end = above ? __ffs(above) - 1 : 32;        // first lane after this lane that is not part of its run
    const int len = head ? end - (int)lane : 0;
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    float sx = 0.f + q.x, sy = 0.f + q.y, sz = 0.f + q.z, si = 0.f + q.w;   // 0 + x as in the serial sum (-0 becomes +0)
    for (int d = 1; d < maxlen; ++d) {
      const float nx = __shfl_down_sync(0xffffffffu, q.x, d), ny = __shfl_down_sync(0xffffffffu, q.y, d);
      const float nz = __shfl_down_sync(0xffffffffu, q.z, d), ni = __shfl_down_sync(0xffffffffu, q.w, d);
      if (d < len) { sx += nx; sy += ny; sz += nz; si += ni; }
    }
    if (head) {
      int cnt = len;
      if (end == 32 && !pseudo) {   // the run may go on in the next 32 items
        int u = base + 32;
        while (u < n && keys[u] == k) {
          const float4 r = e.in_pts[vals[u]];
          sx += r.x; sy += r.y; sz += r.z; si += r.w;
          ++u;
        }
        cnt = u - t;
      }
      const float c = (float)cnt;
      const unsigned o = e.outpos[t];
      e.out_pts[o] = make_float4(sx / c, sy / c, sz / c, si / c);
      e.out_aux[o] = e.in_aux[v];
      const unsigned same = __match_any_sync(__activemask(), seg);   // heads of one segment in this warp: one atomic
      if (lane == (unsigned)(__ffs(same) - 1)) atomicAdd(&e.seg_out_cnt[seg], (unsigned)__popc(same));
    }
  }
}

// per-segment output offsets: exclusive scan of seg_out_cnt (at most a few ten thousand segments: one block, one launch)
__global__ void __launch_bounds__(1024) k_vx_seg_offsets(VoxelEngine e) {
  __shared__ unsigned sm[33];
  const int nsegs = *e.d_nsegs;
  unsigned carry = 0;
  for (int base = 0; base < nsegs; base += blockDim.x) {
    const int s = base + threadIdx.x;
    const unsigned v = s < nsegs ? e.seg_out_cnt[s] : 0u;
    unsigned tot;
    const unsigned ex = block_excl_scan(v, sm, &tot);
    if (s < nsegs) e.seg_out_start[s] = carry + ex;
    carry += tot;
  }
}

// n_items_cap / n_segs_cap bound the grids for this invocation (<= engine capacities).
static inline void lvo_voxel_run(cudaStream_t st, const VoxelEngine& e, int n_items_cap, int n_segs_cap, int max_seg_bits, long long* launches) {
  const int gi = max(1, min(lvo_div_up(n_items_cap, 256), 1184));
  const int gs = max(1, min(lvo_div_up(n_segs_cap, 256), 148));
  k_vx_reset<<<gs, 256, 0, st>>>(e);
  k_vx_bbox<<<gi, 256, 0, st>>>(e);
  k_vx_prepare<<<gs, 256, 0, st>>>(e);
  k_vx_keys<<<gi, 256, 0, st>>>(e);
  lvo_sort_pairs(st, e.sort, e.d_n, n_items_cap, e.d_bits, 32 + max_seg_bits, launches);
  k_vx_heads<<<gi, 256, 0, st>>>(e);
  lvo_scan_exclusive(st, e.outpos, e.d_n, n_items_cap, e.d_n_out, e.scan, launches);
  k_vx_centroid<<<gi, 256, 0, st>>>(e);
  k_vx_seg_offsets<<<1, 1024, 0, st>>>(e);
  if (launches) *launches += 7;
}
