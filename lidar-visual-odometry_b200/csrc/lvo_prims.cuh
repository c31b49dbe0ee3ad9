// lvo_prims.cuh — device-wide primitives written for this library (no CUB/Thrust): exclusive scan and a stable
// LSD radix sort of (u64 key, u32 value) pairs.  Both take their problem size from DEVICE memory (`d_n`) so that
// a frame can be enqueued without host synchronisation; grids are sized from capacities and grid-stride.
//
// The sort replaces three unstable std::sort calls of the reference path with a *stable* order, which is the
// tie-break contract of the oracle (SURVEY §7 hard part 3): voxel index inside pcl::VoxelGrid
// (scanRegistration.cpp:401-405, laserMapping.cpp:543-549,793-799).
#pragma once
#include "lvo_internal.h"

#define LVO_SCAN_THREADS 256
#define LVO_SCAN_ITEMS 8
#define LVO_SCAN_TILE (LVO_SCAN_THREADS * LVO_SCAN_ITEMS)

#define LVO_SORT_THREADS 256
#define LVO_SORT_ITEMS 16
#define LVO_SORT_TILE (LVO_SORT_THREADS * LVO_SORT_ITEMS)  // 4096 keys per tile
#define LVO_SORT_WARPS (LVO_SORT_THREADS / 32)

// ---------------------------------------------------------------------------------------------------------------
// block-level helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned warp_incl_scan(unsigned v) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= (unsigned)o) v += t;
  }
  return v;
}
// Exclusive scan of one value per thread over the block (blockDim.x multiple of 32, <= 1024).  Returns the
// exclusive prefix; *total gets the block sum.  `sm` needs 33 unsigned.
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned* sm, unsigned* total) {
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned inc = warp_incl_scan(v);
  if (lane == 31) sm[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned x = lane < nw ? sm[lane] : 0;
    unsigned xi = warp_incl_scan(x);
    sm[lane] = xi - x;
    if (lane == 31) sm[32] = xi;
  }
  __syncthreads();
  unsigned r = sm[w] + inc - v;
  *total = sm[32];
  __syncthreads();
  return r;
}

// Stable compaction rank of one FLAG per thread over the block, for loops that compact chunk after chunk: rank = number of set flags in
// lower threads, *total = set flags in the block.  One ballot per warp and ONE barrier per call (block_excl_scan: two shuffle scans and
// three barriers); `sm` needs 2 * 32 unsigned and `phase` must alternate 0 / 1 between consecutive calls (double buffering).
__device__ __forceinline__ unsigned block_flag_rank(bool flag, unsigned* sm, int phase, unsigned* total) {
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  unsigned* cnt = sm + 32 * (phase & 1);
  if (lane == 0) cnt[w] = __popc(m);
  __syncthreads();
  unsigned pre = 0, tot = 0;
  for (unsigned ww = 0; ww < nw; ++ww) { const unsigned c = cnt[ww]; tot += c; if (ww < w) pre += c; }
  *total = tot;
  return pre + __popc(m & ((1u << lane) - 1u));
}

// ---------------------------------------------------------------------------------------------------------------
// exclusive scan, in place, n from device memory
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_scan_reduce(const unsigned* __restrict__ data, const int* __restrict__ d_n, unsigned* __restrict__ partial) {
  __shared__ unsigned sm[33];
  const int n = *d_n;
  const int ntiles = (n + LVO_SCAN_TILE - 1) / LVO_SCAN_TILE;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    unsigned s = 0;
    const int base = tile * LVO_SCAN_TILE;
#pragma unroll
    for (int k = 0; k < LVO_SCAN_ITEMS; ++k) {
      int i = base + k * LVO_SCAN_THREADS + threadIdx.x;
      if (i < n) s += data[i];
    }
    unsigned tot;
    block_excl_scan(s, sm, &tot);
    if (threadIdx.x == 0) partial[tile] = tot;
  }
}
// single block: exclusive scan of the tile partials (any count), optional grand total
__global__ void k_scan_partials(unsigned* __restrict__ partial, const int* __restrict__ d_n, unsigned* __restrict__ d_total) {
  __shared__ unsigned sm[33];
  const int n = *d_n;
  const int ntiles = (n + LVO_SCAN_TILE - 1) / LVO_SCAN_TILE;
  unsigned carry = 0;
  for (int base = 0; base < ntiles; base += blockDim.x) {
    int i = base + threadIdx.x;
    unsigned v = i < ntiles ? partial[i] : 0;
    unsigned tot;
    unsigned ex = block_excl_scan(v, sm, &tot);
    if (i < ntiles) partial[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0 && d_total) *d_total = carry;
}
__global__ void k_scan_apply(unsigned* __restrict__ data, const int* __restrict__ d_n, const unsigned* __restrict__ partial) {
  __shared__ unsigned sm[33];
  const int n = *d_n;
  const int ntiles = (n + LVO_SCAN_TILE - 1) / LVO_SCAN_TILE;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // thread t owns ITEMS consecutive elements so that the block scan of per-thread sums gives element prefixes
    const int base = tile * LVO_SCAN_TILE + threadIdx.x * LVO_SCAN_ITEMS;
    unsigned v[LVO_SCAN_ITEMS];
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < LVO_SCAN_ITEMS; ++k) {
      v[k] = (base + k < n) ? data[base + k] : 0;
      s += v[k];
    }
    unsigned tot;
    unsigned ex = block_excl_scan(s, sm, &tot) + partial[tile];
#pragma unroll
    for (int k = 0; k < LVO_SCAN_ITEMS; ++k) {
      if (base + k < n) data[base + k] = ex;
      ex += v[k];
    }
  }
}

struct LvoScanScratch { unsigned* partial; int cap_tiles; };

// data[0..n) <- exclusive prefix sums; *d_total (optional) <- sum.  3 launches.
static inline void lvo_scan_exclusive(cudaStream_t st, unsigned* data, const int* d_n, int n_cap, unsigned* d_total, const LvoScanScratch& sc,
                                      long long* launches) {
  int tiles = lvo_div_up(n_cap, LVO_SCAN_TILE);
  int grid = tiles < 1184 ? (tiles < 1 ? 1 : tiles) : 1184;
  k_scan_reduce<<<grid, LVO_SCAN_THREADS, 0, st>>>(data, d_n, sc.partial);
  k_scan_partials<<<1, 1024, 0, st>>>(sc.partial, d_n, d_total);
  k_scan_apply<<<grid, LVO_SCAN_THREADS, 0, st>>>(data, d_n, sc.partial);
  if (launches) *launches += 3;
}

// ---------------------------------------------------------------------------------------------------------------
// stable LSD radix sort of (u64 key, u32 value), 8 bits per pass, ONE kernel per pass ("onesweep" structure):
//   k_sort_ghist     one sweep over the keys builds the digit histograms of ALL passes (shared-memory atomics, one global atomic per
//                    bin and block); it also opens a new epoch for the tile status words;
//   k_sort_onesweep  per pass: a block takes the next tile (atomic ticket, so a tile's predecessors are always running or done),
//                    ranks its 4096 keys by digit (warp match + per-warp counters, stable), publishes the tile's digit counts in a
//                    status word per (tile, digit) and obtains its global offsets by DECOUPLED LOOK-BACK over the preceding tiles
//                    (thread d walks digit d: add aggregates until a tile with an inclusive prefix is met), then scatters.
// Against the former histogram / 3-kernel scan / scatter sequence this is 1 launch per pass instead of 5, no per-tile histogram
// array in HBM and one read of the keys less per pass.  A status word is (tag << 32 | count) with tag = 2 * (8 * epoch + pass) + flag:
// the epoch is a DEVICE counter bumped by k_sort_ghist, so stale words of earlier sorts (or earlier replays of a captured graph)
// never match and the array needs no clearing.
// ---------------------------------------------------------------------------------------------------------------
struct LvoSortBufs {
  unsigned long long* keys[2];
  unsigned* vals[2];
  unsigned* ghist;               // [8][256] digit histograms of all passes, then [8] tile tickets
  unsigned long long* status;    // [cap_tiles][256]
  unsigned* d_epoch;             // device scalar
  int cap;                       // capacity in pairs
};
#define LVO_SORT_GHIST_WORDS (8 * 256 + 8)

__device__ __forceinline__ int lvo_sort_passes(int bits) { return (bits + 7) >> 3; }

__global__ void __launch_bounds__(256) k_sort_ghist(const unsigned long long* __restrict__ keys, const int* __restrict__ d_n, const int* __restrict__ d_bits,
                                                    unsigned* __restrict__ ghist, unsigned* __restrict__ d_epoch) {
  __shared__ unsigned h[8][256];
  const int passes = min(lvo_sort_passes(*d_bits), 8);
  const int n = *d_n;
  if (blockIdx.x == 0 && threadIdx.x == 0) *d_epoch = *d_epoch + 1u;   // read by the pass kernels, which start after this kernel has finished
  for (int k = threadIdx.x; k < 8 * 256; k += blockDim.x) (&h[0][0])[k] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long key = keys[i];
    for (int p = 0; p < passes; ++p) atomicAdd(&h[p][(unsigned)(key >> (8 * p)) & 255u], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < passes * 256; k += blockDim.x) { const unsigned v = (&h[0][0])[k]; if (v) atomicAdd(&ghist[k], v); }
}

__global__ void __launch_bounds__(LVO_SORT_THREADS, 2)
k_sort_onesweep(unsigned long long* __restrict__ keys0, unsigned long long* __restrict__ keys1, unsigned* __restrict__ vals0,
                unsigned* __restrict__ vals1, const int* __restrict__ d_n, const int* __restrict__ d_bits, int pass,
                unsigned* __restrict__ ghist, unsigned long long* status, const unsigned* __restrict__ d_epoch) {
  __shared__ unsigned wcnt[LVO_SORT_WARPS][256];
  __shared__ unsigned gbase[256];
  __shared__ unsigned sm_scan[33];
  __shared__ int s_tile;
  const int bits = *d_bits;
  if (pass * 8 >= bits) return;
  const int n = *d_n;
  const int ntiles = (n + LVO_SORT_TILE - 1) / LVO_SORT_TILE;
  const unsigned long long* kin = (pass & 1) ? keys1 : keys0;
  unsigned long long* kout = (pass & 1) ? keys0 : keys1;
  const unsigned* vin = (pass & 1) ? vals1 : vals0;
  unsigned* vout = (pass & 1) ? vals0 : vals1;
  const int shift = pass * 8;
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const unsigned tag = 2u * (8u * (*d_epoch) + (unsigned)pass);   // + 1: inclusive prefix
  unsigned* ticket = ghist + 8 * 256 + pass;
  // first output position of every digit: exclusive scan of this pass's histogram (thread d <-> digit d)
  unsigned tot_unused;
  const unsigned digit_base = block_excl_scan(ghist[pass * 256 + threadIdx.x], sm_scan, &tot_unused);
  while (true) {
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
    for (int k = threadIdx.x; k < LVO_SORT_WARPS * 256; k += LVO_SORT_THREADS) (&wcnt[0][0])[k] = 0;
    __syncthreads();
    const int tile = s_tile;
    if (tile >= ntiles) break;
    unsigned long long key[LVO_SORT_ITEMS];
    unsigned val[LVO_SORT_ITEMS];
    unsigned short lpos[LVO_SORT_ITEMS];
    const int wbase = tile * LVO_SORT_TILE + w * (32 * LVO_SORT_ITEMS);
#pragma unroll
    for (int r = 0; r < LVO_SORT_ITEMS; ++r) {
      const int i = wbase + r * 32 + lane;
      const bool valid = i < n;
      key[r] = valid ? kin[i] : 0ull;
      val[r] = valid ? vin[i] : 0u;
      const unsigned digit = (unsigned)(key[r] >> shift) & 255u;
      const unsigned m = __match_any_sync(0xffffffffu, valid ? digit : 256u);
      const unsigned rank = __popc(m & lt);
      unsigned base = 0;
      if (valid) base = wcnt[w][digit];
      __syncwarp();
      if (valid && rank == 0) wcnt[w][digit] = base + __popc(m);
      __syncwarp();
      lpos[r] = (unsigned short)(base + rank);
    }
    __syncthreads();
    {
      const unsigned d = threadIdx.x;  // one thread per digit (blockDim == 256)
      unsigned run = 0;
#pragma unroll
      for (int ww = 0; ww < LVO_SORT_WARPS; ++ww) {
        unsigned t = wcnt[ww][d];
        wcnt[ww][d] = run;
        run += t;
      }
      // decoupled look-back over the preceding tiles of this pass
      volatile unsigned long long* st = status;
      unsigned excl = 0;
      if (tile > 0) {
        st[(size_t)tile * 256 + d] = ((unsigned long long)tag << 32) | run;   // aggregate of this tile
        for (int t = tile - 1; t >= 0; --t) {
          unsigned long long v;
          do { v = st[(size_t)t * 256 + d]; } while (((unsigned)(v >> 32) | 1u) != (tag | 1u));
          excl += (unsigned)v;
          if ((unsigned)(v >> 32) & 1u) break;   // inclusive prefix: everything before tile t is in it
        }
      }
      st[(size_t)tile * 256 + d] = ((unsigned long long)(tag | 1u) << 32) | (excl + run);
      gbase[d] = digit_base + excl;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < LVO_SORT_ITEMS; ++r) {
      const int i = wbase + r * 32 + lane;
      if (i < n) {
        const unsigned digit = (unsigned)(key[r] >> shift) & 255u;
        const unsigned pos = gbase[digit] + wcnt[w][digit] + lpos[r];
        kout[pos] = key[r];
        vout[pos] = val[r];
      }
    }
    __syncthreads();
  }
}

// Sorts pairs held in bufs.keys[0]/vals[0]; after the call the result is in keys[p&1]/vals[p&1] with
// p = ceil(*d_bits / 8) (use lvo_sort_passes on the device).  max_bits bounds the launches.
static inline void lvo_sort_pairs(cudaStream_t st, const LvoSortBufs& b, const int* d_n, int n_cap, const int* d_bits, int max_bits,
                                  long long* launches) {
  int tiles = lvo_div_up(n_cap, LVO_SORT_TILE);
  int grid = tiles < 592 ? (tiles < 1 ? 1 : tiles) : 592;
  int passes = (max_bits + 7) / 8;
  if (passes > 8) passes = 8;
  cudaMemsetAsync(b.ghist, 0, sizeof(unsigned) * LVO_SORT_GHIST_WORDS, st);
  k_sort_ghist<<<max(1, min(lvo_div_up(n_cap, 256 * 16), 592)), 256, 0, st>>>(b.keys[0], d_n, d_bits, b.ghist, b.d_epoch);
  for (int p = 0; p < passes; ++p)
    k_sort_onesweep<<<grid, LVO_SORT_THREADS, 0, st>>>(b.keys[0], b.keys[1], b.vals[0], b.vals[1], d_n, d_bits, p, b.ghist, b.status, b.d_epoch);
  if (launches) *launches += 1 + passes;
}
