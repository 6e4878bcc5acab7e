// Two-level binning: replaces duplicateWithKeys + cub::DeviceRadixSort::SortPairs + identifyTileRanges
// ($RAST/cuda_rasterizer/rasterizer_impl.cu:70-138, 278-321).
//
// The reference sorts R (tile<<32 | depth bits) keys with a 6-pass LSD radix sort (~152 B of HBM traffic
// per tile instance).  The tile id is the most significant field and per-tile counts are a by-product of
// the projection kernel, so the same total order is produced with far less traffic:
//   level 1  counting sort by tile: exclusive scan of the tile histogram gives the tile ranges directly
//            (no memset + identifyTileRanges pass), then every Gaussian scatters one 8-byte entry
//            (depth bits << 32 | gaussian index) per touched tile through a per-tile cursor;
//   level 2  each tile's segment is sorted by that 64-bit entry inside shared memory (one CTA per tile)
//            and the sorted order is materialised as a contiguous slab of 48-byte splat records, which is
//            what the render kernels stream with TMA bulk copies.
// Order parity: the reference's sort is stable and its emit order is ascending Gaussian index
// (rasterizer_impl.cu:98-108), so "stable by (tile, depth bits)" == total order by
// (tile, depth bits, gaussian index); sorting the 64-bit entries reproduces it exactly, independent of the
// (non-deterministic) arrival order of the level-1 scatter.
#include "gsr_internal.cuh"

namespace gsr {

namespace {

// ---- level 1a: exclusive scan over the tile histogram (T <= a few 10^5; one CTA) -----------------
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS) k_tile_scan(ImageView im, int T) {
  __shared__ unsigned long long warp_sums[32];
  __shared__ unsigned long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (T + SCAN_THREADS - 1) / SCAN_THREADS;
  const int beg = min(T, tid * per), end = min(T, beg + per);
  unsigned long long local = 0;
  for (int t = beg; t < end; t++) local += im.tile_count[t];
  // block exclusive scan of `local`
  unsigned long long v = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  if (lane == 31) warp_sums[warp] = v;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += u;
    }
    warp_sums[lane] = w;
    if (lane == 31) carry_s = w;
  }
  __syncthreads();
  unsigned long long run = (v - local) + (warp ? warp_sums[warp - 1] : 0ull);
  const unsigned long long cap = im.hdr->capacity;
  for (int t = beg; t < end; t++) {
    const unsigned c = im.tile_count[t];
    // tiles whose segment does not fit the binning capacity render nothing (pipelined-mode overflow)
    const bool fits = run + c <= cap;
    im.tile_range[t] = (c && fits) ? make_uint2((unsigned)run, (unsigned)(run + c)) : make_uint2(0u, 0u);
    im.tile_cursor[t] = fits ? (unsigned)run : 0xffffffffu;
    run += c;
  }
  if (tid == 0) {
    const unsigned long long R = carry_s;
    im.hdr->num_rendered = R;
    im.hdr->overflow = R > cap ? 1u : 0u;
  }
}

// ---- level 1b: scatter one entry per (Gaussian, touched tile) ------------------------------------
__global__ void __launch_bounds__(256) k_scatter(int P, int gx, GeomView g, ImageView im, BinView b) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  unsigned depth_bits = 0;
  if (idx < P) {
    unpack_rect(g.rect[idx], x0, y0, x1, y1);
    if (x1 > x0) depth_bits = __float_as_uint(g.splat[(size_t)idx * SPLAT_F4 + 1].z);
  }
  // Lanes walking another lane's big rect need that lane's key: fetch it by shuffle inside the visitor.
  const unsigned long long my_key = ((unsigned long long)depth_bits << 32) | (unsigned)idx;
  const unsigned klo = (unsigned)my_key, khi = (unsigned)(my_key >> 32);
  // for_each_tile (see gsr_preprocess.cu) is re-stated here because the visitor needs the source lane.
  const int w = x1 - x0, n = w * (y1 - y0);
  const unsigned lane = threadIdx.x & 31;
  constexpr int kBig = 32;
  auto put = [&](int tile, unsigned long long key) {
    const unsigned cur = im.tile_cursor[tile];
    if (cur == 0xffffffffu) return;  // tile dropped (capacity overflow)
    const unsigned slot = atomicAdd(&im.tile_cursor[tile], 1u);
    b.ents[slot] = key;
  };
  if (n > 0 && n <= kBig) {
    for (int y = y0; y < y1; y++)
      for (int x = x0; x < x1; x++) put(y * gx + x, my_key);
  }
  unsigned big = __ballot_sync(0xffffffffu, n > kBig);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const int bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
    const int bw = __shfl_sync(0xffffffffu, w, src), bn = __shfl_sync(0xffffffffu, n, src);
    const unsigned slo = __shfl_sync(0xffffffffu, klo, src), shi = __shfl_sync(0xffffffffu, khi, src);
    const unsigned long long key = ((unsigned long long)shi << 32) | slo;
    for (int i = lane; i < bn; i += 32) put((by0 + i / bw) * gx + bx0 + i % bw, key);
  }
}

// ---- level 2: per-tile sort + slab gather ---------------------------------------------------------
// Bitonic network in its "all ascending" form (first step of every merge mirrors the partner index), which
// tolerates an arbitrary n by treating indices >= n as +inf: such pairs never swap, so they are skipped.
constexpr int SORT_THREADS = 256;
constexpr int SORT_CAP = 4096;  // entries sorted in shared memory (32 KB); larger tiles sort in global

template <typename Ptr>
__device__ __forceinline__ void cmpxchg(Ptr k, unsigned i, unsigned p) {
  const unsigned long long a = k[i], c = k[p];
  if (a > c) { k[i] = c; k[p] = a; }
}

template <typename Ptr>
__device__ __forceinline__ void bitonic_sort(Ptr k, unsigned n) {
  if (n < 2) return;
  unsigned npad = 1;
  while (npad < n) npad <<= 1;
  const unsigned half = npad >> 1;
  for (unsigned blk = 2; blk <= npad; blk <<= 1) {
    const unsigned hb = blk >> 1;
    for (unsigned t = threadIdx.x; t < half; t += SORT_THREADS) {
      const unsigned base = (t / hb) * blk, off = t % hb;
      const unsigned i = base + off, p = base + blk - 1 - off;
      if (p < n) cmpxchg(k, i, p);
    }
    __syncthreads();
    for (unsigned j = blk >> 2; j > 0; j >>= 1) {
      for (unsigned t = threadIdx.x; t < half; t += SORT_THREADS) {
        const unsigned i = 2 * j * (t / j) + (t % j), p = i + j;
        if (p < n) cmpxchg(k, i, p);
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(SORT_THREADS) k_tile_sort(GeomView g, ImageView im, BinView b) {
  __shared__ unsigned long long keys[SORT_CAP];
  const uint2 range = im.tile_range[blockIdx.x];
  const unsigned n = range.y - range.x;
  if (n == 0) return;
  unsigned long long* seg = b.ents + range.x;
  const unsigned long long* sorted;
  if (n <= SORT_CAP) {
    for (unsigned i = threadIdx.x; i < n; i += SORT_THREADS) keys[i] = seg[i];
    __syncthreads();
    bitonic_sort(keys, n);
    sorted = keys;
  } else {
    __syncthreads();
    bitonic_sort(seg, n);  // rare: a tile with more instances than fit in shared memory
    sorted = seg;
  }
  // gather the splat records in sorted order into the tile's contiguous slab
  float4* out = b.slab + (size_t)range.x * SPLAT_F4;
  for (unsigned i = threadIdx.x; i < n; i += SORT_THREADS) {
    const unsigned id = (unsigned)sorted[i];
    const float4* s = g.splat + (size_t)id * SPLAT_F4;
    const float4 q0 = s[0], q1 = s[1], q2 = s[2];
    out[(size_t)i * SPLAT_F4 + 0] = q0;
    out[(size_t)i * SPLAT_F4 + 1] = q1;
    out[(size_t)i * SPLAT_F4 + 2] = q2;
  }
}

}  // namespace

void launch_tile_scan(ImageView im, int T, cudaStream_t st) { k_tile_scan<<<1, SCAN_THREADS, 0, st>>>(im, T); }

void launch_scatter(int P, int gx, GeomView g, ImageView im, BinView b, cudaStream_t st) {
  k_scatter<<<(P + 255) / 256, 256, 0, st>>>(P, gx, g, im, b);
}

void launch_tile_sort(int T, GeomView g, ImageView im, BinView b, cudaStream_t st) {
  k_tile_sort<<<T, SORT_THREADS, 0, st>>>(g, im, b);
}

}  // namespace gsr
