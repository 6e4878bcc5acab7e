// Two-level binning: replaces duplicateWithKeys + cub::DeviceRadixSort::SortPairs + identifyTileRanges
// ($RAST/cuda_rasterizer/rasterizer_impl.cu:70-138, 278-321).  cub-free.
//
// The reference sorts R (tile<<32 | depth bits) keys with a 6-pass LSD radix sort (~152 B of HBM traffic
// per tile instance).  The tile id is the most significant field and per-tile counts are a by-product of
// the projection kernel, so the same total order is produced with far less traffic:
//   level 1  counting sort by tile: an exclusive scan of the tile histogram gives the tile ranges directly
//            (no memset + identifyTileRanges pass); every Gaussian then scatters one 8-byte entry
//            (depth bits << 32 | gaussian index) per touched tile through per-tile write cursors.  Counters
//            and cursors are split into SUBBINS per tile so the returning atomics do not serialise on one
//            L2 address for a crowded tile;
//   level 2  each tile's segment is radix-sorted by one CTA -- in shared memory when it fits, in place in
//            global memory (L2-resident) otherwise -- and the sorted order is materialised as a contiguous
//            slab of 48-byte splat records, which is what the render kernels stream with TMA bulk copies.
// Order parity: the reference's sort is stable and its emit order is ascending Gaussian index
// (rasterizer_impl.cu:98-108), so "stable by (tile, depth bits)" == total order by
// (tile, depth bits, gaussian index).  Level 2 sorts by depth bits with a stable LSD radix sort and, only if
// it then finds equal-depth neighbours out of index order, redoes the sort over the full 64-bit entry; the
// result never depends on the (non-deterministic) arrival order of the level-1 scatter.
#include "gsr_internal.cuh"

namespace gsr {

namespace {

constexpr unsigned FULL = 0xffffffffu;

// ---- level 1a: tile totals + exclusive scan (one CTA; T is at most a few 10^5) --------------------
// Tiles are taken in rounds of SCAN_THREADS consecutive tiles (coalesced, 16 independent loads per thread),
// each round is block-scanned and chained through a running carry.
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS) k_tile_scan(ImageView im, int T) {
  __shared__ unsigned warp_sums[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long cap = im.hdr->capacity;
  unsigned long long carry = 0;
  for (int t0 = 0; t0 < T; t0 += SCAN_THREADS) {
    const int t = t0 + tid;
    unsigned cnt[SUBBINS];
    unsigned c = 0;
#pragma unroll
    for (int s = 0; s < SUBBINS; s++) { cnt[s] = t < T ? im.tile_count[s * T + t] : 0u; c += cnt[s]; }
    unsigned v = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(FULL, v, o); if (lane >= o) v += u; }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    unsigned wv = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(FULL, wv, o); if (lane >= o) wv += u; }
    const unsigned round_total = __shfl_sync(FULL, wv, 31);
    const unsigned before_warp = __shfl_sync(FULL, wv, max(warp - 1, 0));
    unsigned long long run = carry + (warp ? before_warp : 0u) + (v - c);
    if (t < T) {
      // tiles whose segment does not fit the binning capacity render nothing (pipelined-mode overflow)
      const bool fits = run + c <= cap;
      im.tile_range[t] = (c && fits) ? make_uint2((unsigned)run, (unsigned)(run + c)) : make_uint2(0u, 0u);
#pragma unroll
      for (int s = 0; s < SUBBINS; s++) {
        im.tile_cursor[s * T + t] = fits ? (unsigned)run : 0xffffffffu;
        run += cnt[s];
      }
    }
    carry += round_total;
    __syncthreads();
  }
  if (tid == 0) {
    im.hdr->num_rendered = carry;
    im.hdr->overflow = carry > cap ? 1u : 0u;
  }
}

// ---- level 1b: scatter one entry per (Gaussian, touched tile) ------------------------------------
__global__ void __launch_bounds__(256) k_scatter(int P, int gx, int T, GeomView g, ImageView im, BinView b) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  unsigned depth_bits = 0;
  if (idx < P) {
    unpack_rect(g.rect[idx], x0, y0, x1, y1);
    if (x1 > x0) depth_bits = __float_as_uint(g.splat[(size_t)idx * SPLAT_F4 + 1].z);
  }
  const int w = x1 - x0, n = w * (y1 - y0);
  const unsigned lane = threadIdx.x & 31;
  constexpr int kBig = 32;
  auto put = [&](int tile, unsigned dbits, int gidx) {
    unsigned* cur = &im.tile_cursor[subbin_of(gidx) * T + tile];
    if (*cur == 0xffffffffu) return;  // tile dropped (capacity overflow)
    const unsigned slot = atomicAdd(cur, 1u);
    b.ents[slot] = ((unsigned long long)dbits << 32) | (unsigned)gidx;
  };
  if (n > 0 && n <= kBig) {
    for (int y = y0; y < y1; y++)
      for (int x = x0; x < x1; x++) put(y * gx + x, depth_bits, idx);
  }
  // a splat covering many tiles is walked by the whole warp (see for_each_tile in gsr_preprocess.cu)
  unsigned big = __ballot_sync(FULL, n > kBig);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const int bx0 = __shfl_sync(FULL, x0, src), by0 = __shfl_sync(FULL, y0, src);
    const int bw = __shfl_sync(FULL, w, src), bn = __shfl_sync(FULL, n, src);
    const unsigned sd = __shfl_sync(FULL, depth_bits, src);
    const int sidx = __shfl_sync(FULL, idx, src);
    for (int i = lane; i < bn; i += 32) put((by0 + i / bw) * gx + bx0 + i % bw, sd, sidx);
  }
}

// ---- level 2: per-tile sort + slab gather ---------------------------------------------------------
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_CAP = 4096;       // entries per buffer kept in shared memory (2 x 32 KB)
constexpr int SMALL_N = 256;         // up to here a bitonic network in shared memory is cheaper
constexpr size_t SORT_SMEM = 2 * SORT_CAP * sizeof(unsigned long long);

// Bitonic network in its "all ascending" form (first step of every merge mirrors the partner index), which
// tolerates an arbitrary n by treating indices >= n as +inf: such pairs never swap, so they are skipped.
__device__ __forceinline__ void bitonic_small(unsigned long long* k, unsigned n) {
  unsigned npad = 1;
  while (npad < n) npad <<= 1;
  const unsigned half = npad >> 1;
  for (unsigned blk = 2; blk <= npad; blk <<= 1) {
    const unsigned hb = blk >> 1;
    for (unsigned t = threadIdx.x; t < half; t += SORT_THREADS) {
      const unsigned base = (t / hb) * blk, off = t % hb;
      const unsigned i = base + off, p = base + blk - 1 - off;
      if (p < n) { const unsigned long long a = k[i], c = k[p]; if (a > c) { k[i] = c; k[p] = a; } }
    }
    __syncthreads();
    for (unsigned j = blk >> 2; j > 0; j >>= 1) {
      for (unsigned t = threadIdx.x; t < half; t += SORT_THREADS) {
        const unsigned i = 2 * j * (t / j) + (t % j), p = i + j;
        if (p < n) { const unsigned long long a = k[i], c = k[p]; if (a > c) { k[i] = c; k[p] = a; } }
      }
      __syncthreads();
    }
  }
}

struct RadixShared {
  unsigned hw[SORT_WARPS][256];  // per-warp digit counts, then running write offsets
  unsigned tot[256];
  unsigned wsum[SORT_WARPS];
  unsigned flag;
};

// Lanes of `act` holding the same 8-bit digit (what __match_any_sync returns), built from 8 ballots:
// MATCH.ANY goes through the MIO pipe with a long latency on sm_100, VOTE does not.
__device__ __forceinline__ unsigned match_digit(unsigned act, unsigned d) {
  unsigned peers = act;
#pragma unroll
  for (int b = 0; b < 8; b++) {
    const bool bit = (d >> b) & 1u;
    const unsigned m = __ballot_sync(act, bit);
    peers &= bit ? m : ~m;
  }
  return peers;
}

// One stable LSD pass on the byte at bit `shift` of the 64-bit entries: src[0,n) -> dst[0,n).
// Warp w owns the contiguous chunk [w*m, (w+1)*m); stability = (chunk, position).  Ranks inside a 32-entry
// round come from ballots (match_digit), so no per-element atomics are needed.
__device__ __forceinline__ void radix_pass(const unsigned long long* src, unsigned long long* dst, unsigned n,
                                           unsigned shift, RadixShared& sh) {
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned m = (n + SORT_WARPS - 1) / SORT_WARPS;
  const unsigned beg = min(n, warp * m), end = min(n, beg + m);
  for (unsigned i = tid; i < SORT_WARPS * 256; i += SORT_THREADS) (&sh.hw[0][0])[i] = 0;
  __syncthreads();
  for (unsigned base = beg; base < end; base += 32) {
    const unsigned i = base + lane;
    const bool valid = i < end;
    const unsigned act = __ballot_sync(FULL, valid);
    if (valid) {
      const unsigned d = (unsigned)(src[i] >> shift) & 255u;
      const unsigned peers = match_digit(act, d);
      if ((peers & ((1u << lane) - 1)) == 0) sh.hw[warp][d] += __popc(peers);  // leader of each digit group
    }
    __syncwarp();
  }
  __syncthreads();
  // offsets: digit-major, warp-minor exclusive scan of hw[w][d]
  {
    const unsigned d = tid;  // SORT_THREADS == 256 digits
    unsigned run = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) { const unsigned c = sh.hw[w][d]; sh.hw[w][d] = run; run += c; }
    unsigned v = run;  // total of digit d; block exclusive scan over d
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(FULL, v, o); if (lane >= o) v += u; }
    if (lane == 31) sh.wsum[warp] = v;
    __syncthreads();
    unsigned basew = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) basew += (w < (int)warp) ? sh.wsum[w] : 0u;
    sh.tot[d] = basew + v - run;
  }
  __syncthreads();
  for (unsigned base = beg; base < end; base += 32) {
    const unsigned i = base + lane;
    const bool valid = i < end;
    const unsigned act = __ballot_sync(FULL, valid);
    unsigned long long key = 0;
    unsigned d = 0, peers = 0, pos = 0;
    if (valid) {
      key = src[i];
      d = (unsigned)(key >> shift) & 255u;
      peers = match_digit(act, d);
      pos = sh.tot[d] + sh.hw[warp][d] + __popc(peers & ((1u << lane) - 1));
    }
    __syncwarp();
    if (valid) {
      if ((peers & ((1u << lane) - 1)) == 0) sh.hw[warp][d] += __popc(peers);
      dst[pos] = key;
    }
    __syncwarp();
  }
  __syncthreads();
}

// Sorts bits [lo, lo+32) of the entries with as many byte passes as there are non-constant bytes.
// Returns the buffer holding the result (a or b).
__device__ __forceinline__ unsigned long long* radix_field(unsigned long long* a, unsigned long long* b, unsigned n,
                                                           unsigned lo, RadixShared& sh) {
  // which bytes vary inside this tile?  (depths in one tile usually share sign/exponent bytes)
  if (threadIdx.x == 0) sh.flag = 0;
  __syncthreads();
  const unsigned first = (unsigned)(a[0] >> lo);
  unsigned diff = 0;
  for (unsigned i = threadIdx.x; i < n; i += SORT_THREADS) diff |= (unsigned)(a[i] >> lo) ^ first;
  diff = __reduce_or_sync(FULL, diff);
  if ((threadIdx.x & 31) == 0 && diff) atomicOr(&sh.flag, diff);
  __syncthreads();
  diff = sh.flag;
  __syncthreads();
  for (unsigned byte = 0; byte < 4; byte++) {
    if (((diff >> (8 * byte)) & 255u) == 0) continue;
    radix_pass(a, b, n, lo + 8 * byte, sh);
    unsigned long long* t = a; a = b; b = t;
  }
  return a;
}

__global__ void __launch_bounds__(SORT_THREADS) k_tile_sort(GeomView g, ImageView im, BinView b) {
  extern __shared__ __align__(16) unsigned long long sort_smem[];
  __shared__ RadixShared sh;
  const uint2 range = im.tile_range[blockIdx.x];
  const unsigned n = range.y - range.x;
  if (n == 0) return;
  unsigned long long* seg = b.ents + range.x;
  float4* out = b.slab + (size_t)range.x * SPLAT_F4;
  unsigned long long *A, *B;
  if (n <= SORT_CAP) {
    A = sort_smem; B = sort_smem + SORT_CAP;
    for (unsigned i = threadIdx.x; i < n; i += SORT_THREADS) A[i] = seg[i];
    __syncthreads();
  } else {
    // rare: more instances than fit in shared memory.  Ping-pong between the entry segment and the tile's
    // (not yet written) slab region, both L2-resident.
    A = seg; B = reinterpret_cast<unsigned long long*>(out);
    __syncthreads();
  }
  const unsigned long long* sorted;
  if (n <= SMALL_N) {
    bitonic_small(A, n);
    sorted = A;
  } else {
    unsigned long long* r = radix_field(A, B, n, 32, sh);
    // equal depth bits must come in ascending Gaussian index (stable-sort parity); arrival order is arbitrary
    unsigned bad = 0;
    for (unsigned i = threadIdx.x + 1; i < n; i += SORT_THREADS) bad |= (r[i] < r[i - 1]);
    if (__syncthreads_or(bad)) {
      unsigned long long* o = (r == A) ? B : A;
      r = radix_field(r, o, n, 0, sh);
      o = (r == A) ? B : A;
      r = radix_field(r, o, n, 32, sh);
    }
    sorted = r;
  }
  if (n > SORT_CAP && sorted != seg) {
    // the result sits in the slab region that the gather below overwrites: move it back first
    for (unsigned i = threadIdx.x; i < n; i += SORT_THREADS) seg[i] = sorted[i];
    __syncthreads();
    sorted = seg;
  }
  // gather the splat records in sorted order into the tile's contiguous slab
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

void launch_scatter(int P, int gx, int T, GeomView g, ImageView im, BinView b, cudaStream_t st) {
  k_scatter<<<(P + 255) / 256, 256, 0, st>>>(P, gx, T, g, im, b);
}

void launch_tile_sort(int T, GeomView g, ImageView im, BinView b, cudaStream_t st) {
  cudaFuncSetAttribute(k_tile_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT_SMEM);  // per device
  k_tile_sort<<<T, SORT_THREADS, SORT_SMEM, st>>>(g, im, b);
}

}  // namespace gsr
