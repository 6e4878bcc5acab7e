// Per-tile alpha compositing, forward and backward: replaces renderCUDA
// ($RAST/cuda_rasterizer/forward.cu:261-397 and backward.cu:415-610).
//
// What is kept exactly (observable behaviour): a pixel evaluates the Gaussians of its 16x16
// tile's sorted list in order, with the reference's per-pair arithmetic (same FMA contraction as its SASS,
// precise expf), thresholds (power>0, alpha<1/255, T<1e-4, median at T crossing 0.5), contributor counting
// and un-normalised depth / no-background colour outputs.
//
// What is different (B200):
//  * a dedicated producer warp walks the tile's sorted index list and stages the 48-byte splat records of
//    the next batches into a shared-memory ring with asynchronous 16-byte copies (cp.async, completion on an
//    mbarrier), NSTAGE batches ahead of the 8 consumer warps -- no CTA-wide barrier in the loop, and
//    rgb/depth come from shared memory instead of per-pair global loads (forward.cu:365-366).  Only the
//    part of the list a tile really consumes is ever gathered (early termination / n_contrib bound);
//  * each warp owns an 8x4 pixel sub-tile and first tests every staged Gaussian against its sub-tile with a
//    conservative bound on alpha (minimum of the conic form over the rectangle); only survivors are
//    evaluated.  A pair is skipped only if the reference would `continue` past it for every pixel of the
//    sub-tile (alpha < 1/255), so results are unchanged while most pair evaluations disappear;
//  * backward: the 10 per-Gaussian gradient components are reduced across the warp with a transposed
//    shuffle butterfly and issued as ONE predicated red.global per (warp, Gaussian) into a packed 48-byte
//    accumulator, instead of 11-12 atomicAdd per (pixel, Gaussian) pair (backward.cu:559-607); traversal
//    starts at the tile's largest n_contrib instead of the end of the list.
#include "gsr_internal.cuh"
#include <cstdlib>

namespace gsr {

namespace {

constexpr int RB = 128;                       // records per pipeline stage
constexpr int NSTAGE = 4;                     // ring depth: consumer warps may drift this many batches apart
constexpr int STAGE_F4 = RB * SPLAT_F4;       // float4 per stage (6 KB)
constexpr int NCONS = TILE_PIX / 32;          // 8 consumer warps (one 8x4 sub-tile each) + 1 producer warp
constexpr int RENDER_THREADS = TILE_PIX + 32;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// asynchronous 16-byte global->shared copy (LDGSTS, L2 only) and its completion hook on an mbarrier
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(unsigned long long* bar) {
  // the arrival fires once all of this thread's earlier cp.async have landed; .noinc: counted in the init value
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// producer warp: gather `cnt` records listed in ids[0,cnt) into a stage
__device__ __forceinline__ void stage_gather(float4* dst, const float4* __restrict__ splat,
                                             const uint32_t* __restrict__ ids, int cnt, int lane,
                                             unsigned long long* full) {
  for (int r = lane; r < cnt; r += 32) {
    const float4* src = splat + (size_t)ids[r] * SPLAT_F4;
    cp_async16(dst + r * SPLAT_F4 + 0, src + 0);
    cp_async16(dst + r * SPLAT_F4 + 1, src + 1);
    cp_async16(dst + r * SPLAT_F4 + 2, src + 2);
  }
  cp_async_arrive(full);
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned b = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(b),
      "r"(parity)
      : "memory");
}

// Conservative sub-tile test.  Returns false only if alpha = min(0.99, o*exp(power)) < 1/255 for EVERY pixel
// centre in [rx0,rx1]x[ry0,ry1] (the reference skips such pairs, forward.cu:353-355 / backward.cu:535-537).
// With q(d) = A dx^2 + 2B dx dy + C dy^2 = -2*power, alpha >= 1/255 needs q <= tau = 2 ln(255 o).  q is convex
// (conic positive definite), so its minimum over the rectangle is 0 if the centre is inside, else it lies on
// an edge facing the centre; each facing edge is a 1-D quadratic minimised in closed form.  The margin covers
// the rounding of the per-pixel evaluation (relative 1e-5 of the largest term magnitude + 1e-3 absolute);
// any non-finite / non-PD input keeps the pair.
__device__ __forceinline__ bool may_touch(const float4 q0, const float4 q1, float rx0, float ry0, float rx1, float ry1) {
  const float A = q0.z, B = q0.w, C = q1.x, o = q1.y;
  if (o < 0.0039f) return false;  // alpha <= o < 1/255 everywhere (exp(power) <= 1)
  const float dxlo = q0.x - rx1, dxhi = q0.x - rx0, dylo = q0.y - ry1, dyhi = q0.y - ry0;
  const bool inx = dxlo <= 0.f && dxhi >= 0.f, iny = dylo <= 0.f && dyhi >= 0.f;
  if (inx && iny) return true;
  if (!(A > 0.f && C > 0.f && A * C - B * B > 0.f)) return true;
  float qmin = 3.0e38f;
  if (!inx) {
    const float dxe = dxlo > 0.f ? dxlo : dxhi;
    const float dys = fminf(fmaxf(-B * dxe / C, dylo), dyhi);
    qmin = A * dxe * dxe + 2.f * B * dxe * dys + C * dys * dys;
  }
  if (!iny) {
    const float dye = dylo > 0.f ? dylo : dyhi;
    const float dxs = fminf(fmaxf(-B * dye / A, dxlo), dxhi);
    qmin = fminf(qmin, A * dxs * dxs + 2.f * B * dxs * dye + C * dye * dye);
  }
  const float mx = fmaxf(fabsf(dxlo), fabsf(dxhi)), my = fmaxf(fabsf(dylo), fabsf(dyhi));
  const float S = A * mx * mx + C * my * my + 2.f * fabsf(B) * mx * my;
  const float tau = 2.f * logf(255.f * o);
  return !(qmin > tau + 1e-5f * S + 1e-3f);
}

// ---------------------------------------------------------------------------------------------
// Shared-memory ring shared by both kernels: NSTAGE record batches, `full` barriers completed by the
// producer lanes' cp.async arrivals, `empty` barriers by one arrival per consumer warp.  Warp NCONS (the 9th)
// is the producer: it refills a stage as soon as every consumer warp released it.
// Consumer warps never meet at a CTA-wide barrier inside the loop, so a sub-tile with little work does not
// wait for a crowded one batch by batch.
// ---------------------------------------------------------------------------------------------
struct Ring {
  float4 buf[NSTAGE][STAGE_F4];
  unsigned long long full[NSTAGE];
  unsigned long long empty[NSTAGE];
  unsigned ndone;           // consumer warps that have no pixel left (forward early exit)
  unsigned maxc[NCONS];
};

__device__ __forceinline__ void ring_init(Ring& r) {
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; s++) { mbar_init(&r.full[s], 32); mbar_init(&r.empty[s], NCONS); }
    r.ndone = 0;
    fence_mbar_init();
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// HALF (experimental, GSR_HALFWARP=1; NOT yet validated on a GPU -- see DESIGN.md 7b.1): every consumer warp owns two 4x4
// blocks instead of one 8x4 block; each half-warp walks its own survivor mask, so one loop iteration composites two
// different Gaussians and the 16-px-wide cull removes pairs the 8x4 rectangle keeps.  Per-pixel order is unchanged.
template <bool HALF>
__global__ void __launch_bounds__(RENDER_THREADS) k_render_fwd(int W, int H, int gx, ImageView im, BinView bin,
                                                               const float4* __restrict__ splat,
                                                               float* __restrict__ out_color,
                                                               float* __restrict__ out_depth,
                                                               float* __restrict__ out_median,
                                                               float* __restrict__ out_opacity) {
  __shared__ __align__(128) Ring ring;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, tx = tile % gx, ty = tile / gx;
  const uint2 range = im.tile_range[tile];
  const int n = (int)(range.y - range.x);
  const int nb = (n + RB - 1) / RB;
  const uint32_t* ids = bin.point_list + range.x;
  ring_init(ring);

  if (warp == NCONS) {
    // ---------------- producer warp ----------------
    int issued = 0;
    for (int b = 0; b < nb; b++) {
      const int s = b % NSTAGE;
      bool stop = false;
      if (lane == 0) {
        if (b >= NSTAGE) {
          const unsigned par = (unsigned)((b / NSTAGE - 1) & 1);
          while (!mbar_try(&ring.empty[s], par)) {
            if (*(volatile unsigned*)&ring.ndone == NCONS) { stop = true; break; }
          }
        }
        if (*(volatile unsigned*)&ring.ndone == NCONS) stop = true;  // every pixel of the tile is finished
      }
      if (__shfl_sync(FULL, (int)stop, 0)) break;
      stage_gather(ring.buf[s], splat, ids + b * RB, min(RB, n - b * RB), lane, &ring.full[s]);
      issued = b + 1;
    }
    // every copy must have landed before the CTA's shared memory is released
    for (int b = max(0, issued - NSTAGE); b < issued; b++) mbar_wait(&ring.full[b % NSTAGE], (unsigned)((b / NSTAGE) & 1));
    return;
  }

  // ---------------- consumers: warp w owns the 8x4 sub-tile at (w&1, w>>1) ----------------
  const int sx0 = tx * TILE_X + (warp & 1) * 8, sy0 = ty * TILE_Y + (warp >> 1) * 4;
  const int px = HALF ? sx0 + 4 * (lane >> 4) + (lane & 3) : sx0 + (lane & 7);
  const int py = HALF ? sy0 + ((lane >> 2) & 3) : sy0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float rx0 = (float)sx0, ry0 = (float)sy0;
  const float rx1 = (float)min(sx0 + (HALF ? 3 : 7), W - 1), ry1 = (float)min(sy0 + 3, H - 1);
  const float rbx0 = (float)(sx0 + 4), rbx1 = (float)min(sx0 + 7, W - 1);  // HALF: the right-hand 4x4 block

  bool done = !inside;
  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
  float med_d = 15.0f, med_w = 0.f, med_id = 0.f;  // forward.cu:310-312
  unsigned last_contributor = 0;
  bool warp_done = __all_sync(FULL, done);
  if (warp_done && lane == 0) atomicAdd(&ring.ndone, 1u);

  for (int b = 0; b < nb; b++) {
    const int s = b % NSTAGE;
    const unsigned par = (unsigned)((b / NSTAGE) & 1);
    if (warp_done) {
      // nothing left for this warp: keep releasing stages until the whole tile is finished
      bool stop = false;
      while (!mbar_try(&ring.full[s], par)) {
        if (*(volatile unsigned*)&ring.ndone == NCONS) { stop = true; break; }
      }
      if (stop || *(volatile unsigned*)&ring.ndone == NCONS) break;
      if (lane == 0) mbar_arrive(&ring.empty[s]);
      continue;
    }
    mbar_wait(&ring.full[s], par);
    const int cnt = min(RB, n - b * RB);
    const float4* sb = ring.buf[s];
    for (int base = 0; base < cnt; base += 32) {
      const int j = base + lane;
      bool keep = false;
      if (j < cnt) keep = may_touch(sb[j * SPLAT_F4], sb[j * SPLAT_F4 + 1], rx0, ry0, rx1, ry1);
      unsigned mask = __ballot_sync(FULL, keep);
      unsigned maskB = 0;
      if constexpr (HALF) {
        bool keepB = false;
        if (j < cnt && rbx0 <= rbx1) keepB = may_touch(sb[j * SPLAT_F4], sb[j * SPLAT_F4 + 1], rbx0, ry0, rbx1, ry1);
        maskB = __ballot_sync(FULL, keepB);
      }
      while (HALF ? (mask | maskB) : mask) {
        int jj;
        if constexpr (HALF) {
          const unsigned mine = (lane & 16) ? maskB : mask;  // each half-warp pops its own next survivor
          jj = base + __ffs(mine) - 1;
          mask &= mask - 1;
          maskB &= maskB - 1;
          if (mine == 0 || done) continue;
        } else {
          jj = base + __ffs(mask) - 1;
          mask &= mask - 1;
          if (done) continue;
        }
        const float4 q0 = sb[jj * SPLAT_F4], q1 = sb[jj * SPLAT_F4 + 1];
        // forward.cu:343-356 with the contraction of the reference SASS
        const float dx = q0.x - pxf, dy = q0.y - pyf;
        const float t1 = __fmul_rn(__fmul_rn(dy, q1.x), dy);
        const float t2 = __fmul_rn(dx, q0.z);
        const float t3 = __fmul_rn(__fmul_rn(dx, q0.w), dy);
        const float power = __fmaf_rn(__fmaf_rn(dx, t2, t1), -0.5f, -t3);
        if (power > 0.0f) continue;
        const float alpha = fminf(0.99f, __fmul_rn(q1.y, expf(power)));
        if (alpha < 1.0f / 255.0f) continue;
        const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
        if (test_T < 0.0001f) { done = true; continue; }
        const float4 q2 = sb[jj * SPLAT_F4 + 2];
        C0 = __fmaf_rn(T, __fmul_rn(alpha, q1.w), C0);
        C1 = __fmaf_rn(T, __fmul_rn(alpha, q2.x), C1);
        C2 = __fmaf_rn(T, __fmul_rn(alpha, q2.y), C2);
        D = __fmaf_rn(T, __fmul_rn(alpha, q1.z), D);
        if (T > 0.5f && test_T < 0.5f) {
          med_d = q1.z;
          med_w = __fmul_rn(alpha, T);
          med_id = (float)__float_as_int(q2.z);
        }
        T = test_T;
        last_contributor = (unsigned)(b * RB + jj + 1);
      }
      warp_done = __all_sync(FULL, done);
      if (warp_done) break;
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&ring.empty[s]);
      if (warp_done) atomicAdd(&ring.ndone, 1u);
    }
  }

  if (inside) {
    const size_t pid = (size_t)py * W + px, HW = (size_t)W * H;
    im.final_T[pid] = T;
    im.n_contrib[pid] = last_contributor;
    out_color[pid] = C0;  // no background blend (forward.cu:389-390)
    out_color[HW + pid] = C1;
    out_color[2 * HW + pid] = C2;
    out_depth[pid] = D;
    out_median[pid] = med_d;
    out_median[HW + pid] = med_w;
    out_median[2 * HW + pid] = med_id;
    out_opacity[pid] = 1 - T;
  }
  // largest n_contrib of the tile bounds the backward traversal
  const unsigned wmax = __reduce_max_sync(FULL, last_contributor);
  if (lane == 0) ring.maxc[warp] = wmax;
  asm volatile("bar.sync 1, %0;" ::"n"(TILE_PIX) : "memory");  // consumer warps only (the producer has left)
  if (tid == 0) {
    unsigned m = 0;
#pragma unroll
    for (int w = 0; w < NCONS; w++) m = max(m, ring.maxc[w]);
    im.tile_maxc[tile] = m;
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <bool HALF>
__global__ void __launch_bounds__(RENDER_THREADS) k_render_bwd(int W, int H, int gx, const float* __restrict__ bg,
                                                               ImageView im, BinView bin,
                                                               const float4* __restrict__ splat,
                                                               float* __restrict__ grad,
                                                               const float* __restrict__ dL_dpix,
                                                               const float* __restrict__ dL_ddepthpix,
                                                               const float* __restrict__ dL_dmedpix,
                                                               const float* __restrict__ dL_dopacpix,
                                                               const float ddelx_dx, const float ddely_dy) {
  __shared__ __align__(128) Ring ring;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, tx = tile % gx, ty = tile / gx;
  const uint2 range = im.tile_range[tile];
  const int nmax = (int)min(im.tile_maxc[tile], range.y - range.x);
  if (nmax == 0) return;
  const int nb = (nmax + RB - 1) / RB;
  const uint32_t* ids = bin.point_list + range.x;
  ring_init(ring);

  // batch b (counted from the back) covers list positions [lo_b, hi_b), hi_b = nmax - b*RB
  if (warp == NCONS) {
    for (int b = 0; b < nb; b++) {
      const int s = b % NSTAGE;
      if (b >= NSTAGE) {
        if (lane == 0) mbar_wait(&ring.empty[s], (unsigned)((b / NSTAGE - 1) & 1));
        __syncwarp();
      }
      const int hi = nmax - b * RB, lo = max(0, hi - RB);
      stage_gather(ring.buf[s], splat, ids + lo, hi - lo, lane, &ring.full[s]);
    }
    // the consumers wait on every batch, so all copies have landed when they leave; nothing to drain
    return;
  }

  const int sx0 = tx * TILE_X + (warp & 1) * 8, sy0 = ty * TILE_Y + (warp >> 1) * 4;
  const int px = HALF ? sx0 + 4 * (lane >> 4) + (lane & 3) : sx0 + (lane & 7);
  const int py = HALF ? sy0 + ((lane >> 2) & 3) : sy0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float rx0 = (float)sx0, ry0 = (float)sy0;
  const float rx1 = (float)min(sx0 + (HALF ? 3 : 7), W - 1), ry1 = (float)min(sy0 + 3, H - 1);
  const float rbx0 = (float)(sx0 + 4), rbx1 = (float)min(sx0 + 7, W - 1);  // HALF: the right-hand 4x4 block
  const size_t pid = (size_t)py * W + px, HW = (size_t)W * H;

  const float T_final = inside ? im.final_T[pid] : 0.f;
  float T = T_final;
  const int last_contributor = inside ? (int)im.n_contrib[pid] : 0;
  float gp0 = 0.f, gp1 = 0.f, gp2 = 0.f, gD = 0.f, gMed = 0.f, gO = 0.f;
  if (inside) {
    gp0 = dL_dpix[pid]; gp1 = dL_dpix[HW + pid]; gp2 = dL_dpix[2 * HW + pid];
    gD = dL_ddepthpix[pid];
    gMed = dL_dmedpix[pid];  // channel 0 of the [3,H,W] median grad only (backward.cu:482, quirk 4)
    gO = dL_dopacpix[pid];
  }
  float bg_dot = 0.f;  // backward.cu:584-586
  bg_dot += bg[0] * gp0; bg_dot += bg[1] * gp1; bg_dot += bg[2] * gp2;
  const float bg_term = -T_final * bg_dot;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;
  float acc_d = 0.f, last_d = 0.f, acc_o = 0.f, last_o = 0.f, last_alpha = 0.f;
  const int warp_max = (int)__reduce_max_sync(FULL, (unsigned)last_contributor);
  // HALF: traversal bounds of the two 4x4 blocks (a block whose pixels all stopped early skips the tail entries)
  const int maxA = HALF ? (int)__reduce_max_sync(FULL, (lane & 16) ? 0u : (unsigned)last_contributor) : 0;
  const int maxB = HALF ? (int)__reduce_max_sync(FULL, (lane & 16) ? (unsigned)last_contributor : 0u) : 0;
  // which lane publishes which reduced component (see the butterflies below)
  const bool pub = HALF ? (((lane & 1) == 0) || (lane & 7) == 1) : (((lane & 3) == 0) || lane == 1 || lane == 17);
  const int slot = HALF ? (((lane & 1) == 0) ? ((lane & 8) ? 5 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0)
                                             : ((lane & 8) ? 9 : 4))
                        : (((lane & 3) == 0) ? (lane >> 2) : (lane == 1 ? 8 : 9));

  for (int b = 0; b < nb; b++) {
    const int s = b % NSTAGE;
    mbar_wait(&ring.full[s], (unsigned)((b / NSTAGE) & 1));
    const int hi = nmax - b * RB, lo = max(0, hi - RB), cnt = hi - lo;
    const float4* sb = ring.buf[s];
    if (lo < warp_max) {
      for (int base = 0; base < cnt; base += 32) {
        const int j = cnt - 1 - (base + lane);  // lane 0 = farthest entry of this chunk
        bool keep = false;
        if (j >= 0 && lo + j < (HALF ? maxA : warp_max))
          keep = may_touch(sb[j * SPLAT_F4], sb[j * SPLAT_F4 + 1], rx0, ry0, rx1, ry1);
        unsigned mask = __ballot_sync(FULL, keep);
        unsigned maskB = 0;
        if constexpr (HALF) {
          bool keepB = false;
          if (j >= 0 && lo + j < maxB && rbx0 <= rbx1)
            keepB = may_touch(sb[j * SPLAT_F4], sb[j * SPLAT_F4 + 1], rbx0, ry0, rbx1, ry1);
          maskB = __ballot_sync(FULL, keepB);
        }
        while (HALF ? (mask | maskB) : mask) {
          int jj;
          bool has = true;
          if constexpr (HALF) {
            const unsigned mine = (lane & 16) ? maskB : mask;  // each half-warp pops its own next survivor
            has = mine != 0;
            jj = has ? cnt - 1 - (base + __ffs(mine) - 1) : 0;
            mask &= mask - 1;
            maskB &= maskB - 1;
          } else {
            jj = cnt - 1 - (base + __ffs(mask) - 1);
            mask &= mask - 1;
          }
          const float4 q0 = sb[jj * SPLAT_F4], q1 = sb[jj * SPLAT_F4 + 1];
          const float dx = q0.x - pxf, dy = q0.y - pyf;
          const float t1 = __fmul_rn(__fmul_rn(dy, q1.x), dy);
          const float t2 = __fmul_rn(dx, q0.z);
          const float t3 = __fmul_rn(__fmul_rn(dx, q0.w), dy);
          const float power = __fmaf_rn(__fmaf_rn(dx, t2, t1), -0.5f, -t3);
          const float G = expf(power);
          const float alpha = fminf(0.99f, q1.y * G);
          // backward.cu:520-537: entries at or beyond n_contrib, power > 0 and alpha < 1/255 are skipped
          const bool valid = has && (lo + jj < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
          unsigned vb = 0;  // HALF: which half-warps have a contributing pixel
          if constexpr (HALF) {
            vb = __ballot_sync(FULL, valid);
            if (!vb) continue;
          } else {
            if (!__any_sync(FULL, valid)) continue;
          }
          const float4 q2 = sb[jj * SPLAT_F4 + 2];
          float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f, v6 = 0.f, v7 = 0.f, v8 = 0.f, v9 = 0.f;
          if (valid) {
            float inv;  // 1 / (1 - alpha), 1 - alpha in [0.01, 1]: one MUFU.RCP (gradient tolerance 1e-3)
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.f - alpha));
            const float test_T = T * inv;
            const float w = alpha * test_T;
            const float oma = 1.f - last_alpha;
            float dL_dalpha;
            const float c0 = q1.w, c1 = q2.x, c2 = q2.y, c_d = q1.z;
            acc0 = last_alpha * lc0 + oma * acc0; lc0 = c0; dL_dalpha = (c0 - acc0) * gp0; v0 = w * gp0;
            acc1 = last_alpha * lc1 + oma * acc1; lc1 = c1; dL_dalpha += (c1 - acc1) * gp1; v1 = w * gp1;
            acc2 = last_alpha * lc2 + oma * acc2; lc2 = c2; dL_dalpha += (c2 - acc2) * gp2; v2 = w * gp2;
            acc_d = last_alpha * last_d + oma * acc_d; last_d = c_d;
            dL_dalpha += (c_d - acc_d) * gD;
            v3 = w * gD;
            if (test_T > 0.5f && T < 0.5f) v3 += gMed;  // backward.cu:566-569
            acc_o = last_alpha * last_o + oma * acc_o; last_o = 1.f;
            dL_dalpha += (1.f - acc_o) * gO;
            v4 = w * gO;  // direct term, backward.cu:575 (quirk 5)
            dL_dalpha *= test_T;
            T = test_T;
            last_alpha = alpha;
            dL_dalpha += bg_term * inv;  // (-T_final / (1 - alpha)) * bg_dot, backward.cu:584-587
            const float dL_dG = q1.y * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * q0.z - gdy * q0.w;
            const float dG_ddely = -gdy * q1.x - gdx * q0.w;
            v5 = dL_dG * dG_ddelx * ddelx_dx;
            v6 = dL_dG * dG_ddely * ddely_dy;
            const float h = -0.5f * dL_dG;
            v7 = h * gdx * dx;
            v8 = h * gdx * dy;
            v9 = h * gdy * dy;
            v4 += G * dL_dalpha;
          }
          if constexpr (HALF) {
            // 16-lane transposed butterfly per half-warp (12 shuffles): lane bits 8/4/2 pick the component,
            // comp = (bit8 ? 5 : 0) + (bit4 ? 2 : 0) + (bit2 ? 1 : 0) in `c`; components 4 / 9 (by bit8) in `e`
            const bool h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
            float a0 = h8 ? v5 : v0, a1 = h8 ? v6 : v1, a2 = h8 ? v7 : v2, a3 = h8 ? v8 : v3, a4 = h8 ? v9 : v4;
            a0 += __shfl_xor_sync(FULL, h8 ? v0 : v5, 8);
            a1 += __shfl_xor_sync(FULL, h8 ? v1 : v6, 8);
            a2 += __shfl_xor_sync(FULL, h8 ? v2 : v7, 8);
            a3 += __shfl_xor_sync(FULL, h8 ? v3 : v8, 8);
            a4 += __shfl_xor_sync(FULL, h8 ? v4 : v9, 8);
            float b0 = h4 ? a2 : a0, b1 = h4 ? a3 : a1;
            b0 += __shfl_xor_sync(FULL, h4 ? a0 : a2, 4);
            b1 += __shfl_xor_sync(FULL, h4 ? a1 : a3, 4);
            float e = a4 + __shfl_xor_sync(FULL, a4, 4);
            float c = h2 ? b1 : b0;
            c += __shfl_xor_sync(FULL, h2 ? b0 : b1, 2);
            c += __shfl_xor_sync(FULL, c, 1);
            e += __shfl_xor_sync(FULL, e, 2);
            e += __shfl_xor_sync(FULL, e, 1);
            // each half-warp publishes to its own Gaussian (q2.z differs between the halves)
            if (pub && (vb & ((lane & 16) ? 0xffff0000u : 0x0000ffffu)))
              atomicAdd(grad + (size_t)__float_as_int(q2.z) * GRAD_F + slot, ((lane & 1) == 0) ? c : e);
          } else {
            // transposed butterfly: 8 components -> lanes 4c hold the total of component c (c = lane>>2)
            const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
            float a0 = h16 ? v4 : v0, a1 = h16 ? v5 : v1, a2 = h16 ? v6 : v2, a3 = h16 ? v7 : v3;
            a0 += __shfl_xor_sync(FULL, h16 ? v0 : v4, 16);
            a1 += __shfl_xor_sync(FULL, h16 ? v1 : v5, 16);
            a2 += __shfl_xor_sync(FULL, h16 ? v2 : v6, 16);
            a3 += __shfl_xor_sync(FULL, h16 ? v3 : v7, 16);
            float b0 = h8 ? a2 : a0, b1 = h8 ? a3 : a1;
            b0 += __shfl_xor_sync(FULL, h8 ? a0 : a2, 8);
            b1 += __shfl_xor_sync(FULL, h8 ? a1 : a3, 8);
            float c = h4 ? b1 : b0;
            c += __shfl_xor_sync(FULL, h4 ? b0 : b1, 4);
            c += __shfl_xor_sync(FULL, c, 2);
            c += __shfl_xor_sync(FULL, c, 1);
            // the remaining two components: lanes < 16 end with v8's total, lanes >= 16 with v9's
            float e = h16 ? v9 : v8;
            e += __shfl_xor_sync(FULL, h16 ? v8 : v9, 16);
            e += __shfl_xor_sync(FULL, e, 8);
            e += __shfl_xor_sync(FULL, e, 4);
            e += __shfl_xor_sync(FULL, e, 2);
            e += __shfl_xor_sync(FULL, e, 1);
            // accumulator slots: 0-2 colour, 3 depth, 4 opacity, 5-6 mean2D, 7-9 conic (xx, xy, yy)
            if (pub) atomicAdd(grad + (size_t)__float_as_int(q2.z) * GRAD_F + slot, ((lane & 3) == 0) ? c : e);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ring.empty[s]);
  }
}

}  // namespace

// Experimental half-warp compositing: opt-in with GSR_HALFWARP=1, read once per process.
static bool halfwarp_enabled() {
  static const bool on = [] { const char* e = getenv("GSR_HALFWARP"); return e && e[0] == '1'; }();
  return on;
}

void launch_render_fwd(int W, int H, int gx, int gy, ImageView im, BinView b, GeomView g, float* out_color,
                       float* out_depth, float* out_median, float* out_opacity, cudaStream_t st) {
  if (halfwarp_enabled())
    k_render_fwd<true><<<gx * gy, RENDER_THREADS, 0, st>>>(W, H, gx, im, b, g.splat, out_color, out_depth, out_median,
                                                           out_opacity);
  else
    k_render_fwd<false><<<gx * gy, RENDER_THREADS, 0, st>>>(W, H, gx, im, b, g.splat, out_color, out_depth, out_median,
                                                            out_opacity);
}

void launch_render_bwd(int W, int H, int gx, int gy, const float* bg, ImageView im, BinView b, GeomView g,
                       const float* dL_dpix, const float* dL_ddepth, const float* dL_dmedian,
                       const float* dL_dopacity, cudaStream_t st) {
  // d(pixel coordinate)/d(ndc): backward.cu:493-494 (double-precision product rounded to float)
  const float ddelx_dx = (float)(0.5 * W), ddely_dy = (float)(0.5 * H);
  if (halfwarp_enabled())
    k_render_bwd<true><<<gx * gy, RENDER_THREADS, 0, st>>>(W, H, gx, bg, im, b, g.splat, g.grad, dL_dpix, dL_ddepth,
                                                           dL_dmedian, dL_dopacity, ddelx_dx, ddely_dy);
  else
    k_render_bwd<false><<<gx * gy, RENDER_THREADS, 0, st>>>(W, H, gx, bg, im, b, g.splat, g.grad, dL_dpix, dL_ddepth,
                                                            dL_dmedian, dL_dopacity, ddelx_dx, ddely_dy);
}

}  // namespace gsr
