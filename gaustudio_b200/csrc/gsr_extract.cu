// Extraction post-pass on the device: what gaustudio/scripts/extract_pcd.py does to every
// rendered view right after the rasterizer -- and does on the CPU through OpenCV with a D2H/H2D round trip per view.
//
//   k_dilate_minmax    cv2.dilate of the invalid mask + nanmin/nanmax of the surviving depths  (extract_pcd.py:199-214)
//   k_bilateral        cv2.bilateralFilter on the normalised depth, de-normalise, restore     (extract_pcd.py:216-232)
//   k_extract_normals  depth2normal, -1 fill, normal2worldnormal, validity, negation          (extract_pcd.py:325-335)
//   k_fusion_pass      one weighted accumulation pass of normal_fusion                        (extract_pcd.py:117-136,143-165)
//   k_fusion_mean      mean = normalize(sum / weight)                                         (extract_pcd.py:139-140,167-168)
//
// All of it is streaming work: one thread per pixel (or per list entry), 4-13 bytes in, 1-28 bytes out, nothing
// that leaves HBM twice.  No host synchronisation anywhere (the depth range travels through two device words).
#include "gsr_internal.cuh"
#include "../../include/gsr.h"
#include <cmath>

namespace gsr {

// order-preserving float <-> uint map, so that atomicMin/atomicMax on the key order the floats
__device__ __forceinline__ unsigned f2key(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void k_minmax_init(unsigned* scratch) { scratch[0] = 0xffffffffu; scratch[1] = 0u; }

// new_mask = every pixel of the (2r+1)^2 window that lies inside the image is valid (cv2.dilate of the invalid
// mask with its default border, which never contributes); min / max of depth over new_mask & !isnan.
__global__ void k_dilate_minmax(const float* __restrict__ depth, const unsigned char* __restrict__ mask, int W, int H,
                                int r, unsigned char* __restrict__ out_mask, unsigned* __restrict__ scratch) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
  unsigned kmin = 0xffffffffu, kmax = 0u;
  if (u < W && v < H) {
    bool ok = true;
    for (int dy = -r; dy <= r; dy++) {
      const int y = v + dy;
      if (y < 0 || y >= H) continue;
      for (int dx = -r; dx <= r; dx++) {
        const int x = u + dx;
        if (x < 0 || x >= W) continue;
        ok = ok && mask[(size_t)y * W + x] != 0;
      }
    }
    out_mask[(size_t)v * W + u] = ok;
    const float d = depth[(size_t)v * W + u];
    if (ok && d == d) kmin = kmax = f2key(d);
  }
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  __shared__ unsigned smin[8], smax[8];
  const int warp = (threadIdx.y * blockDim.x + threadIdx.x) >> 5, lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31;
  if (lane == 0) { smin[warp] = kmin; smax[warp] = kmax; }
  __syncthreads();
  if (warp == 0) {
    kmin = lane < 8 ? smin[lane] : 0xffffffffu;
    kmax = lane < 8 ? smax[lane] : 0u;
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0 && kmin != 0xffffffffu) { atomicMin(scratch, kmin); atomicMax(scratch + 1, kmax); }
  }
}

__device__ __forceinline__ int reflect101(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// OpenCV's 32F bilateral filter on (depth - min) / (max - min) with invalid pixels at 0: disc support, centre weight
// 1, REFLECT_101 border; then filtered * (max - min) + min in separate roundings like the numpy expression.
__global__ void k_bilateral(const float* __restrict__ depth, const unsigned char* __restrict__ new_mask, int W, int H,
                            int r, float gauss_color, const __grid_constant__ SpaceKernel sk,
                            const unsigned* __restrict__ scratch, float* __restrict__ out) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
  if (u >= W || v >= H) return;
  const size_t pix = (size_t)v * W + u;
  const float d0 = depth[pix];
  const unsigned kmin = scratch[0];
  float res = d0;
  if (kmin != 0xffffffffu && new_mask[pix] && d0 == d0) {
    const float vmin = key2f(kmin), rng = __fsub_rn(key2f(scratch[1]), vmin);
    const float c = __fdiv_rn(__fsub_rn(d0, vmin), rng);
    float sum = c, wsum = 1.0f;
    const int dd = 2 * r + 1;
    for (int dy = -r; dy <= r; dy++) {
      const int y = reflect101(v + dy, H);
      for (int dx = -r; dx <= r; dx++) {
        const float ws = sk.w[(dy + r) * dd + dx + r];
        if (ws == 0.0f) continue;
        const int x = reflect101(u + dx, W);
        const size_t q = (size_t)y * W + x;
        const float dq = depth[q];
        const float nb = (new_mask[q] && dq == dq) ? __fdiv_rn(__fsub_rn(dq, vmin), rng) : 0.0f;
        const float diff = fabsf(nb - c);
        const float w = ws * expf(diff * diff * gauss_color);
        sum = fmaf(nb, w, sum);
        wsum += w;
      }
    }
    res = __fadd_rn(__fmul_rn(__fdiv_rn(sum, wsum), rng), vmin);
  }
  out[pix] = res;
}

__global__ void k_extract_normals(const float* __restrict__ depth, const unsigned char* __restrict__ fg,
                                  const float* __restrict__ opacity, const float* __restrict__ median_depth, int W,
                                  int H, float ifx, float ify, float ox, float oy, const float* __restrict__ rot,
                                  float depth_limit, float opacity_min, float* __restrict__ cam_normals,
                                  float* __restrict__ neg_world, unsigned char* __restrict__ valid) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
  if (u >= W || v >= H) return;
  const size_t pix = (size_t)v * W + u;
  float c0, c1, c2;
  if (!fg[pix] || !cross_normal(depth, u, v, W, H, ifx, ify, ox, oy, 1e-3f, 100000.0f, c0, c1, c2)) c0 = c1 = c2 = -1.f;
  // normal @ inverse(extrinsics[:3,:3]).t() -- applied to the -1 fill as well, as the reference does
  const float w0 = c0 * rot[0] + c1 * rot[3] + c2 * rot[6];
  const float w1 = c0 * rot[1] + c1 * rot[4] + c2 * rot[7];
  const float w2 = c0 * rot[2] + c1 * rot[5] + c2 * rot[8];
  const bool ok = (w0 + w1 + w2 > -3.0f) && median_depth[pix] < depth_limit && opacity[pix] > opacity_min;
  if (cam_normals) { cam_normals[3 * pix] = c0; cam_normals[3 * pix + 1] = c1; cam_normals[3 * pix + 2] = c2; }
  neg_world[3 * pix] = -w0; neg_world[3 * pix + 1] = -w1; neg_world[3 * pix + 2] = -w2;
  valid[pix] = ok;
}

__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__global__ void k_fusion_pass(long long n, const long long* __restrict__ ids, const float* __restrict__ normals,
                              const float* __restrict__ conf, int P, const float* __restrict__ xyz, float tx, float ty,
                              float tz, const float* __restrict__ mean, float thresh, float* __restrict__ sum_normals,
                              float* __restrict__ sum_weights, unsigned char* __restrict__ touched) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long id = ids[i];
  if (id < 0 || id >= P) return;
  const float nx = normals[3 * i], ny = normals[3 * i + 1], nz = normals[3 * i + 2];
  const float vx = tx - xyz[3 * id], vy = ty - xyz[3 * id + 1], vz = tz - xyz[3 * id + 2];
  const float dist = sqrtf(vx * vx + vy * vy + vz * vz);
  const float view_w = fabsf((vx / dist) * nx + (vy / dist) * ny + (vz / dist) * nz);
  const float w = conf[i] * view_w * (1.0f / (dist + 1e-6f));
  if (touched) touched[id] = 1;
  if (mean) {
    const float ex = nx - mean[3 * id], ey = ny - mean[3 * id + 1], ez = nz - mean[3 * id + 2];
    if (!(sqrtf(ex * ex + ey * ey + ez * ez) < thresh)) return;
  }
  red_add(sum_normals + 3 * id, nx * w);
  red_add(sum_normals + 3 * id + 1, ny * w);
  red_add(sum_normals + 3 * id + 2, nz * w);
  red_add(sum_weights + id, w);
}

__global__ void k_fusion_mean(int P, const float* __restrict__ sum_normals, const float* __restrict__ sum_weights,
                              float* __restrict__ mean) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float w = sum_weights[i];
  const float x = sum_normals[3 * i] / w, y = sum_normals[3 * i + 1] / w, z = sum_normals[3 * i + 2] / w;
  const float len = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);  // F.normalize(p=2, eps=1e-12); NaN stays NaN
  mean[3 * i] = x / len; mean[3 * i + 1] = y / len; mean[3 * i + 2] = z / len;
}

void launch_masked_bilateral(const float* depth, const unsigned char* mask, int W, int H, int r, float gauss_color,
                             const SpaceKernel& sk, float* out_depth, unsigned char* out_mask, unsigned* scratch,
                             cudaStream_t st) {
  dim3 blk(32, 8), grd((W + 31) / 32, (H + 7) / 8);
  k_minmax_init<<<1, 1, 0, st>>>(scratch);
  k_dilate_minmax<<<grd, blk, 0, st>>>(depth, mask, W, H, r, out_mask, scratch);
  k_bilateral<<<grd, blk, 0, st>>>(depth, out_mask, W, H, r, gauss_color, sk, scratch, out_depth);
}

void launch_extract_normals(const float* depth, const unsigned char* fg, const float* opacity, const float* median_depth,
                            int W, int H, float fx, float fy, float cx, float cy, const float* rot, float depth_limit,
                            float opacity_min, float* cam_normals, float* neg_world, unsigned char* valid,
                            cudaStream_t st) {
  dim3 blk(32, 8), grd((W + 31) / 32, (H + 7) / 8);
  k_extract_normals<<<grd, blk, 0, st>>>(depth, fg, opacity, median_depth, W, H, 1.0f / fx, 1.0f / fy, -cx / fx, -cy / fy,
                                         rot, depth_limit, opacity_min, cam_normals, neg_world, valid);
}

void launch_fusion_pass(long long n, const long long* ids, const float* normals, const float* conf, int P,
                        const float* xyz, float tx, float ty, float tz, const float* mean, float thresh,
                        float* sum_normals, float* sum_weights, unsigned char* touched, cudaStream_t st) {
  if (n <= 0) return;
  k_fusion_pass<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, ids, normals, conf, P, xyz, tx, ty, tz, mean, thresh,
                                                             sum_normals, sum_weights, touched);
}

void launch_fusion_mean(int P, const float* sum_normals, const float* sum_weights, float* mean, cudaStream_t st) {
  k_fusion_mean<<<(P + 255) / 256, 256, 0, st>>>(P, sum_normals, sum_weights, mean);
}

}  // namespace gsr

// ---- k nearest neighbours on a uniform grid (neighbour search of extract_pcd.py:170-181, there scipy cKDTree on the host)
// One thread per query point (points are sorted by cell, so a warp's queries are neighbours in space and walk the same
// cells).  Rings of cells are scanned outwards; after the cube of radius r cells every point outside it is farther
// than r * cell_size, so the search stops once the k-th best distance is within that bound.
namespace gsr {
constexpr int KNN_MAX_K = 16;
template <int K>
__global__ void __launch_bounds__(128) k_knn_grid(int n, const float* __restrict__ pts, const int* __restrict__ cell_start,
                                                  const float* __restrict__ grid, int* __restrict__ out_index,
                                                  float* __restrict__ out_dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float ox = grid[0], oy = grid[1], oz = grid[2], inv_h = grid[3];
  const int dx = (int)grid[4], dy = (int)grid[5], dz = (int)grid[6];
  const float h = 1.0f / inv_h;
  const float qx = pts[3 * i], qy = pts[3 * i + 1], qz = pts[3 * i + 2];
  const int cx = min(dx - 1, max(0, (int)floorf((qx - ox) * inv_h)));
  const int cy = min(dy - 1, max(0, (int)floorf((qy - oy) * inv_h)));
  const int cz = min(dz - 1, max(0, (int)floorf((qz - oz) * inv_h)));
  float bd[K];
  int bi[K];
#pragma unroll
  for (int k = 0; k < K; k++) { bd[k] = 3.0e38f; bi[k] = -1; }
  const int rmax = max(dx, max(dy, dz));
  for (int r = 0; r <= rmax; r++) {
    // the shell of cells at Chebyshev distance exactly r from the query's cell
    auto scan = [&](int x, int y, int z) {
      const int c = (z * dy + y) * dx + x;
      for (int j = cell_start[c]; j < cell_start[c + 1]; j++) {
        const float ex = pts[3 * j] - qx, ey = pts[3 * j + 1] - qy, ez = pts[3 * j + 2] - qz;
        float d = ex * ex + ey * ey + ez * ez;
        if (d < bd[K - 1] || (d == bd[K - 1] && j < bi[K - 1])) {
          int id = j;
#pragma unroll
          for (int k = 0; k < K; k++) {  // insertion into the ascending list (ties: lower index first)
            const bool before = d < bd[k] || (d == bd[k] && id < bi[k]);
            const float td = before ? bd[k] : d;
            const int ti = before ? bi[k] : id;
            bd[k] = before ? d : bd[k];
            bi[k] = before ? id : bi[k];
            d = td; id = ti;
          }
        }
      }
    };
    for (int z = max(0, cz - r); z <= min(dz - 1, cz + r); z++)
      for (int y = max(0, cy - r); y <= min(dy - 1, cy + r); y++) {
        const bool face = (z == cz - r) || (z == cz + r) || (y == cy - r) || (y == cy + r);
        if (face || r == 0) {
          for (int x = max(0, cx - r); x <= min(dx - 1, cx + r); x++) scan(x, y, z);
        } else {  // interior row of the shell: only its two end cells
          if (cx - r >= 0) scan(cx - r, y, z);
          if (cx + r <= dx - 1) scan(cx + r, y, z);
        }
      }
    const float bound = (float)r * h;  // everything not scanned yet is farther than this
    if (bi[K - 1] >= 0 && bd[K - 1] <= bound * bound) break;
  }
#pragma unroll
  for (int k = 0; k < K; k++) {
    out_index[(size_t)i * K + k] = bi[k];
    out_dist[(size_t)i * K + k] = bi[k] >= 0 ? sqrtf(bd[k]) : __int_as_float(0x7f800000);
  }
}

int launch_knn_grid(int n, int k, const float* pts, const int* cell_start, const float* grid, int* out_index,
                    float* out_dist, cudaStream_t st) {
  const int blocks = (n + 127) / 128;
  switch (k) {
    case 1: k_knn_grid<1><<<blocks, 128, 0, st>>>(n, pts, cell_start, grid, out_index, out_dist); break;
    case 4: k_knn_grid<4><<<blocks, 128, 0, st>>>(n, pts, cell_start, grid, out_index, out_dist); break;
    case 8: k_knn_grid<8><<<blocks, 128, 0, st>>>(n, pts, cell_start, grid, out_index, out_dist); break;
    case 10: k_knn_grid<10><<<blocks, 128, 0, st>>>(n, pts, cell_start, grid, out_index, out_dist); break;
    case 16: k_knn_grid<16><<<blocks, 128, 0, st>>>(n, pts, cell_start, grid, out_index, out_dist); break;
    default: return -1;
  }
  return 0;
}
}  // namespace gsr

// ---------------------------------------------------------------------------------------------------------------
// Fused multi-tensor Adam / AdamW: the reference steps torch.optim.AdamW over five parameter
// groups (gaustudio/pipelines/optimizers/base.py:19-20, configs/vanilla.yaml:30-46).  One launch updates every
// group: per element 16 B read (param, grad, exp_avg, exp_avg_sq) and 12-16 B written (the three states, plus the
// zeroed gradient when zero_grad is fused in) -- a pure HBM stream, float4 wide where the tensors allow it.
// Arithmetic follows torch's single-tensor Adam: p *= 1 - lr*wd (AdamW) | g += wd*p (Adam); m = m + (g-m)(1-b1);
// v = v*b2 + (1-b2) g g; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps).
namespace gsr {

struct AdamTable {
  float* p[GSR_ADAM_MAX_GROUPS]; float* g[GSR_ADAM_MAX_GROUPS]; float* m[GSR_ADAM_MAX_GROUPS]; float* v[GSR_ADAM_MAX_GROUPS];
  float* g2[GSR_ADAM_MAX_GROUPS];  // optional second gradient tensor per group (nullptr: none)
  long long n[GSR_ADAM_MAX_GROUPS];
  float step_size[GSR_ADAM_MAX_GROUPS], decay[GSR_ADAM_MAX_GROUPS];  // lr / bc1; AdamW: 1 - lr*wd, Adam: wd
  unsigned first_block[GSR_ADAM_MAX_GROUPS + 1];
  unsigned char vec4[GSR_ADAM_MAX_GROUPS];
  int groups;
};
constexpr int ADAM_THREADS = 256, ADAM_PER_BLOCK = ADAM_THREADS * 4;

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float step_size, float decay, float b1c,
                                         float b2, float b2c, float inv_bc2s, float eps, float gscale, int decoupled) {
  float grad = g * gscale;
  if (decoupled) p = p * decay; else grad = grad + decay * p;
  m = m + (grad - m) * b1c;
  v = v * b2 + b2c * grad * grad;
  const float denom = sqrtf(v) * inv_bc2s + eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(ADAM_THREADS) k_adam(const __grid_constant__ AdamTable t, float b1c, float b2, float b2c,
                                                       float inv_bc2s, float eps, float gscale, int decoupled, int zero_grad) {
  int gi = 0;
  while (gi + 1 < t.groups && blockIdx.x >= t.first_block[gi + 1]) gi++;
  const long long base = (long long)(blockIdx.x - t.first_block[gi]) * ADAM_PER_BLOCK;
  float* __restrict__ P = t.p[gi]; float* __restrict__ G = t.g[gi]; float* __restrict__ M = t.m[gi]; float* __restrict__ V = t.v[gi];
  float* __restrict__ G2 = t.g2[gi];
  const long long n = t.n[gi];
  const float ss = t.step_size[gi], dc = t.decay[gi];
  const long long i = base + threadIdx.x * 4;
  if (t.vec4[gi] && i + 3 < n) {
    float4 p = *reinterpret_cast<float4*>(P + i), g = *reinterpret_cast<float4*>(G + i);
    float4 m = *reinterpret_cast<float4*>(M + i), v = *reinterpret_cast<float4*>(V + i);
    if (G2) {
      const float4 h = *reinterpret_cast<float4*>(G2 + i);
      g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
      if (zero_grad) *reinterpret_cast<float4*>(G2 + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    adam_one(p.x, g.x, m.x, v.x, ss, dc, b1c, b2, b2c, inv_bc2s, eps, gscale, decoupled);
    adam_one(p.y, g.y, m.y, v.y, ss, dc, b1c, b2, b2c, inv_bc2s, eps, gscale, decoupled);
    adam_one(p.z, g.z, m.z, v.z, ss, dc, b1c, b2, b2c, inv_bc2s, eps, gscale, decoupled);
    adam_one(p.w, g.w, m.w, v.w, ss, dc, b1c, b2, b2c, inv_bc2s, eps, gscale, decoupled);
    *reinterpret_cast<float4*>(P + i) = p; *reinterpret_cast<float4*>(M + i) = m; *reinterpret_cast<float4*>(V + i) = v;
    if (zero_grad) *reinterpret_cast<float4*>(G + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (long long k = i; k < n && k < i + 4; k++) {
      float p = P[k], g = G[k], m = M[k], v = V[k];
      if (G2) { g += G2[k]; if (zero_grad) G2[k] = 0.f; }
      adam_one(p, g, m, v, ss, dc, b1c, b2, b2c, inv_bc2s, eps, gscale, decoupled);
      P[k] = p; M[k] = m; V[k] = v;
      if (zero_grad) G[k] = 0.f;
    }
  }
}

int launch_adam(int n_groups, const gsr_adam_group* groups, double beta1, double beta2, double eps, long long step,
                int decoupled, float grad_scale, int zero_grad, cudaStream_t st) {
  AdamTable t;
  const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
  unsigned blocks = 0;
  int k = 0;
  for (int i = 0; i < n_groups; i++) {
    const gsr_adam_group& q = groups[i];
    if (q.numel <= 0) continue;
    t.p[k] = q.param; t.g[k] = q.grad; t.m[k] = q.exp_avg; t.v[k] = q.exp_avg_sq; t.n[k] = q.numel;
    t.step_size[k] = (float)((double)q.lr / bc1);
    t.decay[k] = decoupled ? (float)(1.0 - (double)q.lr * (double)q.weight_decay) : q.weight_decay;
    t.g2[k] = q.grad2;
    t.vec4[k] = ((reinterpret_cast<size_t>(q.param) | reinterpret_cast<size_t>(q.grad) | reinterpret_cast<size_t>(q.exp_avg) |
                  reinterpret_cast<size_t>(q.exp_avg_sq) | reinterpret_cast<size_t>(q.grad2)) & 15) == 0;
    t.first_block[k] = blocks;
    blocks += (unsigned)((q.numel + ADAM_PER_BLOCK - 1) / ADAM_PER_BLOCK);
    k++;
  }
  t.first_block[k] = blocks;
  t.groups = k;
  if (!k) return 0;
  k_adam<<<blocks, ADAM_THREADS, 0, st>>>(t, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)(1.0 / std::sqrt(bc2)), (float)eps, grad_scale,
                                          decoupled, zero_grad);
  return 0;
}

}  // namespace gsr
