// CPU restatement of the gaustudio differentiable 3DGS rasterizer hot path.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / reference legs of bench.py may load this library; the product
// path (gaustudio_b200/) never does and fails loudly without its CUDA library.
//
// Parity status: the reference ships no tests / golden vectors for this path
// (SURVEY.md §4, §8c).  The oracle is pinned against outputs of the UNMODIFIED
// reference CUDA extension run on a B200 (oracle/build_ref.py -> oracle/_ref,
// fixtures under tests/golden/, generator tests/golden/make_golden_ref.py).
//
// Every function cites the reference file:line it restates; paths are relative
// to /root/reference/submodules/gaustudio-diff-gaussian-rasterization/.
// The code is templated on the scalar type: `float` mirrors the reference's
// arithmetic (including the double-precision ndc2Pix), `double` is used for
// finite-difference validation of the backward restatement.
//
// Build: see oracle/Makefile  (g++ -O2 -fopenmp -ffp-contract=off -mavx2 -mfma)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int TILE = 16;  // config.h:16-17 (BLOCK_X, BLOCK_Y) -- observable behaviour

// auxiliary.h:22-39
const double kSH_C0 = 0.28209479177387814;
const double kSH_C1 = 0.4886025119029199;
const double kSH_C2[5] = {1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
                          -1.0925484305920792, 0.5462742152960396};
const double kSH_C3[7] = {-0.5900435899266435, 2.890611442640554, -0.4570457994644658,
                          0.3731763325901154,  -0.4570457994644658, 1.445305721320277,
                          -0.5900435899266435};

template <typename real> inline real C_(double v) {
  // the reference declares the SH constants as float (auxiliary.h:22-39)
  return (real)(std::is_same<real, float>::value ? (double)(float)v : v);
}

// CUDA float->int conversion saturates and maps NaN to 0 (cvt.rzi.s32.f32).
template <typename real> inline int f2i(real v) {
  if (v != v) return 0;
  if (v >= (real)2147483647.0) return std::numeric_limits<int>::max();
  if (v <= (real)-2147483648.0) return std::numeric_limits<int>::min();
  return (int)v;
}

// column-major 3x3 like glm: m[c][r]; product order of
// third_party/glm/glm/detail/type_mat3x3.inl:486-518
template <typename real> struct M3 {
  real m[3][3];
};
template <typename real> inline M3<real> mul(const M3<real>& A, const M3<real>& B) {
  M3<real> R;
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++)
      R.m[c][r] = A.m[0][r] * B.m[c][0] + A.m[1][r] * B.m[c][1] + A.m[2][r] * B.m[c][2];
  return R;
}
template <typename real> inline M3<real> tr(const M3<real>& A) {
  M3<real> R;
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) R.m[c][r] = A.m[r][c];
  return R;
}

template <typename real> struct State {
  int P = 0, D = 0, M = 0, W = 0, H = 0, gx = 0, gy = 0;
  int64_t R = 0;
  // GeometryState (rasterizer_impl.h:33-47)
  std::vector<real> depths, cov3D, rgb, means2D, conic_opacity;
  std::vector<uint8_t> clamped;
  std::vector<int> radii;
  std::vector<uint32_t> tiles_touched;
  // BinningState / ImageState (rasterizer_impl.h:49-66)
  std::vector<uint32_t> point_list;
  std::vector<uint32_t> ranges;  // 2 per tile
  std::vector<real> final_T;
  std::vector<uint32_t> n_contrib;
};

// auxiliary.h:58-77
template <typename real> inline void xf4x3(const real* p, const real* m, real* o) {
  o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
  o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
  o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
template <typename real> inline void xf4x4(const real* p, const real* m, real* o) {
  o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
  o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
  o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
  o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

// auxiliary.h:41-44 -- evaluated in double because of the 1.0 / 0.5 literals
template <typename real> inline real ndc2pix(real v, int S) {
  return (real)((((double)v + 1.0) * (double)S - 1.0) * 0.5);
}

// auxiliary.h:46-56
template <typename real>
inline void get_rect(real px, real py, int max_radius, int gx, int gy, int* rmin, int* rmax) {
  const real r = (real)max_radius;
  rmin[0] = std::min(gx, std::max(0, f2i<real>((px - r) / (real)TILE)));
  rmin[1] = std::min(gy, std::max(0, f2i<real>((py - r) / (real)TILE)));
  rmax[0] = std::min(gx, std::max(0, f2i<real>((px + r + (real)(TILE - 1)) / (real)TILE)));
  rmax[1] = std::min(gy, std::max(0, f2i<real>((py + r + (real)(TILE - 1)) / (real)TILE)));
}

// forward.cu:118-152 (un-normalised quaternion: quirk 2)
template <typename real>
inline void cov3d_from_scale_rot(const real* s, real mod, const real* q, real* out, M3<real>* Mout,
                                 M3<real>* Rout) {
  M3<real> S;
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) S.m[c][r] = (real)0;
  S.m[0][0] = mod * s[0];
  S.m[1][1] = mod * s[1];
  S.m[2][2] = mod * s[2];
  const real r = q[0], x = q[1], y = q[2], z = q[3];
  M3<real> R;
  R.m[0][0] = (real)1 - (real)2 * (y * y + z * z);
  R.m[0][1] = (real)2 * (x * y - r * z);
  R.m[0][2] = (real)2 * (x * z + r * y);
  R.m[1][0] = (real)2 * (x * y + r * z);
  R.m[1][1] = (real)1 - (real)2 * (x * x + z * z);
  R.m[1][2] = (real)2 * (y * z - r * x);
  R.m[2][0] = (real)2 * (x * z - r * y);
  R.m[2][1] = (real)2 * (y * z + r * x);
  R.m[2][2] = (real)1 - (real)2 * (x * x + y * y);
  M3<real> Mm = mul(S, R);
  M3<real> Sg = mul(tr(Mm), Mm);
  out[0] = Sg.m[0][0];
  out[1] = Sg.m[0][1];
  out[2] = Sg.m[0][2];
  out[3] = Sg.m[1][1];
  out[4] = Sg.m[1][2];
  out[5] = Sg.m[2][2];
  if (Mout) *Mout = Mm;
  if (Rout) *Rout = R;
}

// shared by forward.cu:74-113 and backward.cu:144-274
template <typename real> struct Cov2DCtx {
  real t[3];
  real txtz, tytz, limx, limy;
  M3<real> J, Wm, T, Vrk, cov;
};
template <typename real>
inline void cov2d(const real* mean, real fx, real fy, real tanx, real tany, const real* c3,
                  const real* view, Cov2DCtx<real>& o) {
  xf4x3(mean, view, o.t);
  o.limx = (real)1.3f * tanx;
  o.limy = (real)1.3f * tany;
  if (std::is_same<real, double>::value) {
    o.limx = (real)1.3 * tanx;
    o.limy = (real)1.3 * tany;
  }
  o.txtz = o.t[0] / o.t[2];
  o.tytz = o.t[1] / o.t[2];
  o.t[0] = std::min(o.limx, std::max(-o.limx, o.txtz)) * o.t[2];
  o.t[1] = std::min(o.limy, std::max(-o.limy, o.tytz)) * o.t[2];
  const real tz = o.t[2];
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) o.J.m[c][r] = (real)0;
  o.J.m[0][0] = fx / tz;
  o.J.m[0][2] = -(fx * o.t[0]) / (tz * tz);
  o.J.m[1][1] = fy / tz;
  o.J.m[1][2] = -(fy * o.t[1]) / (tz * tz);
  o.Wm.m[0][0] = view[0];
  o.Wm.m[0][1] = view[4];
  o.Wm.m[0][2] = view[8];
  o.Wm.m[1][0] = view[1];
  o.Wm.m[1][1] = view[5];
  o.Wm.m[1][2] = view[9];
  o.Wm.m[2][0] = view[2];
  o.Wm.m[2][1] = view[6];
  o.Wm.m[2][2] = view[10];
  o.T = mul(o.Wm, o.J);
  o.Vrk.m[0][0] = c3[0];
  o.Vrk.m[0][1] = c3[1];
  o.Vrk.m[0][2] = c3[2];
  o.Vrk.m[1][0] = c3[1];
  o.Vrk.m[1][1] = c3[3];
  o.Vrk.m[1][2] = c3[4];
  o.Vrk.m[2][0] = c3[2];
  o.Vrk.m[2][1] = c3[4];
  o.Vrk.m[2][2] = c3[5];
  o.cov = mul(mul(tr(o.T), tr(o.Vrk)), o.T);
  o.cov.m[0][0] += (real)0.3f;
  o.cov.m[1][1] += (real)0.3f;
  if (std::is_same<real, double>::value) {
    o.cov.m[0][0] += (real)0.3 - (real)0.3f;
    o.cov.m[1][1] += (real)0.3 - (real)0.3f;
  }
}

// forward.cu:20-71
template <typename real>
inline void sh_to_rgb(int deg, int M, const real* pos, const real* campos, const real* sh /*[M][3]*/,
                      real* rgb, uint8_t* clamped) {
  real d[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
  real len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  const real x = d[0] / len, y = d[1] / len, z = d[2] / len;
  const real C0 = C_<real>(kSH_C0), C1 = C_<real>(kSH_C1);
  for (int c = 0; c < 3; c++) {
    auto S = [&](int k) { return sh[3 * k + c]; };
    real res = C0 * S(0);
    if (deg > 0) {
      res = res - C1 * y * S(1) + C1 * z * S(2) - C1 * x * S(3);
      if (deg > 1) {
        const real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        res = res + C_<real>(kSH_C2[0]) * xy * S(4) + C_<real>(kSH_C2[1]) * yz * S(5) +
              C_<real>(kSH_C2[2]) * ((real)2 * zz - xx - yy) * S(6) + C_<real>(kSH_C2[3]) * xz * S(7) +
              C_<real>(kSH_C2[4]) * (xx - yy) * S(8);
        if (deg > 2) {
          res = res + C_<real>(kSH_C3[0]) * y * ((real)3 * xx - yy) * S(9) +
                C_<real>(kSH_C3[1]) * xy * z * S(10) +
                C_<real>(kSH_C3[2]) * y * ((real)4 * zz - xx - yy) * S(11) +
                C_<real>(kSH_C3[3]) * z * ((real)2 * zz - (real)3 * xx - (real)3 * yy) * S(12) +
                C_<real>(kSH_C3[4]) * x * ((real)4 * zz - xx - yy) * S(13) +
                C_<real>(kSH_C3[5]) * z * (xx - yy) * S(14) +
                C_<real>(kSH_C3[6]) * x * (xx - (real)3 * yy) * S(15);
        }
      }
    }
    res += (real)0.5;
    clamped[c] = res < 0;
    rgb[c] = std::max(res, (real)0);
  }
}

// --------------------------------------------------------------------------
// Forward: forward.cu:155-256 (K1), rasterizer_impl.cu:70-138,278-321 (binning),
// forward.cu:261-397 (K4)
// --------------------------------------------------------------------------
template <typename real>
int64_t forward(State<real>& st, int P, int D, int M, int W, int H, const real* means3D, const real* shs,
                const real* colors_precomp, const real* opacities, const real* scales, real scale_modifier,
                const real* rotations, const real* cov3D_precomp, const real* view, const real* proj,
                const real* campos, real tanx, real tany, real* out_color, real* out_depth, real* out_median,
                real* out_opacity, int* out_radii) {
  st.P = P; st.D = D; st.M = M; st.W = W; st.H = H;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  st.gx = gx; st.gy = gy;
  const real fy = (real)H / ((real)2 * tany), fx = (real)W / ((real)2 * tanx);  // rasterizer_impl.cu:225-226
  st.depths.assign(P, 0); st.cov3D.assign((size_t)P * 6, 0); st.rgb.assign((size_t)P * 3, 0);
  st.means2D.assign((size_t)P * 2, 0); st.conic_opacity.assign((size_t)P * 4, 0);
  st.clamped.assign((size_t)P * 3, 0); st.radii.assign(P, 0); st.tiles_touched.assign(P, 0);

#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    const real* p = means3D + 3 * (size_t)i;
    real pv[3];
    xf4x3(p, view, pv);
    if (pv[2] <= (real)0.2f) continue;  // auxiliary.h:154 (near plane only)
    real ph[4];
    xf4x4(p, proj, ph);
    const real pw = (real)1 / (ph[3] + (real)0.0000001f);
    const real pp[2] = {ph[0] * pw, ph[1] * pw};
    const real* c3;
    if (cov3D_precomp) {
      c3 = cov3D_precomp + 6 * (size_t)i;
    } else {
      cov3d_from_scale_rot<real>(scales + 3 * (size_t)i, scale_modifier, rotations + 4 * (size_t)i,
                                 &st.cov3D[6 * (size_t)i], nullptr, nullptr);
      c3 = &st.cov3D[6 * (size_t)i];
    }
    Cov2DCtx<real> cc;
    cov2d<real>(p, fx, fy, tanx, tany, c3, view, cc);
    const real a = cc.cov.m[0][0], b = cc.cov.m[0][1], c = cc.cov.m[1][1];
    const real det = a * c - b * b;
    if (det == (real)0) continue;
    const real det_inv = (real)1 / det;
    const real conic[3] = {c * det_inv, -b * det_inv, a * det_inv};
    const real mid = (real)0.5 * (a + c);
    const real sq = std::sqrt(std::max((real)0.1f, mid * mid - det));
    const real l1 = mid + sq, l2 = mid - sq;
    const real my_radius = std::ceil((real)3 * std::sqrt(std::max(l1, l2)));
    const real px = ndc2pix<real>(pp[0], W), py = ndc2pix<real>(pp[1], H);
    int rmin[2], rmax[2];
    get_rect<real>(px, py, f2i<real>(my_radius), gx, gy, rmin, rmax);
    if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;
    if (!colors_precomp)
      sh_to_rgb<real>(D, M, p, campos, shs + (size_t)i * M * 3, &st.rgb[3 * (size_t)i], &st.clamped[3 * (size_t)i]);
    st.depths[i] = pv[2];
    st.radii[i] = f2i<real>(my_radius);
    st.means2D[2 * (size_t)i] = px;
    st.means2D[2 * (size_t)i + 1] = py;
    st.conic_opacity[4 * (size_t)i + 0] = conic[0];
    st.conic_opacity[4 * (size_t)i + 1] = conic[1];
    st.conic_opacity[4 * (size_t)i + 2] = conic[2];
    st.conic_opacity[4 * (size_t)i + 3] = opacities[i];
    st.tiles_touched[i] = (uint32_t)((rmax[1] - rmin[1]) * (rmax[0] - rmin[0]));
  }
  if (out_radii) std::memcpy(out_radii, st.radii.data(), sizeof(int) * P);

  // ---- binning: duplicateWithKeys + stable SortPairs + identifyTileRanges.
  // Emit order is ascending Gaussian index, so "stable sort by (tile, depth bits)"
  // == sort by (tile, depth bits, Gaussian index).
  const int T = gx * gy;
  std::vector<uint32_t> tcount(T + 1, 0);
  int64_t R = 0;
  for (int i = 0; i < P; i++) R += st.tiles_touched[i];
  st.R = R;
  struct Ent { uint32_t key; uint32_t idx; };
  std::vector<Ent> ents((size_t)R);
  {
    // count per tile
    for (int i = 0; i < P; i++) {
      if (st.radii[i] <= 0) continue;
      int rmin[2], rmax[2];
      get_rect<real>(st.means2D[2 * (size_t)i], st.means2D[2 * (size_t)i + 1], st.radii[i], gx, gy, rmin, rmax);
      for (int y = rmin[1]; y < rmax[1]; y++)
        for (int x = rmin[0]; x < rmax[0]; x++) tcount[y * gx + x + 1]++;
    }
    for (int t = 0; t < T; t++) tcount[t + 1] += tcount[t];
    std::vector<uint32_t> cur(tcount.begin(), tcount.end() - 1);
    for (int i = 0; i < P; i++) {
      if (st.radii[i] <= 0) continue;
      int rmin[2], rmax[2];
      get_rect<real>(st.means2D[2 * (size_t)i], st.means2D[2 * (size_t)i + 1], st.radii[i], gx, gy, rmin, rmax);
      // the 32-bit sort key is the bit pattern of the *float* depth (rasterizer_impl.cu:102)
      float df = (float)st.depths[i];
      uint32_t bits;
      std::memcpy(&bits, &df, 4);
      for (int y = rmin[1]; y < rmax[1]; y++)
        for (int x = rmin[0]; x < rmax[0]; x++) ents[cur[y * gx + x]++] = Ent{bits, (uint32_t)i};
    }
  }
  st.ranges.assign((size_t)T * 2, 0);
  st.point_list.resize((size_t)R);
#pragma omp parallel for schedule(dynamic, 16)
  for (int t = 0; t < T; t++) {
    const uint32_t b = tcount[t], e = tcount[t + 1];
    if (e > b) {
      std::stable_sort(ents.begin() + b, ents.begin() + e, [](const Ent& l, const Ent& r) { return l.key < r.key; });
      st.ranges[2 * (size_t)t] = b;       // identifyTileRanges; empty tiles stay (0,0) (rasterizer_impl.cu:313)
      st.ranges[2 * (size_t)t + 1] = e;
      for (uint32_t k = b; k < e; k++) st.point_list[k] = ents[k].idx;
    }
  }

  // ---- per-pixel blend, forward.cu:261-397
  st.final_T.assign((size_t)W * H, 0);
  st.n_contrib.assign((size_t)W * H, 0);
  const size_t HW = (size_t)W * H;
  const real kAlphaMin = std::is_same<real, float>::value ? (real)(1.0f / 255.0f) : (real)(1.0 / 255.0);
  const real kTMin = std::is_same<real, float>::value ? (real)0.0001f : (real)0.0001;
  const real kAlphaMax = std::is_same<real, float>::value ? (real)0.99f : (real)0.99;
#pragma omp parallel for schedule(dynamic, 4)
  for (int t = 0; t < T; t++) {
    const int tx = t % gx, ty = t / gx;
    const uint32_t rb = st.ranges[2 * (size_t)t], re = st.ranges[2 * (size_t)t + 1];
    real Tr[TILE * TILE], C[3][TILE * TILE], Dp[TILE * TILE], medD[TILE * TILE], medW[TILE * TILE], medI[TILE * TILE];
    uint32_t last[TILE * TILE];
    bool done[TILE * TILE];
    int live = 0;
    for (int k = 0; k < TILE * TILE; k++) {
      const int x = tx * TILE + (k % TILE), y = ty * TILE + (k / TILE);
      Tr[k] = 1; C[0][k] = C[1][k] = C[2][k] = 0; Dp[k] = 0;
      medD[k] = (real)15.0f; medW[k] = 0; medI[k] = 0; last[k] = 0;  // forward.cu:310-312
      done[k] = !(x < W && y < H);
      live += !done[k];
    }
    uint32_t contributor = 0;
    for (uint32_t e = rb; e < re && live > 0; e++) {
      contributor++;
      const uint32_t g = st.point_list[e];
      const real gxm = st.means2D[2 * (size_t)g], gym = st.means2D[2 * (size_t)g + 1];
      const real cA = st.conic_opacity[4 * (size_t)g], cB = st.conic_opacity[4 * (size_t)g + 1],
                 cC = st.conic_opacity[4 * (size_t)g + 2], op = st.conic_opacity[4 * (size_t)g + 3];
      const real* col = colors_precomp ? colors_precomp + 3 * (size_t)g : &st.rgb[3 * (size_t)g];
      const real dep = st.depths[g];
      for (int k = 0; k < TILE * TILE; k++) {
        if (done[k]) continue;
        const real pxf = (real)(tx * TILE + (k % TILE)), pyf = (real)(ty * TILE + (k / TILE));
        const real dx = gxm - pxf, dy = gym - pyf;
        // contraction pattern of the reference SASS
        const real t1 = (dy * cC) * dy, t2 = dx * cA, t3 = (dx * cB) * dy;
        const real s = std::fma(dx, t2, t1);
        const real power = std::fma(s, (real)-0.5, -t3);
        if (power > (real)0) continue;
        const real alpha = std::min(kAlphaMax, op * std::exp(power));
        if (alpha < kAlphaMin) continue;
        const real testT = Tr[k] * ((real)1 - alpha);
        if (testT < kTMin) { done[k] = true; live--; continue; }
        for (int ch = 0; ch < 3; ch++) C[ch][k] = std::fma(Tr[k], alpha * col[ch], C[ch][k]);
        Dp[k] = std::fma(Tr[k], alpha * dep, Dp[k]);
        if (Tr[k] > (real)0.5 && testT < (real)0.5) { medD[k] = dep; medW[k] = alpha * Tr[k]; medI[k] = (real)g; }
        Tr[k] = testT;
        last[k] = contributor;
      }
    }
    for (int k = 0; k < TILE * TILE; k++) {
      const int x = tx * TILE + (k % TILE), y = ty * TILE + (k / TILE);
      if (!(x < W && y < H)) continue;
      const size_t pid = (size_t)y * W + x;
      st.final_T[pid] = Tr[k];
      st.n_contrib[pid] = last[k];
      for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pid] = C[ch][k];  // no bg blend (quirk 1)
      out_depth[pid] = Dp[k];
      out_median[pid] = medD[k];
      out_median[HW + pid] = medW[k];
      out_median[2 * HW + pid] = medI[k];
      out_opacity[pid] = (real)1 - Tr[k];
    }
  }
  return R;
}

// --------------------------------------------------------------------------
// Backward: backward.cu:415-610 (K5), 144-274 (K6), 346-412 + 20-139 + 278-341 (K7)
// --------------------------------------------------------------------------
template <typename real>
void backward(State<real>& st, const real* bg, const real* means3D, const real* shs, const real* colors_precomp,
              const real* scales, real scale_modifier, const real* rotations, const real* cov3D_precomp,
              const real* view, const real* proj, const real* campos, real tanx, real tany, const real* dL_dpix,
              const real* dL_ddepthpix, const real* dL_dmedian, const real* dL_dopacitypix, real* dL_dmean2D /*[P,3]*/,
              real* dL_dconic /*[P,4]*/, real* dL_dopacity /*[P]*/, real* dL_dcolor /*[P,3]*/,
              real* dL_ddepth /*[P]*/, real* dL_dmean3D /*[P,3]*/, real* dL_dcov3D /*[P,6]*/,
              real* dL_dsh /*[P,M,3]*/, real* dL_dscale /*[P,3]*/, real* dL_drot /*[P,4]*/) {
  const int P = st.P, D = st.D, M = st.M, W = st.W, H = st.H, gx = st.gx, gy = st.gy;
  const size_t HW = (size_t)W * H;
  const int T = gx * gy;
  const real fy = (real)H / ((real)2 * tany), fx = (real)W / ((real)2 * tanx);
  std::fill(dL_dmean2D, dL_dmean2D + (size_t)P * 3, (real)0);
  std::fill(dL_dconic, dL_dconic + (size_t)P * 4, (real)0);
  std::fill(dL_dopacity, dL_dopacity + P, (real)0);
  std::fill(dL_dcolor, dL_dcolor + (size_t)P * 3, (real)0);
  std::fill(dL_ddepth, dL_ddepth + P, (real)0);
  std::fill(dL_dmean3D, dL_dmean3D + (size_t)P * 3, (real)0);
  std::fill(dL_dcov3D, dL_dcov3D + (size_t)P * 6, (real)0);
  std::fill(dL_dsh, dL_dsh + (size_t)P * M * 3, (real)0);
  std::fill(dL_dscale, dL_dscale + (size_t)P * 3, (real)0);
  std::fill(dL_drot, dL_drot + (size_t)P * 4, (real)0);

  const real kAlphaMin = std::is_same<real, float>::value ? (real)(1.0f / 255.0f) : (real)(1.0 / 255.0);
  const real kAlphaMax = std::is_same<real, float>::value ? (real)0.99f : (real)0.99;
  const real ddelx_dx = (real)(0.5 * W), ddely_dy = (real)(0.5 * H);  // backward.cu:493-494

  // K5. Per tile the per-instance sums are accumulated locally (the reference uses global
  // atomics, backward.cu:559-607; summation order is unspecified there).
#pragma omp parallel for schedule(dynamic, 4)
  for (int t = 0; t < T; t++) {
    const int tx = t % gx, ty = t / gx;
    const uint32_t rb = st.ranges[2 * (size_t)t], re = st.ranges[2 * (size_t)t + 1];
    const uint32_t n = re - rb;
    if (n == 0) continue;
    std::vector<real> acc((size_t)n * 10, (real)0);
    for (int k = 0; k < TILE * TILE; k++) {
      const int x = tx * TILE + (k % TILE), y = ty * TILE + (k / TILE);
      if (!(x < W && y < H)) continue;
      const size_t pid = (size_t)y * W + x;
      const real pxf = (real)x, pyf = (real)y;
      const real T_final = st.final_T[pid];
      real Tc = T_final;
      const uint32_t lastc = st.n_contrib[pid];
      const real gpix[3] = {dL_dpix[pid], dL_dpix[HW + pid], dL_dpix[2 * HW + pid]};
      const real gD = dL_ddepthpix[pid], gMed = dL_dmedian[pid], gO = dL_dopacitypix[pid];  // channel 0 only: quirk 4
      real accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0};
      real accum_depth = 0, last_depth = 0, accum_op = 0, last_op = 0, last_alpha = 0;
      real bg_dot = 0;
      for (int c = 0; c < 3; c++) bg_dot += bg[c] * gpix[c];
      // traverse back to front; entries at 0-based position >= n_contrib are skipped (backward.cu:520-522)
      for (uint32_t pos = std::min(n, lastc); pos-- > 0;) {
        const uint32_t g = st.point_list[rb + pos];
        const real dx = st.means2D[2 * (size_t)g] - pxf, dy = st.means2D[2 * (size_t)g + 1] - pyf;
        const real cA = st.conic_opacity[4 * (size_t)g], cB = st.conic_opacity[4 * (size_t)g + 1],
                   cC = st.conic_opacity[4 * (size_t)g + 2], op = st.conic_opacity[4 * (size_t)g + 3];
        const real t1 = (dy * cC) * dy, t2 = dx * cA, t3 = (dx * cB) * dy;
        const real power = std::fma(std::fma(dx, t2, t1), (real)-0.5, -t3);
        if (power > (real)0) continue;
        const real G = std::exp(power);
        const real alpha = std::min(kAlphaMax, op * G);
        if (alpha < kAlphaMin) continue;
        const real Tb = Tc / ((real)1 - alpha);
        const real w = alpha * Tb;
        real* a = &acc[(size_t)pos * 10];
        real dL_dalpha = 0;
        const real* col = colors_precomp ? colors_precomp + 3 * (size_t)g : &st.rgb[3 * (size_t)g];
        for (int ch = 0; ch < 3; ch++) {
          accum_rec[ch] = last_alpha * last_color[ch] + ((real)1 - last_alpha) * accum_rec[ch];
          last_color[ch] = col[ch];
          dL_dalpha += (col[ch] - accum_rec[ch]) * gpix[ch];
          a[0 + ch] += w * gpix[ch];
        }
        const real dep = st.depths[g];
        accum_depth = last_alpha * last_depth + ((real)1 - last_alpha) * accum_depth;
        last_depth = dep;
        dL_dalpha += (dep - accum_depth) * gD;
        a[3] += w * gD;
        if (Tb > (real)0.5 && Tc < (real)0.5) a[3] += gMed;  // backward.cu:566-569
        accum_op = last_alpha * last_op + ((real)1 - last_alpha) * accum_op;
        last_op = 1;
        dL_dalpha += ((real)1 - accum_op) * gO;
        a[4] += w * gO;  // direct opacity term (quirk 5), backward.cu:575
        dL_dalpha *= Tb;
        Tc = Tb;
        last_alpha = alpha;
        dL_dalpha += (-T_final / ((real)1 - alpha)) * bg_dot;  // backward.cu:584-587
        const real dL_dG = op * dL_dalpha;
        const real gdx = G * dx, gdy = G * dy;
        const real dG_ddelx = -gdx * cA - gdy * cB;
        const real dG_ddely = -gdy * cC - gdx * cB;
        a[5] += dL_dG * dG_ddelx * ddelx_dx;
        a[6] += dL_dG * dG_ddely * ddely_dy;
        a[7] += (real)-0.5 * gdx * dx * dL_dG;
        a[8] += (real)-0.5 * gdx * dy * dL_dG;
        a[9] += (real)-0.5 * gdy * dy * dL_dG;
        a[4] += G * dL_dalpha;
      }
    }
    for (uint32_t pos = 0; pos < n; pos++) {
      const uint32_t g = st.point_list[rb + pos];
      const real* a = &acc[(size_t)pos * 10];
      real* dst[10] = {&dL_dcolor[3 * (size_t)g],     &dL_dcolor[3 * (size_t)g + 1], &dL_dcolor[3 * (size_t)g + 2],
                       &dL_ddepth[g],                 &dL_dopacity[g],               &dL_dmean2D[3 * (size_t)g],
                       &dL_dmean2D[3 * (size_t)g + 1], &dL_dconic[4 * (size_t)g],     &dL_dconic[4 * (size_t)g + 1],
                       &dL_dconic[4 * (size_t)g + 3]};
      for (int q = 0; q < 10; q++) {
        if (a[q] == (real)0) continue;
#pragma omp atomic
        *dst[q] += a[q];
      }
    }
  }

  // K6 then K7 (K6 assigns dL_dmeans, K7 accumulates: quirk 7)
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    if (!(st.radii[i] > 0)) continue;  // quirk 8
    const real* mean = means3D + 3 * (size_t)i;
    const real* c3 = cov3D_precomp ? cov3D_precomp + 6 * (size_t)i : &st.cov3D[6 * (size_t)i];
    // ---- K6: backward.cu:144-274
    Cov2DCtx<real> cc;
    cov2d<real>(mean, fx, fy, tanx, tany, c3, view, cc);
    const real gcon[3] = {dL_dconic[4 * (size_t)i], dL_dconic[4 * (size_t)i + 1], dL_dconic[4 * (size_t)i + 3]};
    const real xm = (cc.txtz < -cc.limx || cc.txtz > cc.limx) ? (real)0 : (real)1;
    const real ym = (cc.tytz < -cc.limy || cc.tytz > cc.limy) ? (real)0 : (real)1;
    const real a = cc.cov.m[0][0], b = cc.cov.m[0][1], c = cc.cov.m[1][1];
    const real denom = a * c - b * b;
    real da = 0, db = 0, dc = 0;
    const real d2inv = (real)1 / ((denom * denom) + (real)0.0000001f);
    real* dcv = dL_dcov3D + 6 * (size_t)i;
    const M3<real>& Tm = cc.T;
    if (d2inv != (real)0) {
      da = d2inv * (-c * c * gcon[0] + (real)2 * b * c * gcon[1] + (denom - a * c) * gcon[2]);
      dc = d2inv * (-a * a * gcon[2] + (real)2 * a * b * gcon[1] + (denom - a * c) * gcon[0]);
      db = d2inv * (real)2 * (b * c * gcon[0] - (denom + (real)2 * b * b) * gcon[1] + a * b * gcon[2]);
      dcv[0] = Tm.m[0][0] * Tm.m[0][0] * da + Tm.m[0][0] * Tm.m[1][0] * db + Tm.m[1][0] * Tm.m[1][0] * dc;
      dcv[3] = Tm.m[0][1] * Tm.m[0][1] * da + Tm.m[0][1] * Tm.m[1][1] * db + Tm.m[1][1] * Tm.m[1][1] * dc;
      dcv[5] = Tm.m[0][2] * Tm.m[0][2] * da + Tm.m[0][2] * Tm.m[1][2] * db + Tm.m[1][2] * Tm.m[1][2] * dc;
      dcv[1] = (real)2 * Tm.m[0][0] * Tm.m[0][1] * da + (Tm.m[0][0] * Tm.m[1][1] + Tm.m[0][1] * Tm.m[1][0]) * db +
               (real)2 * Tm.m[1][0] * Tm.m[1][1] * dc;
      dcv[2] = (real)2 * Tm.m[0][0] * Tm.m[0][2] * da + (Tm.m[0][0] * Tm.m[1][2] + Tm.m[0][2] * Tm.m[1][0]) * db +
               (real)2 * Tm.m[1][0] * Tm.m[1][2] * dc;
      dcv[4] = (real)2 * Tm.m[0][2] * Tm.m[0][1] * da + (Tm.m[0][1] * Tm.m[1][2] + Tm.m[0][2] * Tm.m[1][1]) * db +
               (real)2 * Tm.m[1][1] * Tm.m[1][2] * dc;
    } else {
      for (int q = 0; q < 6; q++) dcv[q] = 0;
    }
    const M3<real>& V = cc.Vrk;
    real dT[2][3];
    for (int j = 0; j < 3; j++) {
      const real r0 = Tm.m[0][0] * V.m[j][0] + Tm.m[0][1] * V.m[j][1] + Tm.m[0][2] * V.m[j][2];
      const real r1 = Tm.m[1][0] * V.m[j][0] + Tm.m[1][1] * V.m[j][1] + Tm.m[1][2] * V.m[j][2];
      dT[0][j] = (real)2 * r0 * da + r1 * db;
      dT[1][j] = (real)2 * r1 * dc + r0 * db;
    }
    const M3<real>& Wm = cc.Wm;
    const real dJ00 = Wm.m[0][0] * dT[0][0] + Wm.m[0][1] * dT[0][1] + Wm.m[0][2] * dT[0][2];
    const real dJ02 = Wm.m[2][0] * dT[0][0] + Wm.m[2][1] * dT[0][1] + Wm.m[2][2] * dT[0][2];
    const real dJ11 = Wm.m[1][0] * dT[1][0] + Wm.m[1][1] * dT[1][1] + Wm.m[1][2] * dT[1][2];
    const real dJ12 = Wm.m[2][0] * dT[1][0] + Wm.m[2][1] * dT[1][1] + Wm.m[2][2] * dT[1][2];
    const real tz = (real)1 / cc.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const real dtx = xm * -fx * tz2 * dJ02;
    const real dty = ym * -fy * tz2 * dJ12;
    const real dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + ((real)2 * fx * cc.t[0]) * tz3 * dJ02 +
                     ((real)2 * fy * cc.t[1]) * tz3 * dJ12;
    real dmean[3] = {view[0] * dtx + view[1] * dty + view[2] * dtz, view[4] * dtx + view[5] * dty + view[6] * dtz,
                     view[8] * dtx + view[9] * dty + view[10] * dtz};  // assignment (backward.cu:273)

    // ---- K7: backward.cu:346-412
    real mh[4];
    xf4x4(mean, proj, mh);
    const real mw = (real)1 / (mh[3] + (real)0.0000001f);
    const real mul1 = (proj[0] * mean[0] + proj[4] * mean[1] + proj[8] * mean[2] + proj[12]) * mw * mw;
    const real mul2 = (proj[1] * mean[0] + proj[5] * mean[1] + proj[9] * mean[2] + proj[13]) * mw * mw;
    const real g2x = dL_dmean2D[3 * (size_t)i], g2y = dL_dmean2D[3 * (size_t)i + 1];
    dmean[0] += (proj[0] * mw - proj[3] * mul1) * g2x + (proj[1] * mw - proj[3] * mul2) * g2y;
    dmean[1] += (proj[4] * mw - proj[7] * mul1) * g2x + (proj[5] * mw - proj[7] * mul2) * g2y;
    dmean[2] += (proj[8] * mw - proj[11] * mul1) * g2x + (proj[9] * mw - proj[11] * mul2) * g2y;
    const real mul3 = view[2] * mean[0] + view[6] * mean[1] + view[10] * mean[2] + view[14];
    const real gd = dL_ddepth[i];
    dmean[0] += (view[2] - view[3] * mul3) * gd;
    dmean[1] += (view[6] - view[7] * mul3) * gd;
    dmean[2] += (view[10] - view[11] * mul3) * gd;

    if (shs) {
      // backward.cu:20-139
      const real* sh = shs + (size_t)i * M * 3;
      real* dsh = dL_dsh + (size_t)i * M * 3;
      real dorig[3] = {mean[0] - campos[0], mean[1] - campos[1], mean[2] - campos[2]};
      const real len = std::sqrt(dorig[0] * dorig[0] + dorig[1] * dorig[1] + dorig[2] * dorig[2]);
      const real x = dorig[0] / len, y = dorig[1] / len, z = dorig[2] / len;
      real dRGB[3];
      for (int c = 0; c < 3; c++) dRGB[c] = dL_dcolor[3 * (size_t)i + c] * (st.clamped[3 * (size_t)i + c] ? (real)0 : (real)1);
      real ddir[3] = {0, 0, 0};
      const real C0 = C_<real>(kSH_C0), C1 = C_<real>(kSH_C1);
      real w[16] = {0};
      real dx_[16] = {0}, dy_[16] = {0}, dz_[16] = {0};  // d(basis_k)/d(x,y,z)
      w[0] = C0;
      if (D > 0) {
        w[1] = -C1 * y; w[2] = C1 * z; w[3] = -C1 * x;
        dx_[3] = -C1; dy_[1] = -C1; dz_[2] = C1;
        if (D > 1) {
          const real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
          const real c20 = C_<real>(kSH_C2[0]), c21 = C_<real>(kSH_C2[1]), c22 = C_<real>(kSH_C2[2]),
                     c23 = C_<real>(kSH_C2[3]), c24 = C_<real>(kSH_C2[4]);
          w[4] = c20 * xy; w[5] = c21 * yz; w[6] = c22 * ((real)2 * zz - xx - yy); w[7] = c23 * xz; w[8] = c24 * (xx - yy);
          dx_[4] = c20 * y; dx_[6] = c22 * (real)2 * -x; dx_[7] = c23 * z; dx_[8] = c24 * (real)2 * x;
          dy_[4] = c20 * x; dy_[5] = c21 * z; dy_[6] = c22 * (real)2 * -y; dy_[8] = c24 * (real)2 * -y;
          dz_[5] = c21 * y; dz_[6] = c22 * (real)2 * (real)2 * z; dz_[7] = c23 * x;
          if (D > 2) {
            const real c30 = C_<real>(kSH_C3[0]), c31 = C_<real>(kSH_C3[1]), c32 = C_<real>(kSH_C3[2]),
                       c33 = C_<real>(kSH_C3[3]), c34 = C_<real>(kSH_C3[4]), c35 = C_<real>(kSH_C3[5]),
                       c36 = C_<real>(kSH_C3[6]);
            w[9] = c30 * y * ((real)3 * xx - yy); w[10] = c31 * xy * z; w[11] = c32 * y * ((real)4 * zz - xx - yy);
            w[12] = c33 * z * ((real)2 * zz - (real)3 * xx - (real)3 * yy); w[13] = c34 * x * ((real)4 * zz - xx - yy);
            w[14] = c35 * z * (xx - yy); w[15] = c36 * x * (xx - (real)3 * yy);
            dx_[9] = c30 * (real)3 * (real)2 * xy; dx_[10] = c31 * yz; dx_[11] = c32 * (real)-2 * xy;
            dx_[12] = c33 * (real)-3 * (real)2 * xz; dx_[13] = c34 * ((real)-3 * xx + (real)4 * zz - yy);
            dx_[14] = c35 * (real)2 * xz; dx_[15] = c36 * (real)3 * (xx - yy);
            dy_[9] = c30 * (real)3 * (xx - yy); dy_[10] = c31 * xz; dy_[11] = c32 * ((real)-3 * yy + (real)4 * zz - xx);
            dy_[12] = c33 * (real)-3 * (real)2 * yz; dy_[13] = c34 * (real)-2 * xy; dy_[14] = c35 * (real)-2 * yz;
            dy_[15] = c36 * (real)-3 * (real)2 * xy;
            dz_[10] = c31 * xy; dz_[11] = c32 * (real)4 * (real)2 * yz; dz_[12] = c33 * (real)3 * ((real)2 * zz - xx - yy);
            dz_[13] = c34 * (real)4 * (real)2 * xz; dz_[14] = c35 * (xx - yy);
          }
        }
      }
      const int nco = (D + 1) * (D + 1);
      for (int k = 0; k < nco; k++)
        for (int c = 0; c < 3; c++) {
          dsh[3 * k + c] = w[k] * dRGB[c];
          ddir[0] += dx_[k] * sh[3 * k + c] * dRGB[c];
          ddir[1] += dy_[k] * sh[3 * k + c] * dRGB[c];
          ddir[2] += dz_[k] * sh[3 * k + c] * dRGB[c];
        }
      // auxiliary.h:107-117 (dnormvdv)
      const real sum2 = dorig[0] * dorig[0] + dorig[1] * dorig[1] + dorig[2] * dorig[2];
      const real inv32 = (real)1 / std::sqrt(sum2 * sum2 * sum2);
      dmean[0] += ((+sum2 - dorig[0] * dorig[0]) * ddir[0] - dorig[1] * dorig[0] * ddir[1] - dorig[2] * dorig[0] * ddir[2]) * inv32;
      dmean[1] += (-dorig[0] * dorig[1] * ddir[0] + (sum2 - dorig[1] * dorig[1]) * ddir[1] - dorig[2] * dorig[1] * ddir[2]) * inv32;
      dmean[2] += (-dorig[0] * dorig[2] * ddir[0] - dorig[1] * dorig[2] * ddir[1] + (sum2 - dorig[2] * dorig[2]) * ddir[2]) * inv32;
    }
    for (int c = 0; c < 3; c++) dL_dmean3D[3 * (size_t)i + c] = dmean[c];

    if (scales) {
      // backward.cu:278-341 (gradient w.r.t. the un-normalised quaternion, no dnormvdv)
      M3<real> Mm, Rm;
      real tmp[6];
      cov3d_from_scale_rot<real>(scales + 3 * (size_t)i, scale_modifier, rotations + 4 * (size_t)i, tmp, &Mm, &Rm);
      const real s3[3] = {scale_modifier * scales[3 * (size_t)i], scale_modifier * scales[3 * (size_t)i + 1],
                          scale_modifier * scales[3 * (size_t)i + 2]};
      M3<real> dSg;
      dSg.m[0][0] = dcv[0]; dSg.m[0][1] = (real)0.5 * dcv[1]; dSg.m[0][2] = (real)0.5 * dcv[2];
      dSg.m[1][0] = (real)0.5 * dcv[1]; dSg.m[1][1] = dcv[3]; dSg.m[1][2] = (real)0.5 * dcv[4];
      dSg.m[2][0] = (real)0.5 * dcv[2]; dSg.m[2][1] = (real)0.5 * dcv[4]; dSg.m[2][2] = dcv[5];
      M3<real> M2 = Mm;
      for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) M2.m[c][r] = Mm.m[c][r] * (real)2;
      M3<real> dM = mul(M2, dSg);
      M3<real> Rt = tr(Rm), dMt = tr(dM);
      real* dsc = dL_dscale + 3 * (size_t)i;
      for (int k = 0; k < 3; k++) dsc[k] = Rt.m[k][0] * dMt.m[k][0] + Rt.m[k][1] * dMt.m[k][1] + Rt.m[k][2] * dMt.m[k][2];
      for (int k = 0; k < 3; k++)
        for (int r = 0; r < 3; r++) dMt.m[k][r] *= s3[k];
      const real* q = rotations + 4 * (size_t)i;
      const real r = q[0], x = q[1], y = q[2], z = q[3];
      real* dq = dL_drot + 4 * (size_t)i;
      dq[0] = (real)2 * z * (dMt.m[0][1] - dMt.m[1][0]) + (real)2 * y * (dMt.m[2][0] - dMt.m[0][2]) + (real)2 * x * (dMt.m[1][2] - dMt.m[2][1]);
      dq[1] = (real)2 * y * (dMt.m[1][0] + dMt.m[0][1]) + (real)2 * z * (dMt.m[2][0] + dMt.m[0][2]) + (real)2 * r * (dMt.m[1][2] - dMt.m[2][1]) - (real)4 * x * (dMt.m[2][2] + dMt.m[1][1]);
      dq[2] = (real)2 * x * (dMt.m[1][0] + dMt.m[0][1]) + (real)2 * r * (dMt.m[2][0] - dMt.m[0][2]) + (real)2 * z * (dMt.m[1][2] + dMt.m[2][1]) - (real)4 * y * (dMt.m[2][2] + dMt.m[0][0]);
      dq[3] = (real)2 * r * (dMt.m[0][1] - dMt.m[1][0]) + (real)2 * x * (dMt.m[2][0] + dMt.m[0][2]) + (real)2 * y * (dMt.m[1][2] + dMt.m[2][1]) - (real)4 * z * (dMt.m[1][1] + dMt.m[0][0]);
    }
  }
}

// gaustudio/datasets/__init__.py:106-112, 307-380 (depth2point + depth2normal, k=3, camera coords)
template <typename real>
void depth2normal(const real* depth, int W, int H, real fx, real fy, real cx, real cy, real dmin, real dmax,
                  const real* rot /*3x3 row-major or null*/, real* out /*[H,W,3]*/) {
  // K^-1 = [[1/fx,0,-cx/fx],[0,1/fy,-cy/fy],[0,0,1]]
  auto point = [&](int u, int v, real* X) {
    const real z = depth[(size_t)v * W + u];
    const real uz = ((real)u / (real)(W - 1)) * (real)(W - 1) * z, vz = ((real)v / (real)(H - 1)) * (real)(H - 1) * z;
    X[0] = uz * ((real)1 / fx) + z * (-cx / fx);
    X[1] = vz * ((real)1 / fy) + z * (-cy / fy);
    X[2] = z;
  };
#pragma omp parallel for schedule(static)
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      real* o = out + ((size_t)v * W + u) * 3;
      bool valid = u > 0 && v > 0 && u < W - 1 && v < H - 1;
      real Pc[3], Pt[3], Pb[3], Pl[3], Pr[3];
      if (valid) {
        point(u, v, Pc); point(u, v - 1, Pt); point(u, v + 1, Pb); point(u - 1, v, Pl); point(u + 1, v, Pr);
        auto ok = [&](const real* X) { return X[2] > dmin && X[2] < dmax; };
        valid = ok(Pc) && ok(Pt) && ok(Pb) && ok(Pl) && ok(Pr);
      }
      if (!valid) { o[0] = o[1] = o[2] = (real)-1; continue; }
      const real a[3] = {Pt[0] - Pb[0], Pt[1] - Pb[1], Pt[2] - Pb[2]};
      const real b[3] = {Pl[0] - Pr[0], Pl[1] - Pr[1], Pl[2] - Pr[2]};
      real n[3] = {-(a[1] * b[2] - a[2] * b[1]), -(a[2] * b[0] - a[0] * b[2]), -(a[0] * b[1] - a[1] * b[0])};
      const real len = std::max(std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), (real)1e-12);
      n[0] /= len; n[1] /= len; n[2] /= len;
      if (rot) {
        // normal_row @ rot  (rot = inverse(extrinsics[:3,:3]).t())
        const real m0 = n[0] * rot[0] + n[1] * rot[3] + n[2] * rot[6];
        const real m1 = n[0] * rot[1] + n[1] * rot[4] + n[2] * rot[7];
        const real m2 = n[0] * rot[2] + n[1] * rot[5] + n[2] * rot[8];
        n[0] = m0; n[1] = m1; n[2] = m2;
      }
      o[0] = n[0]; o[1] = n[1]; o[2] = n[2];
    }
}

template <typename real> void mark_visible(int P, const real* means3D, const real* view, uint8_t* present) {
  // rasterizer_impl.cu:54-66 + auxiliary.h:139-164
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    real pv[3];
    xf4x3(means3D + 3 * (size_t)i, view, pv);
    present[i] = pv[2] > (real)0.2f;
  }
}

}  // namespace

#define GSO_API(SUF, real)                                                                                             \
  extern "C" void* gso_create_##SUF() { return new State<real>(); }                                                   \
  extern "C" void gso_destroy_##SUF(void* h) { delete (State<real>*)h; }                                               \
  extern "C" int64_t gso_forward_##SUF(void* h, int P, int D, int M, int W, int H, const real* means3D,              \
                                       const real* shs, const real* colors_precomp, const real* opacities,            \
                                       const real* scales, real scale_modifier, const real* rotations,                \
                                       const real* cov3D_precomp, const real* view, const real* proj,                 \
                                       const real* campos, real tanx, real tany, real* out_color, real* out_depth,    \
                                       real* out_median, real* out_opacity, int* radii) {                             \
    return forward<real>(*(State<real>*)h, P, D, M, W, H, means3D, shs, colors_precomp, opacities, scales,            \
                         scale_modifier, rotations, cov3D_precomp, view, proj, campos, tanx, tany, out_color,         \
                         out_depth, out_median, out_opacity, radii);                                                  \
  }                                                                                                                    \
  extern "C" void gso_backward_##SUF(void* h, const real* bg, const real* means3D, const real* shs,                   \
                                     const real* colors_precomp, const real* scales, real scale_modifier,             \
                                     const real* rotations, const real* cov3D_precomp, const real* view,              \
                                     const real* proj, const real* campos, real tanx, real tany, const real* dL_dpix, \
                                     const real* dL_ddepth, const real* dL_dmedian, const real* dL_dopac,             \
                                     real* dL_dmean2D, real* dL_dconic, real* dL_dopacity, real* dL_dcolor,           \
                                     real* dL_ddepths, real* dL_dmean3D, real* dL_dcov3D, real* dL_dsh,               \
                                     real* dL_dscale, real* dL_drot) {                                                \
    backward<real>(*(State<real>*)h, bg, means3D, shs, colors_precomp, scales, scale_modifier, rotations,             \
                   cov3D_precomp, view, proj, campos, tanx, tany, dL_dpix, dL_ddepth, dL_dmedian, dL_dopac,           \
                   dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_ddepths, dL_dmean3D, dL_dcov3D, dL_dsh,          \
                   dL_dscale, dL_drot);                                                                               \
  }                                                                                                                    \
  extern "C" void gso_get_binning_##SUF(void* h, uint32_t* point_list, uint32_t* ranges, uint32_t* n_contrib,         \
                                        real* final_T) {                                                              \
    State<real>& s = *(State<real>*)h;                                                                                \
    if (point_list) std::memcpy(point_list, s.point_list.data(), s.point_list.size() * 4);                            \
    if (ranges) std::memcpy(ranges, s.ranges.data(), s.ranges.size() * 4);                                            \
    if (n_contrib) std::memcpy(n_contrib, s.n_contrib.data(), s.n_contrib.size() * 4);                                \
    if (final_T) std::memcpy(final_T, s.final_T.data(), s.final_T.size() * sizeof(real));                             \
  }                                                                                                                    \
  extern "C" void gso_get_geometry_##SUF(void* h, real* depths, real* means2D, real* conic_opacity, real* rgb,        \
                                         real* cov3D, uint32_t* tiles_touched, uint8_t* clamped) {                    \
    State<real>& s = *(State<real>*)h;                                                                                \
    if (depths) std::memcpy(depths, s.depths.data(), s.depths.size() * sizeof(real));                                 \
    if (means2D) std::memcpy(means2D, s.means2D.data(), s.means2D.size() * sizeof(real));                             \
    if (conic_opacity) std::memcpy(conic_opacity, s.conic_opacity.data(), s.conic_opacity.size() * sizeof(real));     \
    if (rgb) std::memcpy(rgb, s.rgb.data(), s.rgb.size() * sizeof(real));                                             \
    if (cov3D) std::memcpy(cov3D, s.cov3D.data(), s.cov3D.size() * sizeof(real));                                     \
    if (tiles_touched) std::memcpy(tiles_touched, s.tiles_touched.data(), s.tiles_touched.size() * 4);                \
    if (clamped) std::memcpy(clamped, s.clamped.data(), s.clamped.size());                                            \
  }                                                                                                                    \
  extern "C" void gso_mark_visible_##SUF(int P, const real* means3D, const real* view, uint8_t* present) {            \
    mark_visible<real>(P, means3D, view, present);                                                                    \
  }                                                                                                                    \
  extern "C" void gso_depth2normal_##SUF(const real* depth, int W, int H, real fx, real fy, real cx, real cy,         \
                                         real dmin, real dmax, const real* rot, real* out) {                          \
    depth2normal<real>(depth, W, H, fx, fy, cx, cy, dmin, dmax, rot, out);                                            \
  }

GSO_API(f32, float)
GSO_API(f64, double)

extern "C" int gso_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
