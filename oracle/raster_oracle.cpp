// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference forward rasterizer.
// Nothing in the product path (gs2mesh_b200/) may import, link or call this file;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs do.
//
// Restates, in plain scalar C++ (fp32 like the device code, fp64 only where the
// reference promotes), the forward path of
//   third_party/gaussian-splatting/submodules/diff-gaussian-rasterization (= DGR):
//     DGR/cuda_rasterizer/auxiliary.h:22-39   SH constants
//     DGR/cuda_rasterizer/auxiliary.h:41-56   ndc2Pix (double), getRect
//     DGR/cuda_rasterizer/auxiliary.h:58-77   row-vector matrix indexing
//     DGR/cuda_rasterizer/auxiliary.h:139-164 near cull (z_view <= 0.2)
//     DGR/cuda_rasterizer/forward.cu:20-71    SH -> RGB (+0.5, clamp >= 0)
//     DGR/cuda_rasterizer/forward.cu:74-113   EWA 2D covariance (+0.3 low-pass)
//     DGR/cuda_rasterizer/forward.cu:118-152  3D covariance from scale / quaternion
//     DGR/cuda_rasterizer/forward.cu:182-255  preprocess order and early-outs
//     DGR/cuda_rasterizer/forward.cu:275-373  front-to-back blend loop
//     DGR/cuda_rasterizer/rasterizer_impl.cu:35-50   getHigherMsb
//     DGR/cuda_rasterizer/rasterizer_impl.cu:88-107  key = tile<<32 | depth bits
//     DGR/cuda_rasterizer/rasterizer_impl.cu:116-138 tile ranges
//     DGR/cuda_rasterizer/rasterizer_impl.cu:222-234 focal from tan(fov), tile grid
// plus ONE extension that has no reference counterpart: an expected-depth
// channel D = sum_i z_i * alpha_i * T_i accumulated next to the colour.
//
// Pinned (tests/test_oracle_vs_reference.py, -m gpu) against the reference itself built
// for sm_100a (oracle/_ref/libref_dgr.so) and against golden outputs of that build
// committed under tests/golden/.
//
// The file is compiled twice: -ffp-contract=off (liboracle.so) and with FMA contraction
// (liboracle_fma.so) to bracket what nvcc does to the reference's device code.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr int TILE = 16;  // DGR/cuda_rasterizer/config.h:16-17

constexpr float kSH0 = 0.28209479177387814f;
constexpr float kSH1 = 0.4886025119029199f;
constexpr float kSH2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                           -1.0925484305920792f, 0.5462742152960396f};
constexpr float kSH3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                           0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                           -0.5900435899266435f};

struct V3 {
  float x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }

// Column-major 3x3 with glm's multiplication order: c[col][row].
struct M3 {
  float c[3][3];
};
inline M3 mul(const M3& a, const M3& b) {  // glm operator*(mat3,mat3): sum over k ascending
  M3 r;
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) r.c[j][i] = a.c[0][i] * b.c[j][0] + a.c[1][i] * b.c[j][1] + a.c[2][i] * b.c[j][2];
  return r;
}
inline M3 transpose(const M3& a) {
  M3 r;
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) r.c[j][i] = a.c[i][j];
  return r;
}

inline V3 xform43(V3 p, const float* m) {  // auxiliary.h:58-66
  return {m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
          m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]};
}
inline void xform44(V3 p, const float* m, float out[4]) {  // auxiliary.h:68-77
  out[0] = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12];
  out[1] = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13];
  out[2] = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14];
  out[3] = m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15];
}

inline float ndc_to_pix(float v, int s) {  // auxiliary.h:41-44, evaluated in double
  return (float)(((v + 1.0) * s - 1.0) * 0.5);
}

struct Rect {
  uint32_t x0, y0, x1, y1;
};
inline Rect tile_rect(float px, float py, int r, uint32_t gx, uint32_t gy) {  // auxiliary.h:46-56
  auto lo = [](float c, int rad, uint32_t g) {
    int v = (int)((c - rad) / TILE);
    return std::min<uint32_t>(g, (uint32_t)std::max(0, v));
  };
  auto hi = [](float c, int rad, uint32_t g) {
    int v = (int)((c + rad + TILE - 1) / TILE);
    return std::min<uint32_t>(g, (uint32_t)std::max(0, v));
  };
  return {lo(px, r, gx), lo(py, r, gy), hi(px, r, gx), hi(py, r, gy)};
}

void cov3d_from_scale_rot(const float* s, float mod, const float* q, float out[6]) {  // forward.cu:118-152
  M3 S{};
  S.c[0][0] = mod * s[0];
  S.c[1][1] = mod * s[1];
  S.c[2][2] = mod * s[2];
  const float r = q[0], x = q[1], y = q[2], z = q[3];  // w-first, NOT renormalised (forward.cu:127)
  M3 R;
  R.c[0][0] = 1.f - 2.f * (y * y + z * z);
  R.c[0][1] = 2.f * (x * y - r * z);
  R.c[0][2] = 2.f * (x * z + r * y);
  R.c[1][0] = 2.f * (x * y + r * z);
  R.c[1][1] = 1.f - 2.f * (x * x + z * z);
  R.c[1][2] = 2.f * (y * z - r * x);
  R.c[2][0] = 2.f * (x * z - r * y);
  R.c[2][1] = 2.f * (y * z + r * x);
  R.c[2][2] = 1.f - 2.f * (x * x + y * y);
  M3 M = mul(S, R);
  M3 Sg = mul(transpose(M), M);
  out[0] = Sg.c[0][0];
  out[1] = Sg.c[0][1];
  out[2] = Sg.c[0][2];
  out[3] = Sg.c[1][1];
  out[4] = Sg.c[1][2];
  out[5] = Sg.c[2][2];
}

void cov2d(V3 mean, float fx, float fy, float tanx, float tany, const float* c3, const float* vm, float out[3]) {
  // forward.cu:74-113
  V3 t = xform43(mean, vm);
  const float limx = 1.3f * tanx, limy = 1.3f * tany;
  const float txtz = t.x / t.z, tytz = t.y / t.z;
  t.x = std::min(limx, std::max(-limx, txtz)) * t.z;
  t.y = std::min(limy, std::max(-limy, tytz)) * t.z;
  M3 J{};
  J.c[0][0] = fx / t.z;
  J.c[0][2] = -(fx * t.x) / (t.z * t.z);
  J.c[1][1] = fy / t.z;
  J.c[1][2] = -(fy * t.y) / (t.z * t.z);
  M3 W;
  W.c[0][0] = vm[0];
  W.c[0][1] = vm[4];
  W.c[0][2] = vm[8];
  W.c[1][0] = vm[1];
  W.c[1][1] = vm[5];
  W.c[1][2] = vm[9];
  W.c[2][0] = vm[2];
  W.c[2][1] = vm[6];
  W.c[2][2] = vm[10];
  M3 T = mul(W, J);
  M3 V;
  V.c[0][0] = c3[0];
  V.c[0][1] = c3[1];
  V.c[0][2] = c3[2];
  V.c[1][0] = c3[1];
  V.c[1][1] = c3[3];
  V.c[1][2] = c3[4];
  V.c[2][0] = c3[2];
  V.c[2][1] = c3[4];
  V.c[2][2] = c3[5];
  M3 cov = mul(mul(transpose(T), transpose(V)), T);
  out[0] = cov.c[0][0] + 0.3f;
  out[1] = cov.c[0][1];
  out[2] = cov.c[1][1] + 0.3f;
}

V3 sh_to_rgb(int deg, int ncoef, V3 pos, V3 cam, const float* sh_all, int idx) {  // forward.cu:20-71
  V3 d = pos - cam;
  float len = std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
  d = {d.x / len, d.y / len, d.z / len};
  const V3* sh = reinterpret_cast<const V3*>(sh_all) + (size_t)idx * ncoef;
  V3 res = kSH0 * sh[0];
  if (deg > 0) {
    float x = d.x, y = d.y, z = d.z;
    res = res - (kSH1 * y) * sh[1] + (kSH1 * z) * sh[2] - (kSH1 * x) * sh[3];
    if (deg > 1) {
      float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      res = res + (kSH2[0] * xy) * sh[4] + (kSH2[1] * yz) * sh[5] + (kSH2[2] * (2.0f * zz - xx - yy)) * sh[6] +
            (kSH2[3] * xz) * sh[7] + (kSH2[4] * (xx - yy)) * sh[8];
      if (deg > 2) {
        res = res + (kSH3[0] * y * (3.0f * xx - yy)) * sh[9] + (kSH3[1] * xy * z) * sh[10] +
              (kSH3[2] * y * (4.0f * zz - xx - yy)) * sh[11] +
              (kSH3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy)) * sh[12] +
              (kSH3[4] * x * (4.0f * zz - xx - yy)) * sh[13] + (kSH3[5] * z * (xx - yy)) * sh[14] +
              (kSH3[6] * x * (xx - 3.0f * yy)) * sh[15];
      }
    }
  }
  res = {res.x + 0.5f, res.y + 0.5f, res.z + 0.5f};
  return {std::max(res.x, 0.0f), std::max(res.y, 0.0f), std::max(res.z, 0.0f)};
}

inline uint32_t f2u(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}

}  // namespace

extern "C" {

// rasterizer_impl.cu:35-50
uint32_t orc_higher_msb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4, step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb)
      msb += step;
    else
      msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

// Per-Gaussian stage (forward.cu:155-256).  All outputs are length-P arrays (xy: 2P,
// conic_opacity: 4P, rgb: 3P, cov3d: 6P) and may be NULL.  Entries of culled Gaussians
// are left untouched except radii / tiles_touched which are zeroed, as on the device.
void orc_preprocess(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
                    const float* rotations, const float* opacities, const float* shs, const float* cov3D_precomp,
                    const float* colors_precomp, const float* viewmatrix, const float* projmatrix,
                    const float* cam_pos, int W, int H, float tan_fovx, float tan_fovy, int* radii, float* xy,
                    float* depths, float* cov3d, float* rgb, float* conic_opacity, uint32_t* tiles_touched) {
  const float focal_y = H / (2.0f * tan_fovy);  // rasterizer_impl.cu:222-223
  const float focal_x = W / (2.0f * tan_fovx);
  const uint32_t gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const V3 cam = {cam_pos[0], cam_pos[1], cam_pos[2]};
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; ++i) {
    if (radii) radii[i] = 0;
    if (tiles_touched) tiles_touched[i] = 0;
    V3 p = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]};
    V3 pv = xform43(p, viewmatrix);
    if (pv.z <= 0.2f) continue;  // auxiliary.h:154
    float hom[4];
    xform44(p, projmatrix, hom);
    float pw = 1.0f / (hom[3] + 0.0000001f);
    float projx = hom[0] * pw, projy = hom[1] * pw;
    float c3[6];
    if (cov3D_precomp)
      std::memcpy(c3, cov3D_precomp + 6 * (size_t)i, sizeof(c3));
    else
      cov3d_from_scale_rot(scales + 3 * (size_t)i, scale_modifier, rotations + 4 * (size_t)i, c3);
    if (cov3d) std::memcpy(cov3d + 6 * (size_t)i, c3, sizeof(c3));
    float cv[3];
    cov2d(p, focal_x, focal_y, tan_fovx, tan_fovy, c3, viewmatrix, cv);
    float det = cv[0] * cv[2] - cv[1] * cv[1];
    if (det == 0.0f) continue;
    float det_inv = 1.f / det;
    float conic[3] = {cv[2] * det_inv, -cv[1] * det_inv, cv[0] * det_inv};
    float mid = 0.5f * (cv[0] + cv[2]);
    float lambda1 = mid + std::sqrt(std::max(0.1f, mid * mid - det));
    float lambda2 = mid - std::sqrt(std::max(0.1f, mid * mid - det));
    float my_radius = std::ceil(3.f * std::sqrt(std::max(lambda1, lambda2)));
    float px = ndc_to_pix(projx, W), py = ndc_to_pix(projy, H);
    Rect rc = tile_rect(px, py, (int)my_radius, gx, gy);
    if ((rc.x1 - rc.x0) * (rc.y1 - rc.y0) == 0) continue;
    if (rgb) {
      if (colors_precomp == nullptr) {
        V3 c = sh_to_rgb(D, M, p, cam, shs, i);
        rgb[3 * i] = c.x;
        rgb[3 * i + 1] = c.y;
        rgb[3 * i + 2] = c.z;
      } else {
        rgb[3 * i] = colors_precomp[3 * i];
        rgb[3 * i + 1] = colors_precomp[3 * i + 1];
        rgb[3 * i + 2] = colors_precomp[3 * i + 2];
      }
    }
    if (depths) depths[i] = pv.z;
    if (radii) radii[i] = (int)my_radius;
    if (xy) {
      xy[2 * i] = px;
      xy[2 * i + 1] = py;
    }
    if (conic_opacity) {
      conic_opacity[4 * i] = conic[0];
      conic_opacity[4 * i + 1] = conic[1];
      conic_opacity[4 * i + 2] = conic[2];
      conic_opacity[4 * i + 3] = opacities[i];
    }
    if (tiles_touched) tiles_touched[i] = (rc.y1 - rc.y0) * (rc.x1 - rc.x0);
  }
}

// Full forward pass.  Outputs (any may be NULL): out_color [3,H,W], out_depth [H,W]
// (extension), final_T [H,W], n_contrib [H,W], radii [P], point_list (sorted Gaussian
// ids, capacity list_cap), ranges [tiles*2].  Returns num_rendered.
int64_t orc_forward(int P, int D, int M, const float* background, int W, int H, const float* means3D,
                    const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
                    float scale_modifier, const float* rotations, const float* cov3D_precomp,
                    const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                    float tan_fovy, float* out_color, float* out_depth, float* final_T, uint32_t* n_contrib,
                    int* radii, uint32_t* point_list, int64_t list_cap, uint32_t* ranges) {
  if (P == 0) {  // rasterize_points.cu:68-77: outputs stay zero-filled when there are no points
    const size_t n = (size_t)W * H;
    if (out_color) std::memset(out_color, 0, 3 * n * sizeof(float));
    if (out_depth) std::memset(out_depth, 0, n * sizeof(float));
    if (final_T) std::memset(final_T, 0, n * sizeof(float));
    if (n_contrib) std::memset(n_contrib, 0, n * sizeof(uint32_t));
    return 0;
  }
  std::vector<int> rad(P);
  std::vector<float> xy(2 * (size_t)P), depth(P), rgb(3 * (size_t)P), co(4 * (size_t)P);
  std::vector<uint32_t> touched(P);
  orc_preprocess(P, D, M, means3D, scales, scale_modifier, rotations, opacities, shs, cov3D_precomp, colors_precomp,
                 viewmatrix, projmatrix, cam_pos, W, H, tan_fovx, tan_fovy, rad.data(), xy.data(), depth.data(),
                 nullptr, rgb.data(), co.data(), touched.data());
  if (radii) std::memcpy(radii, rad.data(), sizeof(int) * (size_t)P);
  const uint32_t gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const size_t ntiles = (size_t)gx * gy;

  // duplicateWithKeys + stable radix sort by (tile, depth bits) == per-tile lists filled in
  // ascending Gaussian order, each stably sorted by depth bits (rasterizer_impl.cu:88-107,303).
  struct Inst {
    uint32_t key, id;
  };
  std::vector<std::vector<Inst>> bins(ntiles);
  int64_t num_rendered = 0;
  for (int i = 0; i < P; ++i) {
    if (rad[i] <= 0) continue;
    Rect rc = tile_rect(xy[2 * i], xy[2 * i + 1], rad[i], gx, gy);
    uint32_t dk = f2u(depth[i]);
    for (uint32_t y = rc.y0; y < rc.y1; ++y)
      for (uint32_t x = rc.x0; x < rc.x1; ++x) {
        bins[(size_t)y * gx + x].push_back({dk, (uint32_t)i});
        ++num_rendered;
      }
  }
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t t = 0; t < (int64_t)ntiles; ++t)
    std::stable_sort(bins[t].begin(), bins[t].end(), [](const Inst& a, const Inst& b) { return a.key < b.key; });

  if (ranges || point_list) {
    int64_t off = 0;
    for (size_t t = 0; t < ntiles; ++t) {
      // untouched tiles keep (0,0) like the memset in rasterizer_impl.cu:310
      if (ranges) {
        ranges[2 * t] = bins[t].empty() ? 0u : (uint32_t)off;
        ranges[2 * t + 1] = bins[t].empty() ? 0u : (uint32_t)(off + (int64_t)bins[t].size());
      }
      for (const Inst& in : bins[t]) {
        if (point_list && off < list_cap) point_list[off] = in.id;
        ++off;
      }
    }
  }

  const float bg[3] = {background[0], background[1], background[2]};
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t t = 0; t < (int64_t)ntiles; ++t) {
    const uint32_t ty = (uint32_t)(t / gx), tx = (uint32_t)(t % gx);
    const std::vector<Inst>& list = bins[t];
    for (uint32_t ly = 0; ly < TILE; ++ly)
      for (uint32_t lx = 0; lx < TILE; ++lx) {
        const uint32_t pxl = tx * TILE + lx, pyl = ty * TILE + ly;
        if (pxl >= (uint32_t)W || pyl >= (uint32_t)H) continue;
        const float pfx = (float)pxl, pfy = (float)pyl;
        float T = 1.0f, C[3] = {0, 0, 0}, Dacc = 0.0f;
        uint32_t contributor = 0, last = 0;
        for (const Inst& in : list) {  // forward.cu:323-362
          ++contributor;
          const uint32_t g = in.id;
          float dx = xy[2 * g] - pfx, dy = xy[2 * g + 1] - pfy;
          const float* c = &co[4 * (size_t)g];
          float power = -0.5f * (c[0] * dx * dx + c[2] * dy * dy) - c[1] * dx * dy;
          if (power > 0.0f) continue;
          float alpha = std::min(0.99f, c[3] * std::exp(power));
          if (alpha < 1.0f / 255.0f) continue;
          float test_T = T * (1 - alpha);
          if (test_T < 0.0001f) break;  // `done = true`
          for (int ch = 0; ch < 3; ++ch) C[ch] += rgb[3 * (size_t)g + ch] * alpha * T;
          Dacc += depth[g] * alpha * T;  // extension: expected depth
          T = test_T;
          last = contributor;
        }
        const size_t pid = (size_t)pyl * W + pxl;
        if (final_T) final_T[pid] = T;
        if (n_contrib) n_contrib[pid] = last;
        if (out_color)
          for (int ch = 0; ch < 3; ++ch) out_color[(size_t)ch * H * W + pid] = C[ch] + T * bg[ch];
        if (out_depth) out_depth[pid] = Dacc;
      }
  }
  return num_rendered;
}

}  // extern "C"
