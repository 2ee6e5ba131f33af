// Forward Gaussian-splat rasterizer for sm_100a.
//
// Behavioural contract = the reference's forward pass
//   DGR/cuda_rasterizer/rasterizer_impl.cu:198-336 (Rasterizer::forward)
//   DGR/cuda_rasterizer/forward.cu:155-256 (preprocess), :261-374 (render)
// (DGR = third_party/gaussian-splatting/submodules/diff-gaussian-rasterization), but the
// pipeline is organised differently:
//
//   preprocess   one WARP per 32 consecutive Gaussians.  The warp's parameter block is
//                contiguous in every reference tensor ([P,3] xyz / scales, [P,4] quaternions,
//                [P] opacities, [P,M,3] SH = 6 KB per warp) and is staged into shared memory
//                with 1-D bulk TMA copies (cp.async.bulk + mbarrier).  The 6 KB SH block is
//                only fetched when at least one Gaussian of the warp survives culling, and is
//                read back with 16-byte loads.  Results are written as three coalesced record planes:
//                  recA = (px, py, z_view, opacity)  recB = (conic a, b, c, red)  recC = (green, blue)
//   binning      exact (Gaussian,tile) test: a pair is kept only if the tile's pixel lattice can
//                reach alpha >= 1/255 -- the image is bit-identical to rectangle binning
//                (rasterizer_impl.cu:88-107) with far fewer instances to sort and blend.
//   sort         the reference's order inside a tile is (depth bits, Gaussian index) ascending
//                (stable radix sort of tile << 32 | depth keys emitted in index order).  Here: stable
//                LSD radix sort of the P (depth bits, index) pairs, instances emitted in that order,
//                then a stable sort of the R instances on the tile id alone (gsb_radix.cuh).
//   render       per 16x16 tile, four warps that never wait for each other: each streams the tile's
//                list, keeps the records its 8x8 pixel block can see (exact footprint test) in a
//                private compacted shared-memory list and blends them branch-free, two pixels per lane;
//                also accumulates the expected-depth channel (sum z*alpha*T) the TSDF stage consumes.
//
// All kernels run on the caller's stream.
#include <algorithm>
#include <cstdlib>

#include <cub/cub.cuh>

#include "gsb_common.h"
#include "gsb_radix.cuh"

namespace gsb {

thread_local char g_error[512] = {0};
std::atomic<uint64_t> g_launches{0};
Profiler g_prof;
static const char* kStageNames[kStCount] = {"preprocess", "depth_sort", "emit", "tile_sort", "tile_ranges",
                                            "render",     "to_u8", "prepare_depth", "mark_bricks", "integrate"};
static thread_local int64_t g_required_instances = 0;

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kPreThreads = kWarpsPerBlock * 32;
constexpr int kMaxShFloats = 48;  // 16 coefficients x RGB

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D bulk TMA (global -> shared)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
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
// Bounded wait: a bulk copy that never lands (bad pointer / size) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 22)) __trap();
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// Per-Gaussian maths.  Expressions are kept in the reference's algebraic form so nvcc contracts
// them the same way it contracts the reference's device code.
// ---------------------------------------------------------------------------------------------
struct Mat3 {  // column-major, m[col][row] (glm convention)
  float m[3][3];
};
__device__ __forceinline__ Mat3 mat_mul(const Mat3& a, const Mat3& b) {
  Mat3 r;
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) r.m[j][i] = a.m[0][i] * b.m[j][0] + a.m[1][i] * b.m[j][1] + a.m[2][i] * b.m[j][2];
  return r;
}
__device__ __forceinline__ Mat3 mat_t(const Mat3& a) {
  Mat3 r;
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) r.m[j][i] = a.m[i][j];
  return r;
}

struct TileRect {
  uint32_t x0, y0, x1, y1;
};
// auxiliary.h:46-56
__device__ __forceinline__ TileRect tile_rect(float px, float py, int radius, uint32_t gx, uint32_t gy) {
  TileRect r;
  r.x0 = min(gx, (uint32_t)max(0, (int)((px - radius) / kTile)));
  r.y0 = min(gy, (uint32_t)max(0, (int)((py - radius) / kTile)));
  r.x1 = min(gx, (uint32_t)max(0, (int)((px + radius + kTile - 1) / kTile)));
  r.y1 = min(gy, (uint32_t)max(0, (int)((py + radius + kTile - 1) / kTile)));
  return r;
}

// Exact footprint test.  A pixel of the box [x0,x1] x [y0,y1] (pixel centres) can only be blended if
//   power = -0.5*(a dx^2 + c dy^2) - b dx dy <= 0  and  opacity*exp(power) >= 1/255
// (forward.cu:336-346), i.e. q(d) = a dx^2 + 2b dx dy + c dy^2 <= 2 ln(255*opacity).  q is convex,
// so its minimum over the tile's pixel box is 0 if the centre lies inside and otherwise sits on
// one of the four box edges.  A safety margin covers fp32 rounding of both evaluations; pairs
// inside the margin are kept, so no contributing pair is ever dropped.
// __noinline__: the counting and the emitting kernel must execute the same instructions.
struct Footprint {  // per-Gaussian constants of the test
  float a, b, c;      // conic
  float nb_a, nb_c;   // -b/a, -b/c: minimiser slopes along horizontal / vertical box edges
  float two_tau;      // 2 ln(255 * opacity)
  bool degenerate;    // non positive-definite conic: never culled
};
__device__ __forceinline__ Footprint make_footprint(float a, float b, float c, float two_tau) {
  Footprint f;
  f.a = a;
  f.b = b;
  f.c = c;
  f.degenerate = !(a > 0.f && c > 0.f && a * c - b * b > 0.f);
  // approximate division is enough: an error of a few ulp in the clamped minimiser changes q in second order only
  f.nb_a = f.degenerate ? 0.f : __fdividef(-b, a);
  f.nb_c = f.degenerate ? 0.f : __fdividef(-b, c);
  f.two_tau = two_tau;
  return f;
}
__device__ __forceinline__ bool rect_can_contribute(float gxp, float gyp, const Footprint& f, float x0, float y0, float x1,
                                                    float y1) {
  if (f.degenerate) return true;
  const float dx_lo = gxp - x1, dx_hi = gxp - x0;
  const float dy_lo = gyp - y1, dy_hi = gyp - y0;
  if (dx_lo <= 0.f && dx_hi >= 0.f && dy_lo <= 0.f && dy_hi >= 0.f) return true;
  const float DX = fmaxf(fabsf(dx_lo), fabsf(dx_hi)), DY = fmaxf(fabsf(dy_lo), fabsf(dy_hi));
  const float margin = 1e-3f + 8e-6f * (f.a * DX * DX + f.c * DY * DY + 2.f * fabsf(f.b) * DX * DY);
  // The centre lies outside the box, so the minimum of the (convex) quadratic over the box sits on an edge that faces the
  // centre: the vertical edge nearer to it and the horizontal edge nearer to it.  (d = centre - pixel: the nearer edge is
  // the one with the smaller |d|.)  An edge that does not face the centre only yields a value >= the minimum.
  const float dx = fabsf(dx_lo) < fabsf(dx_hi) ? dx_lo : dx_hi;
  const float dy = fminf(dy_hi, fmaxf(dy_lo, f.nb_c * dx));
  const float qv = f.a * dx * dx + 2.f * f.b * dx * dy + f.c * dy * dy;
  const float ey = fabsf(dy_lo) < fabsf(dy_hi) ? dy_lo : dy_hi;
  const float ex = fminf(dx_hi, fmaxf(dx_lo, f.nb_a * ey));
  const float qh = f.a * ex * ex + 2.f * f.b * ex * ey + f.c * ey * ey;
  return fminf(qv, qh) <= f.two_tau + margin;
}
// __noinline__: the counting and the emitting kernel must execute the same instructions.
__device__ __noinline__ bool tile_can_contribute(float gxp, float gyp, float a, float b, float c, float nb_a, float nb_c,
                                                 float two_tau, uint32_t tx, uint32_t ty) {
  Footprint f;
  f.a = a;
  f.b = b;
  f.c = c;
  f.nb_a = nb_a;
  f.nb_c = nb_c;
  f.two_tau = two_tau;
  f.degenerate = false;
  const float x0 = (float)(tx * kTile), y0 = (float)(ty * kTile);
  return rect_can_contribute(gxp, gyp, f, x0, y0, x0 + (kTile - 1), y0 + (kTile - 1));
}

// Candidate tiles under the exact test: the reference rectangle intersected with the tiles that overlap the axis-aligned
// bounding box of the ellipse q(d) <= tau (half extents sqrt(tau c / det), sqrt(tau a / det)).  tau carries the same
// rounding allowance as rect_can_contribute's margin (its S term is <= 4 tau k over the box, k = ac/det); very elongated
// conics (k > 1000) keep the reference rectangle.  Explicitly rounded operations only: the binning, the emitting and the
// validation kernel must all arrive at the same rectangle.
constexpr uint32_t kRectLarge = 0xffffffffu;  // packed-rectangle sentinel: no mask stored, re-test when emitting
__device__ __forceinline__ TileRect shrink_rect(TileRect r, float px, float py, const Footprint& f) {
  if (f.degenerate) return r;
  const float ac = __fmul_rn(f.a, f.c);
  const float det = __fsub_rn(ac, __fmul_rn(f.b, f.b));
  const float k = __fdiv_rn(ac, det);
  if (!(k <= 1000.f)) return r;
  const float tau = __fmul_rn(__fadd_rn(f.two_tau, 2e-3f), __fadd_rn(1.f, __fmul_rn(5e-5f, k)));
  if (!(tau > 0.f)) {  // opacity < 1/255: nothing can be blended
    r.x1 = r.x0;
    r.y1 = r.y0;
    return r;
  }
  const float hx = __fadd_rn(__fmul_rn(__fsqrt_rn(__fdiv_rn(__fmul_rn(tau, f.c), det)), 1.0001f), 0.01f);
  const float hy = __fadd_rn(__fmul_rn(__fsqrt_rn(__fdiv_rn(__fmul_rn(tau, f.a), det)), 1.0001f), 0.01f);
  const float inv = 1.0f / kTile;  // power of two: exact
  const float lox = floorf(__fmul_rn(__fsub_rn(px, hx), inv)), hix = floorf(__fmul_rn(__fadd_rn(px, hx), inv));
  const float loy = floorf(__fmul_rn(__fsub_rn(py, hy), inv)), hiy = floorf(__fmul_rn(__fadd_rn(py, hy), inv));
  // compare in float (the bounds may be negative or huge), then convert
  const float x0 = fmaxf((float)r.x0, lox), x1 = fminf((float)r.x1, hix + 1.f);
  const float y0 = fmaxf((float)r.y0, loy), y1 = fminf((float)r.y1, hiy + 1.f);
  if (!(x0 < x1 && y0 < y1)) {
    r.x1 = r.x0;
    r.y1 = r.y0;
    return r;
  }
  r.x0 = (uint32_t)x0;
  r.x1 = (uint32_t)x1;
  r.y0 = (uint32_t)y0;
  r.y1 = (uint32_t)y1;
  return r;
}

// Warp-flattened binning.  The candidate tiles of the warp's 32 Gaussians (rectangles of up to
// kMaskTiles tiles) are laid end to end and tested 32 at a time with every lane busy, instead of
// each lane looping over its own rectangle.  Returns this lane's kept-tile count; for rectangles
// of <= kMaskTiles tiles `mask` gets one bit per rectangle tile (row-major) so the emit pass can
// replay the decision without re-testing.  Larger rectangles are counted cooperatively (and
// re-tested by the emit pass with the same __noinline__ routine, so both passes agree).
// `stage` = 32 x 2 float4 of per-warp shared memory the lanes publish their footprint in.
constexpr uint32_t kMaskTiles = 64;
__device__ __forceinline__ uint32_t bin_gaussians_warp(bool active, float px, float py, float ca, float cb, float cc,
                                                       float opacity, int radius, uint32_t gx, uint32_t gy, bool exact,
                                                       float4* stage, uint64_t& mask, uint32_t& rect_word) {
  const int lane = threadIdx.x & 31;
  TileRect rc{0, 0, 0, 0};
  uint32_t w = 0, area = 0;
  Footprint fp = make_footprint(1.f, 0.f, 1.f, 0.f);
  bool test = false;
  if (active) {
    rc = tile_rect(px, py, radius, gx, gy);
    if (exact) {
      fp = make_footprint(ca, cb, cc, 2.f * logf(255.f * opacity));
      test = !fp.degenerate;
      rc = shrink_rect(rc, px, py, fp);
    }
    w = rc.x1 - rc.x0;
    area = w * (rc.y1 - rc.y0);
  }
  mask = 0ull;
  const bool small = active && area <= kMaskTiles;
  // what the emit pass needs to replay the mask: x0 (13 bits) | y0 (13 bits) | w - 1 (6 bits)
  rect_word = (small && area != 0) ? (rc.x0 | (rc.y0 << 13) | ((w - 1u) << 26)) : kRectLarge;
  if (small && !test) mask = area == 64 ? ~0ull : ((1ull << area) - 1ull);  // rectangle binning: everything kept
  const uint32_t cand = (small && test) ? area : 0u;
  uint32_t incl = cand;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const uint32_t excl = incl - cand;
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  // publish what a tester needs to know about my Gaussian
  stage[2 * lane] = make_float4(px, py, fp.a, fp.b);
  stage[2 * lane + 1] = make_float4(fp.c, fp.nb_a, fp.nb_c, fp.two_tau);
  const uint32_t pack0 = rc.x0 | (rc.y0 << 16);  // tile coordinates < 65536
  const uint32_t pack1 = w | (excl << 8);          // w <= 64, excl <= 31 * 64
  __syncwarp();
  for (uint32_t w0 = 0; w0 < total; w0 += 32) {
    const uint32_t fidx = w0 + lane;
    // owner = first lane whose inclusive prefix exceeds fidx
    int lo = 0, hi = 31;  // the owner of a valid slot is one of the 32 lanes: 5 halvings settle it
#pragma unroll
    for (int step = 0; step < 5; ++step) {
      const int mid = (lo + hi) >> 1;
      const uint32_t v = __shfl_sync(0xffffffffu, incl, mid);
      if (v <= fidx)
        lo = mid + 1;
      else
        hi = mid;
    }
    const int owner = lo;
    const uint32_t o0 = __shfl_sync(0xffffffffu, pack0, owner), o1 = __shfl_sync(0xffffffffu, pack1, owner);
    bool keep = false;
    if (fidx < total) {
      const float4 q0 = stage[2 * owner], q1 = stage[2 * owner + 1];
      const uint32_t ow = o1 & 0xffu, k = fidx - (o1 >> 8);
      // k / ow for k < 64, ow <= 64: (k + 0.5) / ow is never within rounding distance of an integer
      const uint32_t row = (uint32_t)__float2int_rd(__fdividef((float)k + 0.5f, (float)ow));
      const uint32_t tx = (o0 & 0xffffu) + (k - row * ow), ty = (o0 >> 16) + row;
      Footprint f;
      f.a = q0.z;
      f.b = q0.w;
      f.c = q1.x;
      f.nb_a = q1.y;
      f.nb_c = q1.z;
      f.two_tau = q1.w;
      f.degenerate = false;
      const float x0 = (float)(tx * kTile), y0 = (float)(ty * kTile);
      keep = rect_can_contribute(q0.x, q0.y, f, x0, y0, x0 + (kTile - 1), y0 + (kTile - 1));
    }
    const unsigned votes = __ballot_sync(0xffffffffu, keep);
    // collect the votes that belong to my own rectangle
    const uint32_t my_lo = max(excl, w0), my_hi = min(excl + cand, w0 + 32);
    if (my_lo < my_hi) {
      const uint32_t nbits = my_hi - my_lo;
      const uint32_t bits = (votes >> (my_lo - w0)) & (nbits == 32 ? 0xffffffffu : ((1u << nbits) - 1u));
      mask |= (uint64_t)bits << (my_lo - excl);
    }
  }
  uint32_t kept = small ? (uint32_t)__popcll(mask) : 0u;
  // rectangles too large for a mask: whole warp per Gaussian, count only
  unsigned todo = __ballot_sync(0xffffffffu, active && !small);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const float spx = __shfl_sync(0xffffffffu, px, src), spy = __shfl_sync(0xffffffffu, py, src);
    const float sa = __shfl_sync(0xffffffffu, fp.a, src), sb = __shfl_sync(0xffffffffu, fp.b, src);
    const float sc = __shfl_sync(0xffffffffu, fp.c, src), st = __shfl_sync(0xffffffffu, fp.two_tau, src);
    const float sna = __shfl_sync(0xffffffffu, fp.nb_a, src), snc = __shfl_sync(0xffffffffu, fp.nb_c, src);
    const bool stest = __shfl_sync(0xffffffffu, (int)test, src) != 0;
    const uint32_t sx0 = __shfl_sync(0xffffffffu, rc.x0, src), sy0 = __shfl_sync(0xffffffffu, rc.y0, src);
    const uint32_t sw = __shfl_sync(0xffffffffu, w, src), sarea = __shfl_sync(0xffffffffu, area, src);
    uint32_t cnt = 0;
    for (uint32_t base = 0; base < sarea; base += 32) {
      const uint32_t k = base + lane;
      const bool keep =
          k < sarea && (!stest || tile_can_contribute(spx, spy, sa, sb, sc, sna, snc, st, sx0 + k % sw, sy0 + k / sw));
      cnt += __popc(__ballot_sync(0xffffffffu, keep));
    }
    if (lane == src) kept = cnt;
  }
  return kept;
}

// forward.cu:20-71 (computeColorFromSH): `sh(i)` = i-th float of the Gaussian's [M,3] coefficient block.
template <class Acc>
__device__ __forceinline__ void eval_sh(int D, Acc sh, float3 dir, float res[3]) {
  const float SH_C0 = 0.28209479177387814f, SH_C1 = 0.4886025119029199f;
  const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                          0.5462742152960396f};
  const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                          -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float v = SH_C0 * sh(ch);
    if (D > 0) {
      const float x = dir.x, y = dir.y, z = dir.z;
      v = v - SH_C1 * y * sh(3 + ch) + SH_C1 * z * sh(6 + ch) - SH_C1 * x * sh(9 + ch);
      if (D > 1) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        v = v + SH_C2[0] * xy * sh(12 + ch) + SH_C2[1] * yz * sh(15 + ch) + SH_C2[2] * (2.0f * zz - xx - yy) * sh(18 + ch) +
            SH_C2[3] * xz * sh(21 + ch) + SH_C2[4] * (xx - yy) * sh(24 + ch);
        if (D > 2) {
          v = v + SH_C3[0] * y * (3.0f * xx - yy) * sh(27 + ch) + SH_C3[1] * xy * z * sh(30 + ch) +
              SH_C3[2] * y * (4.0f * zz - xx - yy) * sh(33 + ch) + SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh(36 + ch) +
              SH_C3[4] * x * (4.0f * zz - xx - yy) * sh(39 + ch) + SH_C3[5] * z * (xx - yy) * sh(42 + ch) +
              SH_C3[6] * x * (xx - 3.0f * yy) * sh(45 + ch);
        }
      }
    }
    v += 0.5f;
    res[ch] = max(v, 0.0f);
  }
}

struct PreParams {
  int P, D, M, W, H;
  const float* means3D;
  const float* shs;
  const float* colors;
  const float* opacities;
  const float* scales;
  const float* rotations;
  const float* cov3D;
  float scale_modifier;
  const float* view;
  const float* proj;
  const float* campos;
  float tan_fovx, tan_fovy, focal_x, focal_y;
  uint32_t gx, gy;
  uint32_t flags;
  int use_tma;
  int sh_mode;  // staging of full-degree SH blocks (M = 16): 0 linear + scalar reads, 1 linear + 16-byte reads,
                // 2 one 192-byte bulk copy per lane into padded slots + 16-byte reads (bank-conflict free)
  float4* recA;
  float4* recB;
  float2* recC;
  uint32_t* tiles;
  int* radii;
  unsigned long long* ref_count;  // sum of reference tile rectangles
  uint32_t* depth_keys;           // view-depth sort key per Gaussian (0xffffffff = not binned)
  uint32_t* ids;                  // identity permutation, the sort's payload
  uint64_t* masks;                // kept-tile bitmask of candidate rectangles with <= 64 tiles
  uint32_t* rects;                // packed candidate rectangle the mask refers to (kRectLarge: none)
};

constexpr int kDefaultShMode = 1;                // GSB_PRE_SH overrides (A/B switch)
constexpr int kShPadStride = kMaxShFloats + 4;  // 208-byte slots: 16-byte reads of 32 lanes hit distinct banks
struct WarpStage {  // one warp's staged parameter block; every member offset is a multiple of 128 B
  float sh[32 * kShPadStride];  // 6656 B
  float rot[32 * 4];            // 512 B
  float xyz[32 * 3];            // 384 B
  float scale[32 * 3];          // 384 B
  float opac[32];               // 128 B
  uint64_t bar;
  uint64_t pad[15];
};

__global__ void __launch_bounds__(kPreThreads) preprocess_kernel(const PreParams p) {
  __shared__ __align__(128) WarpStage stage[kWarpsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpStage& st = stage[warp];
  const int base = (blockIdx.x * kWarpsPerBlock + warp) * 32;
  if (base >= p.P) return;
  const int idx = base + lane;
  const int count = min(32, p.P - base);
  const bool has_sr = p.cov3D == nullptr;
  const bool full_tma = p.use_tma && count == 32;

  // ---- stage 1: positions, scales, rotations, opacities -------------------------------------
  if (full_tma) {
    if (lane == 0) {
      mbar_init(&st.bar, 1);
      uint32_t bytes = 384 + 128 + (has_sr ? 384 + 512 : 0);
      mbar_expect_tx(&st.bar, bytes);
      tma_load_1d(st.xyz, p.means3D + (size_t)base * 3, 384, &st.bar);
      tma_load_1d(st.opac, p.opacities + base, 128, &st.bar);
      if (has_sr) {
        tma_load_1d(st.scale, p.scales + (size_t)base * 3, 384, &st.bar);
        tma_load_1d(st.rot, p.rotations + (size_t)base * 4, 512, &st.bar);
      }
    }
    __syncwarp();
    mbar_wait(&st.bar, 0);
  } else {
    for (int k = lane; k < count * 3; k += 32) {
      st.xyz[k] = p.means3D[(size_t)base * 3 + k];
      if (has_sr) st.scale[k] = p.scales[(size_t)base * 3 + k];
    }
    if (lane < count) st.opac[lane] = p.opacities[base + lane];
    if (has_sr)
      for (int k = lane; k < count * 4; k += 32) st.rot[k] = p.rotations[(size_t)base * 4 + k];
    __syncwarp();
  }

  const bool valid = lane < count;
  bool visible = false;
  float px = 0.f, py = 0.f, zv = 0.f, opacity = 0.f, ca = 0.f, cb = 0.f, cc = 0.f;
  float3 pos = make_float3(0.f, 0.f, 0.f);
  int radius = 0;
  uint32_t ntiles = 0, nref = 0;

  if (valid) {
    pos = make_float3(st.xyz[lane * 3], st.xyz[lane * 3 + 1], st.xyz[lane * 3 + 2]);
    const float* vm = p.view;
    const float* pm = p.proj;
    // auxiliary.h:139-164 (in_frustum): view-space point, near cull at 0.2
    float3 t = make_float3(vm[0] * pos.x + vm[4] * pos.y + vm[8] * pos.z + vm[12],
                           vm[1] * pos.x + vm[5] * pos.y + vm[9] * pos.z + vm[13],
                           vm[2] * pos.x + vm[6] * pos.y + vm[10] * pos.z + vm[14]);
    if (t.z > 0.2f) {
      zv = t.z;
      const float hx = pm[0] * pos.x + pm[4] * pos.y + pm[8] * pos.z + pm[12];
      const float hy = pm[1] * pos.x + pm[5] * pos.y + pm[9] * pos.z + pm[13];
      const float hw = pm[3] * pos.x + pm[7] * pos.y + pm[11] * pos.z + pm[15];
      const float p_w = 1.0f / (hw + 0.0000001f);
      const float projx = hx * p_w, projy = hy * p_w;

      // forward.cu:118-152: Sigma = (S R)^T (S R); quaternion is w-first and NOT renormalised
      float c3[6];
      if (has_sr) {
        const float4 q = reinterpret_cast<const float4*>(st.rot)[lane];
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        Mat3 S;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int i = 0; i < 3; ++i) S.m[j][i] = 0.f;
        S.m[0][0] = p.scale_modifier * st.scale[lane * 3];
        S.m[1][1] = p.scale_modifier * st.scale[lane * 3 + 1];
        S.m[2][2] = p.scale_modifier * st.scale[lane * 3 + 2];
        Mat3 R;
        R.m[0][0] = 1.f - 2.f * (y * y + z * z);
        R.m[0][1] = 2.f * (x * y - r * z);
        R.m[0][2] = 2.f * (x * z + r * y);
        R.m[1][0] = 2.f * (x * y + r * z);
        R.m[1][1] = 1.f - 2.f * (x * x + z * z);
        R.m[1][2] = 2.f * (y * z - r * x);
        R.m[2][0] = 2.f * (x * z - r * y);
        R.m[2][1] = 2.f * (y * z + r * x);
        R.m[2][2] = 1.f - 2.f * (x * x + y * y);
        const Mat3 Mx = mat_mul(S, R);
        const Mat3 Sg = mat_mul(mat_t(Mx), Mx);
        c3[0] = Sg.m[0][0];
        c3[1] = Sg.m[0][1];
        c3[2] = Sg.m[0][2];
        c3[3] = Sg.m[1][1];
        c3[4] = Sg.m[1][2];
        c3[5] = Sg.m[2][2];
      } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) c3[k] = p.cov3D[(size_t)idx * 6 + k];
      }

      // forward.cu:74-113: EWA projection with the 1.3*tan(fov) guard band and +0.3 low-pass
      const float limx = 1.3f * p.tan_fovx, limy = 1.3f * p.tan_fovy;
      const float txtz = t.x / t.z, tytz = t.y / t.z;
      t.x = min(limx, max(-limx, txtz)) * t.z;
      t.y = min(limy, max(-limy, tytz)) * t.z;
      Mat3 J;
      J.m[0][0] = p.focal_x / t.z;
      J.m[0][1] = 0.f;
      J.m[0][2] = -(p.focal_x * t.x) / (t.z * t.z);
      J.m[1][0] = 0.f;
      J.m[1][1] = p.focal_y / t.z;
      J.m[1][2] = -(p.focal_y * t.y) / (t.z * t.z);
      J.m[2][0] = 0.f;
      J.m[2][1] = 0.f;
      J.m[2][2] = 0.f;
      Mat3 Wm;
      Wm.m[0][0] = vm[0];
      Wm.m[0][1] = vm[4];
      Wm.m[0][2] = vm[8];
      Wm.m[1][0] = vm[1];
      Wm.m[1][1] = vm[5];
      Wm.m[1][2] = vm[9];
      Wm.m[2][0] = vm[2];
      Wm.m[2][1] = vm[6];
      Wm.m[2][2] = vm[10];
      const Mat3 T = mat_mul(Wm, J);
      Mat3 V;
      V.m[0][0] = c3[0];
      V.m[0][1] = c3[1];
      V.m[0][2] = c3[2];
      V.m[1][0] = c3[1];
      V.m[1][1] = c3[3];
      V.m[1][2] = c3[4];
      V.m[2][0] = c3[2];
      V.m[2][1] = c3[4];
      V.m[2][2] = c3[5];
      const Mat3 cov = mat_mul(mat_mul(mat_t(T), mat_t(V)), T);
      const float cxx = cov.m[0][0] + 0.3f, cxy = cov.m[0][1], cyy = cov.m[1][1] + 0.3f;

      const float det = (cxx * cyy - cxy * cxy);
      if (det != 0.0f) {
        const float det_inv = 1.f / det;
        ca = cyy * det_inv;
        cb = -cxy * det_inv;
        cc = cxx * det_inv;
        const float mid = 0.5f * (cxx + cyy);
        const float lambda1 = mid + sqrt(max(0.1f, mid * mid - det));
        const float lambda2 = mid - sqrt(max(0.1f, mid * mid - det));
        const float my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
        // auxiliary.h:41-44: ndc2Pix is evaluated in double
        px = (float)(((projx + 1.0) * p.W - 1.0) * 0.5);
        py = (float)(((projy + 1.0) * p.H - 1.0) * 0.5);
        const TileRect rc = tile_rect(px, py, (int)my_radius, p.gx, p.gy);
        nref = (rc.x1 - rc.x0) * (rc.y1 - rc.y0);
        if (nref != 0) {
          visible = true;
          radius = (int)my_radius;
          opacity = st.opac[lane];
        }
      }
    }
  }
  // binning, pass 1: how many tiles this Gaussian is binned into
  // The 6 KB SH block is only fetched when some lane survived culling; the bulk copy is issued NOW
  // so that it lands while the warp is busy binning.
  const bool any_visible = __any_sync(0xffffffffu, visible);
  const int shf = p.M * 3;  // SH floats per Gaussian
  const bool need_sh = p.colors == nullptr && any_visible;
  const bool tma_sh = need_sh && full_tma && (shf * 4) % 16 == 0 && shf <= kMaxShFloats;
  const int sh_mode = shf == kMaxShFloats ? p.sh_mode : 0;
  const int sh_stride = sh_mode == 2 ? kShPadStride : shf;
  if (tma_sh) {
    if (lane == 0) mbar_expect_tx(&st.bar, (uint32_t)(32 * shf * 4));
    if (sh_mode == 2) {
      __syncwarp();  // the transaction count is armed before any lane's copy can complete
      tma_load_1d(st.sh + lane * kShPadStride, p.shs + (size_t)(base + lane) * shf, (uint32_t)(shf * 4), &st.bar);
    } else if (lane == 0) {
      tma_load_1d(st.sh, p.shs + (size_t)base * shf, (uint32_t)(32 * shf * 4), &st.bar);
    }
  }
  // the staged rotations / positions / scales are dead by now: their 1280 bytes carry the lanes'
  // footprints during binning (32 x 2 float4 = 1024 bytes)
  __syncwarp();
  uint64_t tile_mask = 0ull;
  uint32_t rect_word = kRectLarge;
  ntiles = bin_gaussians_warp(visible, px, py, ca, cb, cc, opacity, radius, p.gx, p.gy,
                              (p.flags & GSB_RASTER_EXACT_TILE_CULL) != 0, reinterpret_cast<float4*>(st.rot), tile_mask,
                              rect_word);

  // ---- stage 2: colour ------------------------------------------------------------------------
  float cr = 0.f, cg = 0.f, cbl = 0.f;
  if (p.colors == nullptr) {
    if (any_visible) {
      if (tma_sh) {
        mbar_wait(&st.bar, 1);
      } else {
        for (int k = lane; k < count * shf; k += 32) st.sh[(k / shf) * sh_stride + k % shf] = p.shs[(size_t)base * shf + k];
        __syncwarp();
      }
      if (visible) {
        const float3 cam = make_float3(p.campos[0], p.campos[1], p.campos[2]);
        float3 dir = make_float3(pos.x - cam.x, pos.y - cam.y, pos.z - cam.z);
        const float len = sqrt(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
        dir.x = dir.x / len;
        dir.y = dir.y / len;
        dir.z = dir.z / len;
        float res[3];
        const float* sh = st.sh + lane * sh_stride;
        if (sh_mode != 0) {  // 12 x 16-byte reads, then the same arithmetic from registers
          float shreg[kMaxShFloats];
          const float4* s4 = reinterpret_cast<const float4*>(sh);
#pragma unroll
          for (int k = 0; k < kMaxShFloats / 4; ++k) {
            const float4 v4 = s4[k];
            shreg[4 * k] = v4.x;
            shreg[4 * k + 1] = v4.y;
            shreg[4 * k + 2] = v4.z;
            shreg[4 * k + 3] = v4.w;
          }
          eval_sh(p.D, [&](int i) { return shreg[i]; }, dir, res);
        } else {
          eval_sh(p.D, [&](int i) { return sh[i]; }, dir, res);
        }
        cr = res[0];
        cg = res[1];
        cbl = res[2];
      }
    }
  } else if (visible) {
    cr = p.colors[(size_t)idx * 3];
    cg = p.colors[(size_t)idx * 3 + 1];
    cbl = p.colors[(size_t)idx * 3 + 2];
  }

  // ---- outputs --------------------------------------------------------------------------------
  if (valid) {
    if (visible) {
      p.recA[idx] = make_float4(px, py, zv, opacity);
      p.recB[idx] = make_float4(ca, cb, cc, cr);
      p.recC[idx] = make_float2(cg, cbl);
    }
    p.tiles[idx] = ntiles;
    p.radii[idx] = radius;
    if (p.depth_keys) {
      p.depth_keys[idx] = ntiles ? __float_as_uint(zv) : 0xffffffffu;  // z_view > 0.2: bit order == float order
      p.ids[idx] = (uint32_t)idx;
      if (ntiles) {
        p.masks[idx] = tile_mask;
        p.rects[idx] = rect_word;
      }
    }
  }
  unsigned long long wsum = nref;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  if (lane == 0 && wsum) atomicAdd(p.ref_count, wsum);
}

// rasterizer_impl.cu:70-111: one (key, value) per kept (Gaussian, tile) pair, written at the
// Gaussian's prefix-sum offset in ascending tile order (the stable sort then preserves ascending
// Gaussian index among equal keys).
__global__ void __launch_bounds__(256) emit_instances_kernel(int P, const float4* __restrict__ recA,
                                                             const float4* __restrict__ recB,
                                                             const uint32_t* __restrict__ tiles,
                                                             const uint32_t* __restrict__ offsets,
                                                             const int* __restrict__ radii, uint32_t gx, uint32_t gy,
                                                             uint32_t flags, int64_t capacity,
                                                             uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const uint32_t n = tiles[idx];
  if (n == 0) return;
  const uint32_t end = offsets[idx];
  uint32_t off = end - n;
  if ((int64_t)end > capacity) return;  // host reports GSB_ERR_WORKSPACE
  const float4 A = recA[idx];
  const float4 B = recB[idx];
  TileRect rc = tile_rect(A.x, A.y, radii[idx], gx, gy);
  const uint32_t depth_bits = __float_as_uint(A.z);
  const bool exact = flags & GSB_RASTER_EXACT_TILE_CULL;
  const Footprint fp = make_footprint(B.x, B.y, B.z, exact ? 2.f * logf(255.f * A.w) : 0.f);
  const bool test = exact && !fp.degenerate;
  if (exact) rc = shrink_rect(rc, A.x, A.y, fp);
  for (uint32_t ty = rc.y0; ty < rc.y1; ++ty)
    for (uint32_t tx = rc.x0; tx < rc.x1; ++tx) {
      if (test && !tile_can_contribute(A.x, A.y, fp.a, fp.b, fp.c, fp.nb_a, fp.nb_c, fp.two_tau, tx, ty)) continue;
      if (off >= end) return;
      keys[off] = ((uint64_t)(ty * gx + tx) << 32) | depth_bits;
      vals[off] = (uint32_t)idx;
      ++off;
    }
}

// ---------------------------------------------------------------------------------------------
// Default binning pipeline: depth-sort the P Gaussians, emit their (tile, id) instances in that
// order, then split the instance stream by tile with a STABLE sort on the tile id alone.  Inside
// every tile the order is (depth bits, Gaussian index) ascending -- exactly what the reference's
// global stable radix sort on tile|depth keys emitted in Gaussian order produces
// (rasterizer_impl.cu:88-107,303) -- but the R-sized stream is sorted on <= 16 bits instead of 45,
// and nothing waits for the host.
// counters (unsigned long long[8]): [0] reference instance count  [1] binned instances R
//                                   [2] overflow flag (R > capacity)
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 8;
constexpr int kScanBlock = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_reduce_sum_1024(uint32_t v, uint32_t* smem32) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) smem32[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t t = smem32[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  return t;  // every thread holds the block total
}

// tiles-per-Gaussian in depth-sorted order (written by the depth sort's last pass): per-block sums
__global__ void __launch_bounds__(kScanThreads) sorted_block_sums_kernel(int P, const uint32_t* __restrict__ tiles_sorted,
                                                                        uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t red[32];
  uint32_t sum = 0;
  const int base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < P) sum += tiles_sorted[base + k];
  sum = block_reduce_sum_1024(sum, red);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = sum;
}

// exclusive offsets in depth-sorted order; the last block publishes R and the overflow flag
__global__ void __launch_bounds__(kScanThreads) sorted_offsets_kernel(int P, const uint32_t* __restrict__ tiles_sorted,
                                                                     const uint32_t* __restrict__ block_sums,
                                                                     uint32_t* __restrict__ offsets,
                                                                     unsigned long long* __restrict__ counters,
                                                                     int64_t capacity) {
  __shared__ uint32_t red[32];
  __shared__ uint32_t warp_sums[32];
  uint32_t prev = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += kScanThreads) prev += block_sums[b];
  const uint32_t block_base = block_reduce_sum_1024(prev, red);
  const int base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t tsum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = base + k < P ? tiles_sorted[base + k] : 0u;
    tsum += v[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t ws = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, ws, o);
      if (lane >= o) ws += u;
    }
    warp_sums[lane] = ws;
  }
  __syncthreads();
  uint32_t run = block_base + (warp ? warp_sums[warp - 1] : 0u) + incl - tsum;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < P) offsets[base + k] = run;
    run += v[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) {
    counters[1] = run;
    if ((int64_t)run > capacity) counters[2] = 1;
  }
}

// thread k = k-th Gaussian in depth order; the warp writes the (tile id, Gaussian id) instances of
// its 32 Gaussians.  Rectangles with a stored bitmask are replayed from it, flattened across the
// warp (every lane busy, stores run along the output); larger ones are re-tested cooperatively.
__global__ void __launch_bounds__(256) emit_sorted_kernel(int P, const uint32_t* __restrict__ ids_sorted,
                                                          const uint32_t* __restrict__ offsets,
                                                          const float4* __restrict__ recA, const float4* __restrict__ recB,
                                                          const uint32_t* __restrict__ tiles, const int* __restrict__ radii,
                                                          const uint64_t* __restrict__ masks,
                                                          const uint32_t* __restrict__ rects, uint32_t gx, uint32_t gy,
                                                          uint32_t flags, const unsigned long long* __restrict__ counters,
                                                          uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ tile_vals,
                                                          uint32_t* __restrict__ ghist, int hist_passes, int digit_bits) {
  // digit histograms of the tile ids this block emits (what the tile sort's passes need), so the
  // sort does not have to read the instance stream once more just to count
  __shared__ uint32_t shist[kRdxMaxPasses][kRdxBins];
  for (int p = 0; p < hist_passes; ++p) shist[p][threadIdx.x] = 0;
  __syncthreads();
  auto tally = [&](uint32_t tile) {
    for (int p = 0; p < hist_passes; ++p) atomicAdd(&shist[p][(tile >> (digit_bits * p)) & ((1u << digit_bits) - 1u)], 1u);
  };
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool fits = counters[2] == 0;
  uint32_t gid = 0, off = 0, rect = kRectLarge;
  bool active = false;
  if (k < P && fits) {
    gid = ids_sorted[k];
    active = tiles[k] != 0;  // `tiles` is in depth-sorted order here
    off = offsets[k];
  }
  uint64_t mask = 0ull;
  if (active) rect = rects[gid];
  const bool small = active && rect != kRectLarge;
  if (small) mask = masks[gid];
  // ---- flattened replay of the stored masks (bits up to the highest kept tile of each rectangle)
  const uint32_t cand = small ? 64u - (uint32_t)__clzll((long long)mask) : 0u;
  uint32_t incl = cand;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const uint32_t excl = incl - cand;
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  const uint32_t mlo = (uint32_t)mask, mhi = (uint32_t)(mask >> 32);
  for (uint32_t w0 = 0; w0 < total; w0 += 32) {
    const uint32_t fidx = w0 + lane;
    int lo = 0, hi = 31;  // the owner of a valid slot is one of the 32 lanes: 5 halvings settle it
#pragma unroll
    for (int step = 0; step < 5; ++step) {
      const int mid = (lo + hi) >> 1;
      const uint32_t v = __shfl_sync(0xffffffffu, incl, mid);
      if (v <= fidx)
        lo = mid + 1;
      else
        hi = mid;
    }
    const int owner = lo;
    const uint32_t olo = __shfl_sync(0xffffffffu, mlo, owner), ohi = __shfl_sync(0xffffffffu, mhi, owner);
    const uint32_t orect = __shfl_sync(0xffffffffu, rect, owner), oex = __shfl_sync(0xffffffffu, excl, owner);
    const uint32_t ooff = __shfl_sync(0xffffffffu, off, owner), ogid = __shfl_sync(0xffffffffu, gid, owner);
    if (fidx < total) {
      const uint32_t kk = fidx - oex;
      const uint64_t om = ((uint64_t)ohi << 32) | olo;
      if ((om >> kk) & 1ull) {
        const uint32_t ordinal = (uint32_t)__popcll(om & ((1ull << kk) - 1ull));
        const uint32_t ow = (orect >> 26) + 1u;
        // kk / ow for kk < 64, ow <= 64: (kk + 0.5) / ow is never within rounding distance of an integer
        const uint32_t row = (uint32_t)__float2int_rd(__fdividef((float)kk + 0.5f, (float)ow));
        const uint32_t tile = (((orect >> 13) & 0x1fffu) + row) * gx + (orect & 0x1fffu) + (kk - row * ow);
        tile_keys[ooff + ordinal] = tile;
        tile_vals[ooff + ordinal] = ogid;
        tally(tile);
      }
    }
  }
  // ---- rectangles without a mask: re-test, whole warp per Gaussian, ordered by ballot
  const bool exact = (flags & GSB_RASTER_EXACT_TILE_CULL) != 0;
  float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
  TileRect rc{0, 0, 0, 0};
  uint32_t w = 0, area = 0;
  Footprint fp = make_footprint(1.f, 0.f, 1.f, 0.f);
  bool test = false;
  if (active && !small) {
    A = recA[gid];
    rc = tile_rect(A.x, A.y, radii[gid], gx, gy);
    if (exact) {
      const float4 B = recB[gid];
      fp = make_footprint(B.x, B.y, B.z, 2.f * logf(255.f * A.w));
      test = !fp.degenerate;
      rc = shrink_rect(rc, A.x, A.y, fp);
    }
    w = rc.x1 - rc.x0;
    area = w * (rc.y1 - rc.y0);
  }
  const uint32_t lt_mask = (1u << lane) - 1u;
  unsigned todo = __ballot_sync(0xffffffffu, active && !small);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const float spx = __shfl_sync(0xffffffffu, A.x, src), spy = __shfl_sync(0xffffffffu, A.y, src);
    const float sa = __shfl_sync(0xffffffffu, fp.a, src), sb = __shfl_sync(0xffffffffu, fp.b, src);
    const float sc = __shfl_sync(0xffffffffu, fp.c, src), st = __shfl_sync(0xffffffffu, fp.two_tau, src);
    const float sna = __shfl_sync(0xffffffffu, fp.nb_a, src), snc = __shfl_sync(0xffffffffu, fp.nb_c, src);
    const bool stest = __shfl_sync(0xffffffffu, (int)test, src) != 0;
    const uint32_t sx0 = __shfl_sync(0xffffffffu, rc.x0, src), sy0 = __shfl_sync(0xffffffffu, rc.y0, src);
    const uint32_t sw = __shfl_sync(0xffffffffu, w, src), sarea = __shfl_sync(0xffffffffu, area, src);
    const uint32_t soff = __shfl_sync(0xffffffffu, off, src), sgid = __shfl_sync(0xffffffffu, gid, src);
    uint32_t cnt = 0;
    for (uint32_t base = 0; base < sarea; base += 32) {
      const uint32_t kk = base + lane;
      const uint32_t tx = sx0 + kk % sw, ty = sy0 + kk / sw;
      const bool keep = kk < sarea && (!stest || tile_can_contribute(spx, spy, sa, sb, sc, sna, snc, st, tx, ty));
      const unsigned votes = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const uint32_t dst = soff + cnt + __popc(votes & lt_mask);
        tile_keys[dst] = ty * gx + tx;
        tile_vals[dst] = sgid;
        tally(ty * gx + tx);
      }
      cnt += __popc(votes);
    }
  }
  __syncthreads();
  for (int p = 0; p < hist_passes; ++p) {
    const uint32_t c = shist[p][threadIdx.x];
    if (c) atomicAdd(&ghist[p * kRdxBins + threadIdx.x], c);
  }
}

__global__ void __launch_bounds__(256) init_ranges_kernel(uint2* __restrict__ ranges, uint32_t ntiles) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ntiles) ranges[t] = make_uint2(0xffffffffu, 0u);  // atomicMin / atomicMax targets; y <= x means empty
}

// rasterizer_impl.cu:116-138
__global__ void __launch_bounds__(256) tile_ranges_kernel(int64_t L, const uint64_t* __restrict__ keys,
                                                          uint2* __restrict__ ranges) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L) return;
  const uint32_t cur = (uint32_t)(keys[idx] >> 32);
  if (idx == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
    if (cur != prev) {
      ranges[prev].y = (uint32_t)idx;
      ranges[cur].x = (uint32_t)idx;
    }
  }
  if (idx == L - 1) ranges[cur].y = (uint32_t)L;
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, const float2& v) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// forward.cu:261-374, re-organised:
//  * batches of 256 records (position/depth/opacity, conic/red, green/blue) are staged in a
//    DOUBLE-buffered shared-memory ring: while batch i is blended, batch i+1 is already in the other
//    buffer and batch i+2 is in flight in registers -> one block barrier per batch, no global
//    access inside the blend loop;
//  * each warp owns an 8x4 pixel block of the 16x16 tile and first asks, 32 Gaussians at a time
//    (one per lane, exact footprint test), which ones can reach alpha >= 1/255 anywhere in its
//    block; only those are evaluated per pixel.  Skipped Gaussians are exactly the ones every pixel
//    of the block would `continue` past in the reference loop, so the result is bit-identical;
//  * an expected-depth accumulator (sum z*alpha*T) runs next to the colour.
// kFastExp: alpha = opacity * ex2.approx(power * log2 e) instead of the reference's full-precision
// expf (forward.cu:340): ~2e-7 relative on alpha, far inside the 1e-4 parity budget.
template <bool kFastExp>
__global__ void __launch_bounds__(kTilePixels) render_kernel(const uint2* __restrict__ ranges,
                                                            const uint32_t* __restrict__ point_list, int W, int H,
                                                            const float4* __restrict__ recA,
                                                            const float4* __restrict__ recB,
                                                            const float2* __restrict__ recC,
                                                            const float* __restrict__ bg, float* __restrict__ out_color,
                                                            float* __restrict__ out_depth, float* __restrict__ out_T,
                                                            int64_t capacity) {
  __shared__ __align__(16) float4 sA[2][kTilePixels];
  __shared__ __align__(16) float4 sB[2][kTilePixels];
  __shared__ __align__(16) float2 sC[2][kTilePixels];
  const uint32_t aA = smem_u32(&sA[0][0]), aB = smem_u32(&sB[0][0]), aC = smem_u32(&sC[0][0]);
  constexpr uint32_t kStrideAB = kTilePixels * 16, kStrideC = kTilePixels * 8;
  const uint32_t tiles_x = (W + kTile - 1) / kTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // warp -> 8x4 block inside the tile: 2 blocks across, 4 down
  const uint32_t blk_x = blockIdx.x * kTile + (warp & 1) * 8, blk_y = blockIdx.y * kTile + (warp >> 1) * 4;
  const uint32_t pix_x = blk_x + (lane & 7), pix_y = blk_y + (lane >> 3);
  const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
  const float pfx = (float)pix_x, pfy = (float)pix_y;
  const float bx0 = (float)blk_x, by0 = (float)blk_y, bx1 = (float)(blk_x + 7), by1 = (float)(blk_y + 3);
  uint2 range = ranges[blockIdx.y * tiles_x + blockIdx.x];
  if (range.y <= range.x || (int64_t)range.y > capacity) range = make_uint2(0u, 0u);  // empty tile / invalid frame
  const int total = range.y - range.x;
  const int rounds = (total + kTilePixels - 1) / kTilePixels;
  bool done = !inside;
  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dz = 0.f;

  float4 nA = make_float4(0.f, 0.f, 0.f, 0.f), nB = nA;
  float2 nC = make_float2(0.f, 0.f);
  auto fetch = [&](int batch_index) {
    const int slot = batch_index * kTilePixels + (int)threadIdx.x;
    if (slot < total) {
      const uint32_t g = point_list[range.x + slot];
      nA = recA[g];
      nB = recB[g];
      nC = recC[g];
    }
  };
  auto stage = [&](int buf) {
    sts128(aA + buf * kStrideAB + threadIdx.x * 16, nA);
    sts128(aB + buf * kStrideAB + threadIdx.x * 16, nB);
    sts64(aC + buf * kStrideC + threadIdx.x * 8, nC);
  };
  if (rounds > 0) {
    fetch(0);
    stage(0);
    fetch(1);
  }
  __syncthreads();
  for (int i = 0; i < rounds; ++i) {
    const int cur = i & 1;
    if (i + 1 < rounds) {  // buffer cur^1 was released by the barrier that ended iteration i-1
      stage(cur ^ 1);
      fetch(i + 2);
    }
    const uint32_t bA = aA + cur * kStrideAB, bB = aB + cur * kStrideAB, bC = aC + cur * kStrideC;
    const int batch = min(kTilePixels, total - i * kTilePixels);
    bool warp_done = __all_sync(0xffffffffu, done);
    for (int c0 = 0; c0 < batch && !warp_done; c0 += 32) {
      const int jt = c0 + lane;
      bool hit = false;
      if (jt < batch) {
        const float4 A = lds128(bA + jt * 16);
        const float4 B = lds128(bB + jt * 16);
        const Footprint fp = make_footprint(B.x, B.y, B.z, 2.f * __logf(255.f * A.w) + 1e-3f);
        hit = rect_can_contribute(A.x, A.y, fp, bx0, by0, bx1, by1);
      }
      unsigned todo = __ballot_sync(0xffffffffu, hit);
      while (todo) {
        const int j = c0 + __ffs(todo) - 1;
        todo &= todo - 1;
        if (done) continue;
        const float4 A = lds128(bA + j * 16);
        const float4 B = lds128(bB + j * 16);
        const float dx = A.x - pfx, dy = A.y - pfy;
        const float power = -0.5f * (B.x * dx * dx + B.z * dy * dy) - B.y * dx * dy;
        if (power > 0.0f) continue;
        const float alpha = min(0.99f, A.w * (kFastExp ? __expf(power) : exp(power)));
        if (alpha < 1.0f / 255.0f) continue;
        const float test_T = T * (1 - alpha);
        if (test_T < 0.0001f) {
          done = true;
          continue;
        }
        const float2 gb = lds64(bC + j * 8);
        C0 += B.w * alpha * T;
        C1 += gb.x * alpha * T;
        C2 += gb.y * alpha * T;
        Dz += A.z * alpha * T;
        T = test_T;
      }
      warp_done = __all_sync(0xffffffffu, done);
    }
    // one barrier per batch: batch i+1 is now visible, buffer `cur` may be overwritten, and the
    // tile stops as soon as every pixel is saturated
    if (__syncthreads_count(done) == kTilePixels) break;
  }
  if (inside) {
    const size_t pid = (size_t)pix_y * W + pix_x;
    const size_t plane = (size_t)H * W;
    out_color[pid] = C0 + T * bg[0];
    out_color[plane + pid] = C1 + T * bg[1];
    out_color[2 * plane + pid] = C2 + T * bg[2];
    if (out_depth) out_depth[pid] = Dz;
    if (out_T) out_T[pid] = T;
  }
}

// Barrier-free variant: the 8 warps of a tile never wait for each other.  Every warp streams the
// tile's sorted instance list on its own, 32 Gaussians at a time: each lane gathers ONE record
// (the 8 warps of the CTA hit the same lines, so all but the first gather is an L1 hit), tests it
// against the warp's 8x4 pixel block straight from registers, and publishes it in the warp's private
// double-buffered shared-memory slot for the broadcast reads of the blend loop.  The next chunk's
// gathers are in flight while the current one is blended; no __syncthreads anywhere, and a warp
// leaves as soon as its own 32 pixels are saturated.
template <bool kFastExp>
__global__ void __launch_bounds__(kTilePixels, 5) render_warp_kernel(const uint2* __restrict__ ranges,
                                                                 const uint32_t* __restrict__ point_list, int W, int H,
                                                                 const float4* __restrict__ recA,
                                                                 const float4* __restrict__ recB,
                                                                 const float2* __restrict__ recC,
                                                                 const float* __restrict__ bg, float* __restrict__ out_color,
                                                                 float* __restrict__ out_depth, float* __restrict__ out_T,
                                                                 int64_t capacity) {
  __shared__ __align__(16) float4 sA[8][2][32];
  __shared__ __align__(16) float4 sB[8][2][32];
  __shared__ __align__(16) float2 sC[8][2][32];
  const uint32_t tiles_x = (W + kTile - 1) / kTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t aA = smem_u32(&sA[warp][0][0]), aB = smem_u32(&sB[warp][0][0]), aC = smem_u32(&sC[warp][0][0]);
  const uint32_t blk_x = blockIdx.x * kTile + (warp & 1) * 8, blk_y = blockIdx.y * kTile + (warp >> 1) * 4;
  const uint32_t pix_x = blk_x + (lane & 7), pix_y = blk_y + (lane >> 3);
  const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
  const float pfx = (float)pix_x, pfy = (float)pix_y;
  const float bx0 = (float)blk_x, by0 = (float)blk_y, bx1 = (float)(blk_x + 7), by1 = (float)(blk_y + 3);
  uint2 range = ranges[blockIdx.y * tiles_x + blockIdx.x];
  if (range.y <= range.x || (int64_t)range.y > capacity) range = make_uint2(0u, 0u);
  const int total = range.y - range.x;
  bool done = !inside;
  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dz = 0.f;

  float4 cA = make_float4(0.f, 0.f, 0.f, 0.f), cB = cA, nA = cA, nB = cA;
  float2 cC = make_float2(0.f, 0.f), nC = cC;
  if (lane < total) {
    const uint32_t g = point_list[range.x + lane];
    cA = recA[g];
    cB = recB[g];
    cC = recC[g];
  }
  for (int c0 = 0; c0 < total; c0 += 32) {
    if (__all_sync(0xffffffffu, done)) break;
    const int buf = (c0 >> 5) & 1;
    sts128(aA + (buf * 32 + lane) * 16, cA);
    sts128(aB + (buf * 32 + lane) * 16, cB);
    sts64(aC + (buf * 32 + lane) * 8, cC);
    const int nxt = c0 + 32 + lane;
    if (nxt < total) {  // next chunk's gathers fly while this one is blended
      const uint32_t g = point_list[range.x + nxt];
      nA = recA[g];
      nB = recB[g];
      nC = recC[g];
    }
    bool hit = false;
    if (c0 + lane < total) {
      const Footprint fp = make_footprint(cB.x, cB.y, cB.z, 2.f * __logf(255.f * cA.w) + 1e-3f);
      hit = rect_can_contribute(cA.x, cA.y, fp, bx0, by0, bx1, by1);
    }
    __syncwarp();  // the chunk's records are visible to every lane of the warp
    unsigned todo = __ballot_sync(0xffffffffu, hit);
    const uint32_t bA = aA + buf * 512, bB = aB + buf * 512, bC = aC + buf * 256;
    // two hits per trip: their loads, quadratic forms and exponentials are independent and overlap; only
    // the transmittance update is applied in order
    while (todo) {
      const int j0 = __ffs(todo) - 1;
      todo &= todo - 1;
      const bool two = todo != 0;
      const int j1 = two ? __ffs(todo) - 1 : j0;
      todo &= todo - 1;  // no-op when todo is already 0
      if (done) continue;
      const float4 A0 = lds128(bA + j0 * 16), B0 = lds128(bB + j0 * 16);
      const float4 A1 = lds128(bA + j1 * 16), B1 = lds128(bB + j1 * 16);
      const float dx0 = A0.x - pfx, dy0 = A0.y - pfy, dx1 = A1.x - pfx, dy1 = A1.y - pfy;
      const float power0 = -0.5f * (B0.x * dx0 * dx0 + B0.z * dy0 * dy0) - B0.y * dx0 * dy0;
      const float power1 = -0.5f * (B1.x * dx1 * dx1 + B1.z * dy1 * dy1) - B1.y * dx1 * dy1;
      const float alpha0 = min(0.99f, A0.w * (kFastExp ? __expf(power0) : exp(power0)));
      const float alpha1 = min(0.99f, A1.w * (kFastExp ? __expf(power1) : exp(power1)));
      if (!(power0 > 0.0f) && !(alpha0 < 1.0f / 255.0f)) {
        const float test_T = T * (1 - alpha0);
        if (test_T < 0.0001f) {
          done = true;
        } else {
          const float2 gb = lds64(bC + j0 * 8);
          C0 += B0.w * alpha0 * T;
          C1 += gb.x * alpha0 * T;
          C2 += gb.y * alpha0 * T;
          Dz += A0.z * alpha0 * T;
          T = test_T;
        }
      }
      if (two && !done && !(power1 > 0.0f) && !(alpha1 < 1.0f / 255.0f)) {
        const float test_T = T * (1 - alpha1);
        if (test_T < 0.0001f) {
          done = true;
        } else {
          const float2 gb = lds64(bC + j1 * 8);
          C0 += B1.w * alpha1 * T;
          C1 += gb.x * alpha1 * T;
          C2 += gb.y * alpha1 * T;
          Dz += A1.z * alpha1 * T;
          T = test_T;
        }
      }
    }
    cA = nA;
    cB = nB;
    cC = nC;
  }
  if (inside) {
    const size_t pid = (size_t)pix_y * W + pix_x;
    const size_t plane = (size_t)H * W;
    out_color[pid] = C0 + T * bg[0];
    out_color[plane + pid] = C1 + T * bg[1];
    out_color[2 * plane + pid] = C2 + T * bg[2];
    if (out_depth) out_depth[pid] = Dz;
    if (out_T) out_T[pid] = T;
  }
}

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Compacting variant of render_warp_kernel: the lanes whose Gaussian passed the warp's footprint test
// store their record at consecutive 48-byte slots (ballot rank) of the warp's private list, so the
// blend loop walks plain consecutive addresses (no bit scan, no per-hit address arithmetic) and is
// branch-free: the three thresholds of forward.cu:336-353 become predicates that gate a weight
// w = alpha*T (0 when the pixel skips the Gaussian) and the transmittance update.  An all-zero
// sentinel record (opacity 0 -> alpha 0 -> skipped) pads an odd hit count so the loop always takes
// two records per trip.  Thresholds are decided on exactly the same alpha / T values as in
// render_warp_kernel; colours accumulate as fma(c, alpha*T, C) instead of fma(c*alpha, T, C)
// (<= 1 ulp per term).
constexpr int kDefaultRenderImpl = 3;        // 0 block, 1 warp, 2 compact, 3 compact with two pixels per lane (GSB_RENDER_IMPL overrides)
constexpr int kSlotBytes = 48;               // A (16) | B (16) | green, blue (8) | pad (8)
constexpr int kSlotsPerBuf = 33;             // 32 hits + sentinel
// kPix = pixels per lane: 1 -> 8 warps per tile, each an 8x4 block; 2 -> 4 warps per tile, each an 8x8 block whose
// lanes own (x, y) and (x, y + 4): the list walk, the footprint test and the record loads are shared by 64 pixels.
template <bool kFastExp, int kPix>
__global__ void __launch_bounds__(kTilePixels / kPix, kPix == 1 ? 5 : 8)
    render_compact_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
                          const float4* __restrict__ recA, const float4* __restrict__ recB, const float2* __restrict__ recC,
                          const float* __restrict__ bg, float* __restrict__ out_color, float* __restrict__ out_depth,
                          float* __restrict__ out_T, int64_t capacity) {
  constexpr int kWarps = 8 / kPix;
  __shared__ __align__(16) unsigned char slots[kWarps][2][kSlotsPerBuf * kSlotBytes];
  const uint32_t tiles_x = (W + kTile - 1) / kTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t a0 = smem_u32(&slots[warp][0][0]);
  const uint32_t blk_x = blockIdx.x * kTile + (warp & 1) * 8, blk_y = blockIdx.y * kTile + (warp >> 1) * (4 * kPix);
  const uint32_t pix_x = blk_x + (lane & 7), pix_y = blk_y + (lane >> 3);
  const float pfx = (float)pix_x;
  const float bx0 = (float)blk_x, by0 = (float)blk_y, bx1 = (float)(blk_x + 7), by1 = (float)(blk_y + 4 * kPix - 1);
  uint2 range = ranges[blockIdx.y * tiles_x + blockIdx.x];
  if (range.y <= range.x || (int64_t)range.y > capacity) range = make_uint2(0u, 0u);
  const int total = range.y - range.x;
  const uint32_t lt_mask = (1u << lane) - 1u;
  bool inside[kPix], done[kPix];
  float pfy[kPix], T[kPix], C0[kPix], C1[kPix], C2[kPix], Dz[kPix];
#pragma unroll
  for (int k = 0; k < kPix; ++k) {
    inside[k] = pix_x < (uint32_t)W && pix_y + 4 * k < (uint32_t)H;
    done[k] = !inside[k];
    pfy[k] = (float)(pix_y + 4 * k);
    T[k] = 1.0f;
    C0[k] = C1[k] = C2[k] = Dz[k] = 0.f;
  }

  float4 cA = make_float4(0.f, 0.f, 0.f, 0.f), cB = cA, nA = cA, nB = cA;
  float2 cC = make_float2(0.f, 0.f), nC = cC;
  if (lane < total) {
    const uint32_t g = point_list[range.x + lane];
    cA = recA[g];
    cB = recB[g];
    cC = recC[g];
  }
  // one Gaussian of the warp's list against this lane's pixel(s)
  auto blend = [&](uint32_t addr) {
    const float4 A = lds128(addr), B = lds128(addr + 16);
    const float2 gb = lds64(addr + 32);
    const float dx = A.x - pfx;
#pragma unroll
    for (int k = 0; k < kPix; ++k) {
      const float dy = A.y - pfy[k];
      const float power = -0.5f * (B.x * dx * dx + B.z * dy * dy) - B.y * dx * dy;
      const float e = kFastExp ? ex2_ftz(power * 1.4426950408889634f) : exp(power);
      const float alpha = min(0.99f, A.w * e);
      const float test_T = T[k] * (1 - alpha);
      const bool cand = !done[k] && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
      const bool sat = cand && test_T < 0.0001f;
      const bool ok = cand && !sat;
      done[k] = done[k] || sat;
      const float w = ok ? alpha * T[k] : 0.f;
      C0[k] = fmaf(B.w, w, C0[k]);
      C1[k] = fmaf(gb.x, w, C1[k]);
      C2[k] = fmaf(gb.y, w, C2[k]);
      Dz[k] = fmaf(A.z, w, Dz[k]);
      T[k] = ok ? test_T : T[k];
    }
  };
  for (int c0 = 0; c0 < total; c0 += 32) {
    bool all_done = done[0];
#pragma unroll
    for (int k = 1; k < kPix; ++k) all_done = all_done && done[k];
    if (__all_sync(0xffffffffu, all_done)) break;
    const int nxt = c0 + 32 + lane;
    if (nxt < total) {  // next chunk's gathers fly while this one is tested and blended
      const uint32_t g = point_list[range.x + nxt];
      nA = recA[g];
      nB = recB[g];
      nC = recC[g];
    }
    bool hit = false;
    if (c0 + lane < total) {
      const Footprint fp = make_footprint(cB.x, cB.y, cB.z, 2.f * __logf(255.f * cA.w) + 1e-3f);
      hit = rect_can_contribute(cA.x, cA.y, fp, bx0, by0, bx1, by1);
    }
    const unsigned votes = __ballot_sync(0xffffffffu, hit);
    const int n = __popc(votes);
    const uint32_t buf = a0 + ((c0 >> 5) & 1) * (kSlotsPerBuf * kSlotBytes);
    if (hit) {
      const uint32_t dst = buf + __popc(votes & lt_mask) * kSlotBytes;
      sts128(dst, cA);
      sts128(dst + 16, cB);
      sts64(dst + 32, cC);
    }
    if (lane == 0) {  // sentinel after the last hit
      const uint32_t dst = buf + n * kSlotBytes;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      sts128(dst, z);
      sts128(dst + 16, z);
      sts64(dst + 32, make_float2(0.f, 0.f));
    }
    __syncwarp();  // the list is visible to every lane of the warp
    uint32_t addr = buf;
    for (int j = 0; j < n; j += 2, addr += 2 * kSlotBytes) {
      blend(addr);
      blend(addr + kSlotBytes);
    }
    cA = nA;
    cB = nB;
    cC = nC;
  }
  const size_t plane = (size_t)H * W;
#pragma unroll
  for (int k = 0; k < kPix; ++k)
    if (inside[k]) {
      const size_t pid = (size_t)(pix_y + 4 * k) * W + pix_x;
      out_color[pid] = C0[k] + T[k] * bg[0];
      out_color[plane + pid] = C1[k] + T[k] * bg[1];
      out_color[2 * plane + pid] = C2[k] + T[k] * bg[2];
      if (out_depth) out_depth[pid] = Dz[k];
      if (out_T) out_T[pid] = T[k];
    }
}

__global__ void write_counts_kernel(const uint32_t* __restrict__ offsets, int P, const unsigned long long* counters,
                                    int64_t* out) {
  // offsets != NULL: validation path (instance total = last element of the per-Gaussian scan)
  out[0] = offsets ? (P > 0 ? (int64_t)offsets[P - 1] : 0) : (int64_t)counters[1];
  out[1] = (int64_t)counters[0];
  out[2] = (int64_t)counters[2];
  out[3] = 0;
}

__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* __restrict__ means,
                                                           const float* __restrict__ vm, uint8_t* present) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float x = means[3 * (size_t)idx], y = means[3 * (size_t)idx + 1], z = means[3 * (size_t)idx + 2];
  present[idx] = (vm[2] * x + vm[6] * y + vm[10] * z + vm[14]) > 0.2f ? 1 : 0;
}

// cv::saturate_cast<uchar>(v): round half to even, clamp to [0,255]
__global__ void __launch_bounds__(256) to_u8_kernel(const float* __restrict__ chw, int W, int H, uint8_t* __restrict__ hwc) {
  const size_t n = (size_t)W * H;
  const size_t pid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pid >= n) return;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float v = chw[ch * n + pid] * 255.f;
    int q = __float2int_rn(v);
    q = q < 0 ? 0 : (q > 255 ? 255 : q);
    if (!(v == v)) q = 0;
    hwc[pid * 3 + ch] = (uint8_t)q;
  }
}

// rasterizer_impl.cu:35-50
uint32_t higher_msb(uint32_t n) {
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

struct Workspace {
  float4* recA;
  float4* recB;
  float2* recC;
  uint32_t* tiles;
  uint32_t* offsets;
  int* radii;
  unsigned long long* counters;  // [0] reference instance count
  uint2* ranges;
  uint32_t* depth_a;
  uint32_t* depth_b;
  uint32_t* ids_a;
  uint32_t* ids_b;
  uint32_t* sorted_offsets;
  uint64_t* masks;
  uint32_t* rects;
  uint32_t* block_sums;
  uint32_t* radix_scratch;
  uint64_t* keys_in;
  uint64_t* keys_out;
  uint32_t* vals_in;
  uint32_t* vals_out;
  void* scan_temp;
  size_t scan_temp_bytes;
  void* sort_temp;
  size_t sort_temp_bytes;
  size_t total;
};

Workspace carve(void* base, int32_t P, int32_t W, int32_t H, int64_t R) {
  Workspace w{};
  Carver c(base);
  const size_t Pn = (size_t)(P > 1 ? P : 1), Rn = (size_t)(R > 1 ? R : 1);
  const size_t ntiles = (size_t)((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
  w.recA = c.take<float4>(Pn);
  w.recB = c.take<float4>(Pn);
  w.recC = c.take<float2>(Pn);
  w.tiles = c.take<uint32_t>(Pn);
  w.offsets = c.take<uint32_t>(Pn);
  w.radii = c.take<int>(Pn);
  w.counters = c.take<unsigned long long>(8);
  w.ranges = c.take<uint2>(ntiles);
  w.depth_a = c.take<uint32_t>(Pn);
  w.depth_b = c.take<uint32_t>(Pn);
  w.ids_a = c.take<uint32_t>(Pn);
  w.ids_b = c.take<uint32_t>(Pn);
  w.sorted_offsets = c.take<uint32_t>(Pn);
  w.masks = c.take<uint64_t>(Pn);
  w.rects = c.take<uint32_t>(Pn);
  w.block_sums = c.take<uint32_t>((Pn + kScanBlock - 1) / kScanBlock + 1);
  w.radix_scratch = c.take<uint32_t>(radix_scratch_words(std::max(Pn, Rn)));
  w.keys_in = c.take<uint64_t>(Rn);
  w.keys_out = c.take<uint64_t>(Rn);
  w.vals_in = c.take<uint32_t>(Rn);
  w.vals_out = c.take<uint32_t>(Rn);
  cub::DeviceScan::InclusiveSum(nullptr, w.scan_temp_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)Pn);
  w.scan_temp = c.take<char>(w.scan_temp_bytes);
  cub::DeviceRadixSort::SortPairs(nullptr, w.sort_temp_bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int64_t)Rn);
  w.sort_temp = c.take<char>(w.sort_temp_bytes);
  w.total = c.total();
  return w;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace gsb

using namespace gsb;

extern "C" {

const char* gsb_last_error(void) { return g_error; }
int gsb_version(void) { return GSB_VERSION; }
uint64_t gsb_kernel_launch_count(void) { return g_launches.load(); }
int64_t gsb_raster_required_instances(void) { return g_required_instances; }

int gsb_profile_num_stages(void) { return kStCount; }
const char* gsb_profile_stage_name(int stage) { return stage >= 0 && stage < kStCount ? kStageNames[stage] : ""; }
int gsb_profile_enable(int on) {
  g_prof.enabled.store(on != 0);
  return GSB_OK;
}
int gsb_profile_collect(double* total_ms, uint64_t* samples, int n_stages) {
  if (!total_ms || !samples || n_stages < kStCount) return fail(GSB_ERR_INVALID, "profile_collect: need %d stages", (int)kStCount);
  GSB_CUDA_OK(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof.mu);
  for (int i = 0; i < n_stages; ++i) {
    total_ms[i] = 0.0;
    samples[i] = 0;
  }
  for (const Profiler::Rec& r : g_prof.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      total_ms[r.stage] += ms;
      samples[r.stage] += 1;
    }
    g_prof.pool.push_back(r.a);
    g_prof.pool.push_back(r.b);
  }
  g_prof.recs.clear();
  return GSB_OK;
}

size_t gsb_raster_workspace_bytes(int32_t P, int32_t width, int32_t height, int64_t max_instances) {
  return carve(nullptr, P, width, height, max_instances).total;
}

int gsb_raster_forward(const GsbRasterArgs* a, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!a) return fail(GSB_ERR_INVALID, "args is NULL");
  if (a->P < 0 || a->width <= 0 || a->height <= 0) return fail(GSB_ERR_INVALID, "bad P / image size");
  if (a->width > 8191 * kTile || a->height > 8191 * kTile)  // packed candidate rectangles hold 13-bit tile coordinates
    return fail(GSB_ERR_INVALID, "image side exceeds %d pixels", 8191 * kTile);
  if (!a->out_color || !a->background || !a->viewmatrix || !a->projmatrix || !a->cam_pos)
    return fail(GSB_ERR_INVALID, "out_color, background, viewmatrix, projmatrix and cam_pos are required");
  // DGR/diff_gaussian_rasterization/__init__.py:191-195
  if ((a->shs == nullptr) == (a->colors_precomp == nullptr))
    return fail(GSB_ERR_INVALID, "Please provide excatly one of either SHs or precomputed colors!");
  const bool has_sr = a->scales != nullptr && a->rotations != nullptr;
  if (((a->scales == nullptr || a->rotations == nullptr) && a->cov3D_precomp == nullptr) ||
      ((a->scales != nullptr || a->rotations != nullptr) && a->cov3D_precomp != nullptr))
    return fail(GSB_ERR_INVALID,
                "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
  if (a->shs && (a->sh_coeffs <= 0 || a->sh_coeffs > 16 || a->sh_degree < 0 || a->sh_degree > 3 ||
                 (a->sh_degree + 1) * (a->sh_degree + 1) > a->sh_coeffs))
    return fail(GSB_ERR_INVALID, "SH degree %d needs %d coefficients, tensor has %d (max 16)", a->sh_degree,
                (a->sh_degree + 1) * (a->sh_degree + 1), a->sh_coeffs);
  if (!a->workspace) return fail(GSB_ERR_WORKSPACE, "workspace is NULL");
  const int P = a->P, W = a->width, H = a->height;
  const int64_t cap = a->max_instances;
  Workspace ws = carve(a->workspace, P, W, H, cap);
  if (ws.total > a->workspace_bytes)
    return fail(GSB_ERR_WORKSPACE, "workspace has %zu bytes, %zu needed", a->workspace_bytes, ws.total);
  const bool dbg = a->flags & GSB_RASTER_DEBUG_SYNC;
  const uint32_t gx = (W + kTile - 1) / kTile, gy = (H + kTile - 1) / kTile;
  const size_t npix = (size_t)W * H;
  int rc;

  if (P == 0) {  // rasterize_points.cu:77: outputs stay zero-filled
    GSB_CUDA_OK(cudaMemsetAsync(a->out_color, 0, 3 * npix * sizeof(float), stream));
    if (a->out_depth) GSB_CUDA_OK(cudaMemsetAsync(a->out_depth, 0, npix * sizeof(float), stream));
    if (a->out_final_T) GSB_CUDA_OK(cudaMemsetAsync(a->out_final_T, 0, npix * sizeof(float), stream));
    if (a->num_rendered) GSB_CUDA_OK(cudaMemsetAsync(a->num_rendered, 0, 4 * sizeof(int64_t), stream));
    return GSB_OK;
  }

  PreParams pp{};
  pp.P = P;
  pp.D = a->sh_degree;
  pp.M = a->sh_coeffs;
  pp.W = W;
  pp.H = H;
  pp.means3D = a->means3D;
  pp.shs = a->shs;
  pp.colors = a->colors_precomp;
  pp.opacities = a->opacities;
  pp.scales = a->scales;
  pp.rotations = a->rotations;
  pp.cov3D = a->cov3D_precomp;
  pp.scale_modifier = a->scale_modifier;
  pp.view = a->viewmatrix;
  pp.proj = a->projmatrix;
  pp.campos = a->cam_pos;
  pp.tan_fovx = a->tan_fovx;
  pp.tan_fovy = a->tan_fovy;
  pp.focal_y = H / (2.0f * a->tan_fovy);  // rasterizer_impl.cu:222-223
  pp.focal_x = W / (2.0f * a->tan_fovx);
  pp.gx = gx;
  pp.gy = gy;
  pp.flags = a->flags;
  pp.use_tma = !(a->flags & GSB_RASTER_NO_TMA) && aligned16(a->means3D) && aligned16(a->opacities) &&
               (!has_sr || (aligned16(a->scales) && aligned16(a->rotations))) && (!a->shs || aligned16(a->shs));
  static const int sh_mode_default = [] {
    const char* e = getenv("GSB_PRE_SH");  // A/B switch: "scalar" | "vec" | "padded"
    if (e && e[0] == 's') return 0;
    if (e && e[0] == 'v') return 1;
    if (e && e[0] == 'p') return 2;
    return kDefaultShMode;
  }();
  pp.sh_mode = ((a->flags >> 12) & 3u) ? (int)((a->flags >> 12) & 3u) - 1 : sh_mode_default;
  pp.recA = ws.recA;
  pp.recB = ws.recB;
  pp.recC = ws.recC;
  pp.tiles = ws.tiles;
  pp.radii = a->radii ? a->radii : ws.radii;
  pp.ref_count = ws.counters;

  const bool use_cub = (a->flags & GSB_RASTER_CUB_SORT) != 0;
  const size_t ntiles = (size_t)gx * gy;
  pp.depth_keys = use_cub ? nullptr : ws.depth_a;
  pp.ids = use_cub ? nullptr : ws.ids_a;
  pp.masks = ws.masks;
  pp.rects = ws.rects;
  GSB_CUDA_OK(cudaMemsetAsync(ws.counters, 0, 8 * sizeof(unsigned long long), stream));
  const int pre_blocks = (P + kPreThreads - 1) / kPreThreads;
  {
    StageTimer tm(kStPreprocess, stream);
    preprocess_kernel<<<pre_blocks, kPreThreads, 0, stream>>>(pp);
  }
  count_launch();
  if ((rc = check_launch("preprocess_kernel", stream, dbg))) return rc;

  const uint32_t* point_list = nullptr;
  if (use_cub) {
    // ---- validation path: the reference's own pipeline shape (P-sized scan, emit in Gaussian order,
    //      global radix sort of (tile | depth) keys, boundary detection), incl. its host round trip
    {
      StageTimer tm(kStScan, stream);
      GSB_CUDA_OK(cub::DeviceScan::InclusiveSum(ws.scan_temp, ws.scan_temp_bytes, ws.tiles, ws.offsets, P, stream));
    }
    if (a->num_rendered) {
      write_counts_kernel<<<1, 1, 0, stream>>>(ws.offsets, P, ws.counters, a->num_rendered);
      count_launch();
    }
    uint32_t R32 = 0;
    GSB_CUDA_OK(cudaMemcpyAsync(&R32, ws.offsets + (P - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    GSB_CUDA_OK(cudaStreamSynchronize(stream));
    const int64_t R = (int64_t)R32;
    g_required_instances = R;
    if (R > cap)
      return fail(GSB_ERR_WORKSPACE, "frame needs %lld instances, workspace sized for %lld", (long long)R, (long long)cap);
    if (R > 0) {
      {
        StageTimer tm(kStEmit, stream);
        emit_instances_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, ws.recA, ws.recB, ws.tiles, ws.offsets, pp.radii, gx, gy,
                                                                   a->flags, cap, ws.keys_in, ws.vals_in);
      }
      count_launch();
      if ((rc = check_launch("emit_instances_kernel", stream, dbg))) return rc;
      const int bit = (int)higher_msb(gx * gy);
      {
        StageTimer tm(kStSort, stream);
        GSB_CUDA_OK(cub::DeviceRadixSort::SortPairs(ws.sort_temp, ws.sort_temp_bytes, ws.keys_in, ws.keys_out, ws.vals_in,
                                                    ws.vals_out, R, 0, 32 + bit, stream));
      }
      if ((rc = check_launch("radix sort", stream, dbg))) return rc;
    }
    GSB_CUDA_OK(cudaMemsetAsync(ws.ranges, 0, ntiles * sizeof(uint2), stream));
    if (R > 0) {
      {
        StageTimer tm(kStRanges, stream);
        tile_ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(R, ws.keys_out, ws.ranges);
      }
      count_launch();
      if ((rc = check_launch("tile_ranges_kernel", stream, dbg))) return rc;
    }
    point_list = ws.vals_out;
  } else {
    // ---- default pipeline: depth-sort P, emit in that order, stable split by tile; nothing waits
    uint32_t* rs = ws.radix_scratch;
    uint64_t nl = 0;
    const uint32_t* ids_sorted;
    {
      StageTimer tm(kStScan, stream);
      // the last pass also delivers tiles-per-Gaussian in sorted order (ws.offsets is free on this path), so the
      // scan and the emit pass read it contiguously instead of gathering tiles[ids_sorted[k]]
      uint32_t* tiles_sorted = ws.offsets;
      const int which = radix_sort_pairs(ws.depth_a, ws.ids_a, ws.depth_b, ws.ids_b, (uint32_t)P, nullptr, 0, (size_t)P, 32, rs,
                                         nullptr, stream, &nl, /*histogram_ready=*/false, ws.tiles, tiles_sorted);
      ids_sorted = which ? ws.ids_b : ws.ids_a;
      const int sblocks = (P + kScanBlock - 1) / kScanBlock;
      sorted_block_sums_kernel<<<sblocks, kScanThreads, 0, stream>>>(P, tiles_sorted, ws.block_sums);
      sorted_offsets_kernel<<<sblocks, kScanThreads, 0, stream>>>(P, tiles_sorted, ws.block_sums, ws.sorted_offsets, ws.counters,
                                                                 cap);
      nl += 2;
    }
    if ((rc = check_launch("depth sort / offsets", stream, dbg))) return rc;
    // tile-id sort buffers alias the 64-bit key arrays of the validation path
    uint32_t* tk_a = reinterpret_cast<uint32_t*>(ws.keys_in);
    uint32_t* tv_a = tk_a + cap;
    uint32_t* tk_b = reinterpret_cast<uint32_t*>(ws.keys_out);
    uint32_t* tv_b = tk_b + cap;
    const int tile_bits = (int)higher_msb((uint32_t)ntiles);
    {
      StageTimer tm(kStEmit, stream);
      radix_prepare(rs, (size_t)cap, tile_bits, stream);
      emit_sorted_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, ids_sorted, ws.sorted_offsets, ws.recA, ws.recB, ws.offsets,
                                                              pp.radii, ws.masks, ws.rects, gx, gy, a->flags, ws.counters, tk_a, tv_a,
                                                              radix_ghist(rs), (tile_bits + 7) / 8, radix_digit_bits(tile_bits));
      init_ranges_kernel<<<(unsigned)((ntiles + 255) / 256), 256, 0, stream>>>(ws.ranges, (uint32_t)ntiles);
      nl += 2;
    }
    if ((rc = check_launch("emit_sorted_kernel", stream, dbg))) return rc;
    {
      StageTimer tm(kStSort, stream);
      const int which = radix_sort_pairs(tk_a, tv_a, tk_b, tv_b, 0u, ws.counters, cap, (size_t)cap, tile_bits, rs, ws.ranges,
                                         stream, &nl, /*histogram_ready=*/true);
      point_list = which ? tv_b : tv_a;
    }
    count_launch(nl);
    if ((rc = check_launch("tile sort", stream, dbg))) return rc;
    if (a->num_rendered) {
      write_counts_kernel<<<1, 1, 0, stream>>>(nullptr, P, ws.counters, a->num_rendered);
      count_launch();
    }
  }
  {
    StageTimer tm(kStRender, stream);
    static const int render_impl = [] {
      // A/B switch for profiling: "block" = barrier-per-batch variant, "warp" = per-warp bit-scan variant,
      // "compact" = per-warp compacted hit list with a branch-free blend
      const char* e = getenv("GSB_RENDER_IMPL");
      if (e && e[0] == 'b') return 0;
      if (e && e[0] == 'c') return 2;
      if (e && e[0] == 'd') return 3;  // "dual": compact, two pixels per lane
      if (e && e[0] == 'w') return 1;
      return kDefaultRenderImpl;
    }();
    const int impl = ((a->flags >> 8) & 7u) ? (int)((a->flags >> 8) & 7u) - 1 : render_impl;
    const bool fast = (a->flags & GSB_RASTER_FAST_EXP) != 0;
#define GSB_LAUNCH_RENDER_T(THREADS, ...)                                                                                \
  __VA_ARGS__<<<dim3(gx, gy), THREADS, 0, stream>>>(ws.ranges, point_list, W, H, ws.recA, ws.recB, ws.recC, a->background, \
                                                    a->out_color, a->out_depth, a->out_final_T, cap)
#define GSB_LAUNCH_RENDER(KERNEL) GSB_LAUNCH_RENDER_T(kTilePixels, KERNEL)
    if (impl == 3) {
      if (fast)
        GSB_LAUNCH_RENDER_T(kTilePixels / 2, render_compact_kernel<true, 2>);
      else
        GSB_LAUNCH_RENDER_T(kTilePixels / 2, render_compact_kernel<false, 2>);
    } else if (impl == 2) {
      if (fast)
        GSB_LAUNCH_RENDER_T(kTilePixels, render_compact_kernel<true, 1>);
      else
        GSB_LAUNCH_RENDER_T(kTilePixels, render_compact_kernel<false, 1>);
    } else if (impl == 1) {
      if (fast)
        GSB_LAUNCH_RENDER(render_warp_kernel<true>);
      else
        GSB_LAUNCH_RENDER(render_warp_kernel<false>);
    } else {
      if (fast)
        GSB_LAUNCH_RENDER(render_kernel<true>);
      else
        GSB_LAUNCH_RENDER(render_kernel<false>);
    }
#undef GSB_LAUNCH_RENDER
#undef GSB_LAUNCH_RENDER_T
  }
  count_launch();
  if ((rc = check_launch("render_kernel", stream, dbg))) return rc;

  if (!use_cub && !(a->flags & GSB_RASTER_ASYNC)) {
    // synchronous contract: report an undersized workspace now (one wait at the END of the frame;
    // GSB_RASTER_ASYNC callers read num_rendered[2] themselves after their own synchronisation)
    unsigned long long host_counters[4] = {0, 0, 0, 0};
    GSB_CUDA_OK(cudaMemcpyAsync(host_counters, ws.counters, sizeof(host_counters), cudaMemcpyDeviceToHost, stream));
    GSB_CUDA_OK(cudaStreamSynchronize(stream));
    g_required_instances = (int64_t)host_counters[1];
    if (host_counters[2])
      return fail(GSB_ERR_WORKSPACE, "frame needs %lld instances, workspace sized for %lld", (long long)host_counters[1],
                  (long long)cap);
  }
  return GSB_OK;
}

int gsb_raster_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                            uint8_t* present, void* stream_v) {
  (void)projmatrix;  // the reference computes p_proj but only tests z_view (auxiliary.h:154)
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(GSB_ERR_INVALID, "mark_visible: bad arguments");
  if (P == 0) return GSB_OK;
  mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, viewmatrix, present);
  count_launch();
  return check_launch("mark_visible_kernel", stream, false);
}

int gsb_image_to_u8(const float* chw, int32_t width, int32_t height, uint8_t* hwc, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!chw || !hwc || width <= 0 || height <= 0) return fail(GSB_ERR_INVALID, "image_to_u8: bad arguments");
  const size_t n = (size_t)width * height;
  {
    StageTimer tm(kStToU8, stream);
    to_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(chw, width, height, hwc);
  }
  count_launch();
  return check_launch("to_u8_kernel", stream, false);
}

}  // extern "C"
