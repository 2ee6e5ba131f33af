// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin C-ABI shim around the UNMODIFIED reference rasterizer
// (third_party/gaussian-splatting/submodules/diff-gaussian-rasterization,
//  cuda_rasterizer/rasterizer.h:20-53 `CudaRasterizer::Rasterizer::forward`).
// The reference sources are compiled where they lie under /root/reference by
// oracle/build_ref.sh; nothing from them is copied into this repository.  The
// resulting oracle/_ref/libref_dgr.so is used by tests/ and by
// `bench.py --impl reference` as "the reference itself, run on the B200".
//
// It replaces (for test purposes) the torch marshalling in
// rasterize_points.cu:35-115: scratch buffers are grow-only cudaMalloc blocks
// instead of torch byte tensors resized through std::function callbacks.
#include <cstdint>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <cuda_runtime.h>
#include "cuda_rasterizer/rasterizer_impl.h"

namespace {
struct Scratch {
  char* ptr = nullptr;
  size_t cap = 0;
  char* grow(size_t n) {
    if (n > cap) {
      if (ptr) cudaFree(ptr);
      size_t want = n + n / 4 + 256;
      if (cudaMalloc(&ptr, want) != cudaSuccess) throw std::runtime_error("ref_dgr: cudaMalloc failed");
      cap = want;
    }
    return ptr;
  }
};
Scratch g_geom, g_bin, g_img;
char g_err[512] = {0};
}  // namespace

extern "C" {

const char* ref_dgr_last_error() { return g_err; }

// Returns num_rendered (>= 0) or -1 on error.  All pointers are device pointers
// (nullable where the reference allows: shs / colors_precomp, scales+rotations /
// cov3D_precomp).  Kernels run on the legacy default stream like the reference.
int ref_dgr_forward(int P, int D, int M, const float* background, int W, int H,
                    const float* means3D, const float* shs, const float* colors_precomp,
                    const float* opacities, const float* scales, float scale_modifier,
                    const float* rotations, const float* cov3D_precomp,
                    const float* viewmatrix, const float* projmatrix, const float* campos,
                    float tan_fovx, float tan_fovy, int prefiltered,
                    float* out_color, int* radii, int debug) {
  try {
    std::function<char*(size_t)> gf = [](size_t n) { return g_geom.grow(n); };
    std::function<char*(size_t)> bf = [](size_t n) { return g_bin.grow(n); };
    std::function<char*(size_t)> imf = [](size_t n) { return g_img.grow(n); };
    int r = CudaRasterizer::Rasterizer::forward(
        gf, bf, imf, P, D, M, background, W, H, means3D, shs, colors_precomp, opacities,
        scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
        tan_fovx, tan_fovy, prefiltered != 0, out_color, radii, debug != 0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "ref_dgr: %s", cudaGetErrorString(e));
      return -1;
    }
    return r;
  } catch (const std::exception& ex) {
    snprintf(g_err, sizeof(g_err), "ref_dgr: %s", ex.what());
    return -1;
  }
}

// Copies the reference's per-pixel final transmittance (ImageState.accum_alpha,
// rasterizer_impl.h:47) of the LAST forward call into `out` (device, W*H floats).
int ref_dgr_last_final_T(int W, int H, float* out) {
  if (!g_img.ptr) return -1;
  char* chunk = g_img.ptr;
  CudaRasterizer::ImageState img = CudaRasterizer::ImageState::fromChunk(chunk, (size_t)W * H);
  return cudaMemcpy(out, img.accum_alpha, sizeof(float) * (size_t)W * H, cudaMemcpyDeviceToDevice) == cudaSuccess ? 0 : -1;
}

// Copies the reference's per-Gaussian intermediates of the LAST forward call
// (GeometryState, rasterizer_impl.h:29-43).  Any output pointer may be NULL.
int ref_dgr_last_geometry(int P, float* depths, float* means2D, float* conic_opacity,
                          float* rgb, unsigned* tiles_touched) {
  if (!g_geom.ptr) return -1;
  char* chunk = g_geom.ptr;
  CudaRasterizer::GeometryState g = CudaRasterizer::GeometryState::fromChunk(chunk, (size_t)P);
  cudaError_t e = cudaSuccess;
  if (depths && e == cudaSuccess) e = cudaMemcpy(depths, g.depths, 4 * (size_t)P, cudaMemcpyDeviceToDevice);
  if (means2D && e == cudaSuccess) e = cudaMemcpy(means2D, g.means2D, 8 * (size_t)P, cudaMemcpyDeviceToDevice);
  if (conic_opacity && e == cudaSuccess) e = cudaMemcpy(conic_opacity, g.conic_opacity, 16 * (size_t)P, cudaMemcpyDeviceToDevice);
  if (rgb && e == cudaSuccess) e = cudaMemcpy(rgb, g.rgb, 12 * (size_t)P, cudaMemcpyDeviceToDevice);
  if (tiles_touched && e == cudaSuccess) e = cudaMemcpy(tiles_touched, g.tiles_touched, 4 * (size_t)P, cudaMemcpyDeviceToDevice);
  return e == cudaSuccess ? 0 : -1;
}

void ref_dgr_release() {
  for (Scratch* s : {&g_geom, &g_bin, &g_img}) {
    if (s->ptr) cudaFree(s->ptr);
    s->ptr = nullptr;
    s->cap = 0;
  }
}
}
