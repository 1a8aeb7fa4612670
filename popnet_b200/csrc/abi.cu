// Library-level entry points (version, diagnostics) and the eval-time depth preprocessing kernel.
#include "common.cuh"

namespace popnet {
std::atomic<int> g_last_cuda_error{0};
std::atomic<long long> g_launch_count{0};
}  // namespace popnet

extern "C" int popnet_abi_version(void) { return POPNET_ABI_VERSION; }
extern "C" int popnet_last_cuda_error(void) { return popnet::g_last_cuda_error.load(); }
extern "C" long long popnet_launch_count(void) { return popnet::g_launch_count.load(); }

namespace {

// OpenCV INTER_LINEAR (half-pixel centres, replicate border) + clamp + normalise, one thread per
// output pixel; reads are the 2x2 neighbourhood, writes are coalesced.  Replaces
// data_augmentation_2d3d.py:497-522 (Resize) and datasets_kdh3d_rtpose_mpreal.py:CR229-246.
__global__ void __launch_bounds__(256) preprocess_kernel(const float* __restrict__ src, int src_h, int src_w,
                                                         float* __restrict__ dst, int dst_h, int dst_w,
                                                         float depth_max, float depth_mean, float depth_std) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x >= dst_w) return;
  const double sx_scale = (double)src_w / dst_w, sy_scale = (double)src_h / dst_h;
  float fx = (float)((x + 0.5) * sx_scale - 0.5);
  int ix = (int)floorf(fx);
  fx -= ix;
  if (ix < 0) { ix = 0; fx = 0.f; }
  if (ix >= src_w - 1) { ix = src_w - 1; fx = 0.f; }
  float fy = (float)((y + 0.5) * sy_scale - 0.5);
  int iy = (int)floorf(fy);
  fy -= iy;
  if (iy < 0) { iy = 0; fy = 0.f; }
  if (iy >= src_h - 1) { iy = src_h - 1; fy = 0.f; }
  const int ix1 = min(ix + 1, src_w - 1), iy1 = min(iy + 1, src_h - 1);
  const float* s = src + (size_t)b * src_h * src_w;
  const float r0 = __fadd_rn(__fmul_rn(s[iy * src_w + ix], 1.f - fx), __fmul_rn(s[iy * src_w + ix1], fx));
  const float r1 = __fadd_rn(__fmul_rn(s[iy1 * src_w + ix], 1.f - fx), __fmul_rn(s[iy1 * src_w + ix1], fx));
  float v = __fadd_rn(__fmul_rn(r0, 1.f - fy), __fmul_rn(r1, fy));
  v = fminf(fmaxf(v, 0.f), depth_max);
  dst[((size_t)b * dst_h + y) * dst_w + x] = __fdiv_rn(__fsub_rn(v, depth_mean), depth_std);
}

}  // namespace

extern "C" int popnet_preprocess_depth(const float* src, int batch, int src_h, int src_w, float* dst, int dst_h,
                                       int dst_w, float depth_max, float depth_mean, float depth_std, void* stream) {
  if (!src || !dst || batch < 0 || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1 || dst_h > 65535 || batch > 65535)
    return POPNET_ERR_INVALID_ARG;
  if (batch == 0) return POPNET_OK;
  dim3 grid((dst_w + 255) / 256, dst_h, batch);
  preprocess_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, src_h, src_w, dst, dst_h, dst_w,
                                                                        depth_max, depth_mean, depth_std);
  POPNET_AFTER_LAUNCH();
  return POPNET_OK;
}
