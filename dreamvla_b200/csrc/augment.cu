// Device-side pieces of the reference's collator: the random-shift augmentation and the depth resize, which
// the reference runs on CPU data-loader workers (utils/data_utils.py:326-383 `RandomShiftsAug`, :3588-3603 `depth_image_fn`,
// applied in `collator` :1337-1354).
//
// RandomShiftsAug: replicate-pad by `pad`, then grid_sample (bilinear, align_corners=False) on a grid of the padded image's
// own pixel centres moved by an INTEGER number of pixels (sx, sy) in [0, 2*pad] -- every sample point is a pixel centre, the
// bilinear weights are (1, 0, 0, 0), and the op is a shifted crop with clamped indices:
//     out[n, c, y, x] = in[n, c, clamp(y + sy - pad, 0, H-1), clamp(x + sx - pad, 0, W-1)]
// (exact; the reference's fp32 grid arithmetic lands within ~1e-5 pixels of the centres, oracle/augment_oracle.py measures the
// difference).  One pass: 4 B read (fp32 from the loader) + 2 B written (bf16 for the model) per element, HBM-bound.
//
// depth resize: torchvision `Resize(NEAREST)` on a tensor = F.interpolate(mode="nearest"): src = min(floor(dst * in/out), in-1),
// the scale computed in fp32 as torch does.
#include "common.cuh"
#include "../../include/dvla.h"

namespace dvla {
void set_error(const char* fmt, ...);
void count_launch();
int num_sms();

#define DVLA_CHECK_LAUNCH(name)                                                                 \
  do {                                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) { set_error("%s launch: %s", name, cudaGetErrorString(e__)); return DVLA_ERR_CUDA; } \
    count_launch();                                                                             \
  } while (0)

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as_float<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_from_float(T* p, float v);
template <> __device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_from_float<bf16>(bf16* p, float v) { *p = __float2bfloat16(v); }

// warp = one output row (img, channel, y) at a time, lanes stride x: the index arithmetic (two divisions, the clamp of y) is done
// once per row, reads of a row are contiguous (shifted by sx) except at the clamped borders, writes are contiguous
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256) shift_crop_kernel(const Tin* __restrict__ x, Tout* __restrict__ out,
                                                         const int* __restrict__ shifts, long long n, int c, int h, int w, int pad) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long rows = n * c * h;
  for (long long r = warp; r < rows; r += nwarps) {
    const long long img = r / (static_cast<long long>(c) * h);
    const int y = static_cast<int>(r % h);
    const int sx = __ldg(shifts + 2 * img), sy = __ldg(shifts + 2 * img + 1);
    const int ys = min(max(y + sy - pad, 0), h - 1);
    const Tin* src = x + (r - y + ys) * w;            // same image and channel, source row ys
    Tout* dst = out + r * w;
    for (int xx = lane; xx < w; xx += 32) st_from_float(dst + xx, ld_as_float(src + min(max(xx + sx - pad, 0), w - 1)));
  }
}

template <typename Tout>
__global__ void __launch_bounds__(256) resize_nearest_kernel(const float* __restrict__ x, Tout* __restrict__ out, long long n,
                                                             int hin, int win, int hout, int wout, float sh, float sw) {
  const long long plane = static_cast<long long>(hout) * wout;
  const long long total = n * plane;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long img = i / plane;
    const int yx = static_cast<int>(i - img * plane);
    const int y = yx / wout, xx = yx - y * wout;
    const int ys = min(static_cast<int>(floorf(y * sh)), hin - 1), xs = min(static_cast<int>(floorf(xx * sw)), win - 1);
    st_from_float(out + i, __ldg(x + img * hin * win + static_cast<long long>(ys) * win + xs));
  }
}

static unsigned grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = 32LL * num_sms();
  return static_cast<unsigned>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// dtype codes: 0 = fp32, 1 = bf16
int shift_crop_dispatch(const void* x, void* out, const int32_t* shifts, int64_t n, int64_t c, int64_t h, int64_t w, int32_t pad,
                        int32_t in_dtype, int32_t out_dtype, cudaStream_t s) {
  if (n == 0) return DVLA_OK;
  if (!x || !out || !shifts) { set_error("shift_crop: null pointer"); return DVLA_ERR_INVALID; }
  if (n < 0 || c <= 0 || h <= 0 || w <= 0 || pad < 0 || h > 32768 || w > 32768 || c > 65535) { set_error("shift_crop: bad dims"); return DVLA_ERR_INVALID; }
  if ((in_dtype | out_dtype) & ~1) { set_error("shift_crop: dtype codes are 0 (fp32) / 1 (bf16)"); return DVLA_ERR_INVALID; }
  const unsigned g = grid_for(n * c * h * 32);          // one warp per row
#define SC(TI, TO) shift_crop_kernel<TI, TO><<<g, 256, 0, s>>>((const TI*)x, (TO*)out, shifts, n, (int)c, (int)h, (int)w, pad)
  switch (in_dtype * 2 + out_dtype) {
    case 0: SC(float, float); break;
    case 1: SC(float, bf16); break;
    case 2: SC(bf16, float); break;
    default: SC(bf16, bf16); break;
  }
#undef SC
  DVLA_CHECK_LAUNCH("shift_crop");
  return DVLA_OK;
}

int resize_nearest_dispatch(const float* x, void* out, int64_t n, int64_t hin, int64_t win, int64_t hout, int64_t wout,
                            int32_t out_dtype, cudaStream_t s) {
  if (n == 0) return DVLA_OK;
  if (!x || !out) { set_error("resize_nearest: null pointer"); return DVLA_ERR_INVALID; }
  if (n < 0 || hin <= 0 || win <= 0 || hout <= 0 || wout <= 0 || hin > 32768 || win > 32768 || hout > 32768 || wout > 32768) {
    set_error("resize_nearest: bad dims"); return DVLA_ERR_INVALID;
  }
  if (out_dtype & ~1) { set_error("resize_nearest: dtype codes are 0 (fp32) / 1 (bf16)"); return DVLA_ERR_INVALID; }
  if (n == 0) return DVLA_OK;
  const float sh = static_cast<float>(hin) / static_cast<float>(hout), sw = static_cast<float>(win) / static_cast<float>(wout);
  const unsigned g = grid_for(n * hout * wout);
  if (out_dtype == 0) resize_nearest_kernel<float><<<g, 256, 0, s>>>(x, (float*)out, n, (int)hin, (int)win, (int)hout, (int)wout, sh, sw);
  else                resize_nearest_kernel<bf16><<<g, 256, 0, s>>>(x, (bf16*)out, n, (int)hin, (int)win, (int)hout, (int)wout, sh, sw);
  DVLA_CHECK_LAUNCH("resize_nearest");
  return DVLA_OK;
}

}  // namespace dvla
