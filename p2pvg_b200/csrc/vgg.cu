// Data-movement kernels of the vgg_64 backbone (reference models/vgg_64.py): MaxPool2d(2,2), nearest x2 upsampling,
// and the explicit 3x3 / stride-1 / pad-1 lowering used by the fp32 path and by the 3-channel ends of the network
// (the >= 64-channel layers run as implicit GEMMs in conv_gemm.cu, kinds 3-5).  All tensors NHWC; HBM-bound.
#include "common.cuh"

namespace {

inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 64;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// col[(n,y,x), tap*C + c] = x[n, y + sgn*(kh-1), x + sgn*(kw-1), c]  (zero outside the map); columns [9C, ld) = 0
template <typename T>
__global__ void im2col3_kernel(const T* __restrict__ x, T* __restrict__ col, int N, int H, int W, int C, int ld, int sgn) {
  const long long total = (long long)N * H * W * ld;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % ld);
    const long long pix = idx / ld;
    float v = 0.f;
    if (j < 9 * C) {
      const int tap = j / C, c = j - tap * C;
      const int kh = tap / 3, kw = tap - kh * 3;
      const int xx = (int)(pix % W);
      const int yy = (int)((pix / W) % H);
      const long long n = pix / ((long long)W * H);
      const int sy = yy + sgn * (kh - 1), sx = xx + sgn * (kw - 1);
      if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = ld_f(x + ((n * H + sy) * W + sx) * C + c);
    }
    st_f(col + idx, v);
  }
}

// y[(n,y,x), c] = bias[c] + sum_tap col[(n, y-(kh-1), x-(kw-1)), tap*C + c]     (ConvTranspose2d(k=3, s=1, p=1) scatter as a gather)
template <typename T>
__global__ void col2im3_kernel(const T* __restrict__ col, T* __restrict__ y, int N, int H, int W, int C, int ld, const float* __restrict__ bias) {
  const long long total = (long long)N * H * W * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long pix = idx / C;
    const int xx = (int)(pix % W);
    const int yy = (int)((pix / W) % H);
    const long long n = pix / ((long long)W * H);
    float acc = bias ? bias[c] : 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; kh++) {
      const int sy = yy - (kh - 1);
      if (sy < 0 || sy >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; kw++) {
        const int sx = xx - (kw - 1);
        if (sx < 0 || sx >= W) continue;
        acc += ld_f(col + ((n * H + sy) * W + sx) * ld + (kh * 3 + kw) * C + c);
      }
    }
    st_f(y + idx, acc);
  }
}

template <typename T>
__global__ void maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long p = idx / C;
    const int xo = (int)(p % Wo), yo = (int)((p / Wo) % Ho);
    const long long n = p / ((long long)Wo * Ho);
    const T* b = x + ((n * H + 2 * yo) * W + 2 * xo) * C + c;
    const float v = fmaxf(fmaxf(ld_f(b), ld_f(b + C)), fmaxf(ld_f(b + (long long)W * C), ld_f(b + (long long)W * C + C)));
    st_f(y + idx, v);
  }
}

// gradient goes to the first maximum of each window in row-major order (torch's max_pool2d tie rule)
template <typename T>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long p = idx / C;
    const int xo = (int)(p % Wo), yo = (int)((p / Wo) % Ho);
    const long long n = p / ((long long)Wo * Ho);
    const long long base = ((n * H + 2 * yo) * W + 2 * xo) * C + c;
    const long long off[4] = {0, (long long)C, (long long)W * C, (long long)W * C + C};
    int best = 0;
    float bv = ld_f(x + base);
#pragma unroll
    for (int k = 1; k < 4; k++) {
      const float v = ld_f(x + base + off[k]);
      if (v > bv) { bv = v; best = k; }
    }
    const float g = ld_f(dy + idx);
#pragma unroll
    for (int k = 0; k < 4; k++) st_f(dx + base + off[k], k == best ? g : 0.f);
  }
}

template <typename T>
__global__ void upsample2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
  const int Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)N * Ho * Wo * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long p = idx / C;
    const int xo = (int)(p % Wo), yo = (int)((p / Wo) % Ho);
    const long long n = p / ((long long)Wo * Ho);
    y[idx] = x[((n * H + (yo >> 1)) * W + (xo >> 1)) * C + c];
  }
}

template <typename T>
__global__ void upsample2_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C) {
  const long long total = (long long)N * H * W * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long p = idx / C;
    const int xi = (int)(p % W), yi = (int)((p / W) % H);
    const long long n = p / ((long long)W * H);
    const T* b = dy + ((n * 2 * H + 2 * yi) * (2 * W) + 2 * xi) * C + c;
    const long long row = 2LL * W * C;
    st_f(dx + idx, (ld_f(b) + ld_f(b + C)) + (ld_f(b + row) + ld_f(b + row + C)));
  }
}

// dst[g*n + i] += src[grp_src[g]*n + i]   (fp32 addend shared by the groups that reuse a skip frame)
template <typename T>
__global__ void gather_add_kernel(T* __restrict__ dst, const float* __restrict__ src, const int* __restrict__ grp_src, int G, long long n) {
  const long long total = (long long)G * n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long g = idx / n, i = idx - g * n;
    st_f(dst + idx, ld_f(dst + idx) + src[(long long)grp_src[g] * n + i]);
  }
}

}  // namespace

int p2pvg_im2col3_impl(const void* x, void* col, int dtype, int N, int H, int W, int C, int ld, int sgn, cudaStream_t st) {
  P2PVG_REQUIRE(ld >= 9 * C, P2PVG_ERR_BAD_ARG, "im2col3: ld %d < 9*C", ld);
  if (N == 0) return P2PVG_OK;
  const long long total = (long long)N * H * W * ld;
  DISPATCH_DTYPE(dtype, T, (im2col3_kernel<T><<<grid_for(total, 256), 256, 0, st>>>((const T*)x, (T*)col, N, H, W, C, ld, sgn < 0 ? -1 : 1)));
  return p2pvg_check_launch("im2col3");
}

int p2pvg_col2im3_impl(const void* col, void* y, int dtype, int N, int H, int W, int C, int ld, const float* bias, cudaStream_t st) {
  P2PVG_REQUIRE(ld >= 9 * C, P2PVG_ERR_BAD_ARG, "col2im3: ld %d < 9*C", ld);
  if (N == 0) return P2PVG_OK;
  const long long total = (long long)N * H * W * C;
  DISPATCH_DTYPE(dtype, T, (col2im3_kernel<T><<<grid_for(total, 256), 256, 0, st>>>((const T*)col, (T*)y, N, H, W, C, ld, bias)));
  return p2pvg_check_launch("col2im3");
}

int p2pvg_maxpool2_fwd_impl(const void* x, void* y, int dtype, int N, int H, int W, int C, cudaStream_t st) {
  P2PVG_REQUIRE(H % 2 == 0 && W % 2 == 0, P2PVG_ERR_BAD_ARG, "maxpool2: odd map %dx%d", H, W);
  if (N == 0) return P2PVG_OK;
  const long long total = (long long)N * (H / 2) * (W / 2) * C;
  DISPATCH_DTYPE(dtype, T, (maxpool2_fwd_kernel<T><<<grid_for(total, 256), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C)));
  return p2pvg_check_launch("maxpool2_fwd");
}

int p2pvg_maxpool2_bwd_impl(const void* x, const void* dy, void* dx, int dtype, int N, int H, int W, int C, cudaStream_t st) {
  P2PVG_REQUIRE(H % 2 == 0 && W % 2 == 0, P2PVG_ERR_BAD_ARG, "maxpool2: odd map %dx%d", H, W);
  if (N == 0) return P2PVG_OK;
  const long long total = (long long)N * (H / 2) * (W / 2) * C;
  DISPATCH_DTYPE(dtype, T, (maxpool2_bwd_kernel<T><<<grid_for(total, 256), 256, 0, st>>>((const T*)x, (const T*)dy, (T*)dx, N, H, W, C)));
  return p2pvg_check_launch("maxpool2_bwd");
}

int p2pvg_upsample2_fwd_impl(const void* x, void* y, int dtype, int N, int H, int W, int C, cudaStream_t st) {
  if (N == 0) return P2PVG_OK;
  const long long total = (long long)N * 4 * H * W * C;
  DISPATCH_DTYPE(dtype, T, (upsample2_fwd_kernel<T><<<grid_for(total, 256), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C)));
  return p2pvg_check_launch("upsample2_fwd");
}

int p2pvg_upsample2_bwd_impl(const void* dy, void* dx, int dtype, int N, int H, int W, int C, cudaStream_t st) {
  if (N == 0) return P2PVG_OK;
  const long long total = (long long)N * H * W * C;
  DISPATCH_DTYPE(dtype, T, (upsample2_bwd_kernel<T><<<grid_for(total, 256), 256, 0, st>>>((const T*)dy, (T*)dx, N, H, W, C)));
  return p2pvg_check_launch("upsample2_bwd");
}

int p2pvg_gather_add_impl(void* dst, int dtype, const float* src, const int* grp_src, int G, long long n, cudaStream_t st) {
  if (G == 0 || n == 0) return P2PVG_OK;
  DISPATCH_DTYPE(dtype, T, (gather_add_kernel<T><<<grid_for((long long)G * n, 256), 256, 0, st>>>((T*)dst, src, grp_src, G, n)));
  return p2pvg_check_launch("gather_add");
}
