// Fused sigmoid + MSE (forward value and gradient w.r.t. the pre-sigmoid decoder output in one
// pass), the loss finaliser that produces the four host-visible scalars of P2PModel.forward
// (models/p2p_model.py:271), and the flat-arena Adam update with PyTorch-1.0 arithmetic.
#include "common.cuh"

#define MSE_CHUNKS 32

namespace {

template <typename T>
__global__ void __launch_bounds__(256) sigmoid_mse_kernel(const T* __restrict__ raw, const float* __restrict__ x,
                                                          const int* __restrict__ tgt, const float* __restrict__ coef,
                                                          long long E, T* __restrict__ pred, T* __restrict__ d_raw,
                                                          float* __restrict__ partial) {
  const int g = blockIdx.y;
  const float* xt = x + (long long)tgt[g] * E;
  const T* rg = raw + (long long)g * E;
  const float cf = coef[g];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (long long)gridDim.x * blockDim.x) {
    float s = sigmoidf_(ld_f<T>(rg + i));
    float d = s - xt[i];
    acc += (double)d * (double)d;
    if (pred) st_f<T>(pred + (long long)g * E + i, s);
    if (d_raw) st_f<T>(d_raw + (long long)g * E + i, cf * 2.f * d * s * (1.f - s));
  }
  __shared__ double sh[8];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < 8; w++) v += sh[w];
    partial[(long long)g * gridDim.x + blockIdx.x] = (float)v;
  }
}

__global__ void finalize_losses_kernel(const float* __restrict__ mse_partial, int n_recon, int has_cpc, int nchunk, double E,
                                       const float* __restrict__ kl_sum, float batch_size, const float* __restrict__ align_partial,
                                       int n_align, float seq_len, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double mse = 0.0, cpc = 0.0, al = 0.0;
  for (int g = 0; g < n_recon; g++) {
    double s = 0.0;
    for (int k = 0; k < nchunk; k++) s += (double)mse_partial[g * nchunk + k];
    mse += s / E;
  }
  if (has_cpc) {
    double s = 0.0;
    for (int k = 0; k < nchunk; k++) s += (double)mse_partial[n_recon * nchunk + k];
    cpc = s / E;
  }
  for (int s = 0; s < n_align; s++) al += (double)align_partial[s];
  out[0] = (float)(mse / seq_len);
  out[1] = (float)((double)kl_sum[0] / batch_size / seq_len);
  out[2] = (float)(cpc / seq_len);
  out[3] = (float)(al / seq_len);
}

// torch-1.0 Adam: denom = sqrt(v) + eps; p -= lr*sqrt(bc2)/bc1 * m/denom
__global__ void adam_legacy_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                   long long n, double beta1_d, double beta2_d, double eps_d, double lr,
                                   const int* __restrict__ step_ptr) {
  __shared__ float s_step_size;
  if (threadIdx.x == 0) {
    // step counter lives in device memory so that a captured CUDA graph sees the live value
    const double t = (double)step_ptr[0];
    const double bc1 = 1.0 - pow(beta1_d, t), bc2 = 1.0 - pow(beta2_d, t);
    s_step_size = (float)(lr * sqrt(bc2) / bc1);
  }
  __syncthreads();
  const float step_size = s_step_size;
  // scalars are rounded to fp32 exactly as torch rounds python doubles applied to fp32 tensors
  const float beta1 = (float)beta1_d, beta2 = (float)beta2_d, eps = (float)eps_d;
  const float omb1 = (float)(1.0 - beta1_d), omb2 = (float)(1.0 - beta2_d);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i];
    float mi = m[i] * beta1 + omb1 * gi;
    float vi = v[i] * beta2 + omb2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - step_size * (mi / (sqrtf(vi) + eps));
  }
}

__global__ void scale_kernel(float* __restrict__ x, long long n, float a) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= a;
}

}  // namespace

int p2pvg_mse_chunks_impl() { return MSE_CHUNKS; }

int p2pvg_sigmoid_mse_impl(const void* raw, int dtype, const float* x, const int* tgt, const float* coef, int G, long long E,
                           void* pred, void* d_raw, float* partial, cudaStream_t st) {
  if (G == 0 || E == 0) return P2PVG_OK;
  dim3 grid(MSE_CHUNKS, G);
  DISPATCH_DTYPE(dtype, T, (sigmoid_mse_kernel<T><<<grid, 256, 0, st>>>((const T*)raw, x, tgt, coef, E, (T*)pred, (T*)d_raw, partial)));
  return p2pvg_check_launch("sigmoid_mse");
}

int p2pvg_finalize_losses_impl(const float* mse_partial, int n_recon, int has_cpc, double E, const float* kl_sum, float batch_size,
                               const float* align_partial, int n_align, float seq_len, float* out, cudaStream_t st) {
  finalize_losses_kernel<<<1, 32, 0, st>>>(mse_partial, n_recon, has_cpc, MSE_CHUNKS, E, kl_sum, batch_size, align_partial, n_align,
                                           seq_len, out);
  return p2pvg_check_launch("finalize_losses");
}

int p2pvg_adam_legacy_impl(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2, double eps,
                           const int* step_ptr, cudaStream_t st) {
  if (n == 0) return P2PVG_OK;
  P2PVG_REQUIRE(step_ptr != nullptr, P2PVG_ERR_BAD_ARG, "adam: step counter pointer is null");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adam_legacy_kernel<<<(int)blocks, 256, 0, st>>>(p, g, m, v, n, beta1, beta2, eps, lr, step_ptr);
  return p2pvg_check_launch("adam_legacy");
}

int p2pvg_scale_impl(float* x, long long n, float a, cudaStream_t st) {
  if (n == 0) return P2PVG_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  scale_kernel<<<(int)blocks, 256, 0, st>>>(x, n, a);
  return p2pvg_check_launch("scale");
}
