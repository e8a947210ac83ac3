// Probe of cp.async.bulk.tensor.4d (tiled mode) semantics on sm_100a: out-of-bounds / negative coordinates and
// elementStrides (traversal stride) — used to design the implicit-GEMM convolution loaders.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, int rows, float* out, int* done) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(smem);
  for (int i = threadIdx.x; i < rows * 64; i += blockDim.x) tile[i] = __float2bfloat16(-7.f);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(rows * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     smem_u32(smem)),
                 "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
  }
  // bounded wait: report whether the expected byte count (rows*128) completed the barrier
  int ok = 0;
  if (threadIdx.x == 0) {
    for (int it = 0; it < 2000000 && !ok; it++) {
      uint32_t p;
      asm volatile("{\n .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.u32 %0, 1, 0, P; }\n" : "=r"(p) : "r"(smem_u32(&bar)) : "memory");
      ok = p;
    }
    *done = ok;
  }
  __syncthreads();
  // un-swizzle is irrelevant here: all 64 channels of a pixel hold the same value; read element 0 of each 128B row's first chunk
  for (int r = threadIdx.x; r < rows; r += blockDim.x) out[r] = __bfloat162float(tile[r * 64 + ((r & 7) ^ 0) * 0]);
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int N = 3, H = 6, W = 6, C = 64;
  std::vector<__nv_bfloat16> h(N * H * W * C);
  for (int n = 0; n < N; n++)
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++)
        for (int c = 0; c < C; c++) h[((n * H + y) * W + x) * C + c] = __float2bfloat16((float)(n * 36 + y * 6 + x));
  __nv_bfloat16* d;
  cudaMalloc(&d, h.size() * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  float* out;
  int* done;
  cudaMalloc(&out, 4096 * 4);
  cudaMalloc(&done, 4);
  struct T { int box[4]; int es[4]; int crd[4]; int rows; const char* what; };
  T tests[] = {
      {{64, 4, 4, 2}, {1, 1, 1, 1}, {0, -1, -1, 0}, 32, "stride1 box 4x4x2 at (-1,-1,n0): expect OOB zeros on first row/col"},
      {{64, 6, 2, 1}, {1, 1, 1, 1}, {0, 2, 5, 2}, 12, "stride1 box 6x2 at (x=2,y=5,n=2): OOB right/bottom"},
      {{64, 3, 3, 1}, {1, 2, 2, 1}, {0, 0, 0, 1}, 9, "elementStrides 2, box 3x3 at (0,0,n=1)"},
      {{64, 6, 6, 1}, {1, 2, 2, 1}, {0, 0, 0, 1}, 9, "elementStrides 2, box 6x6 (expect 3x3 loaded?) at (0,0,n=1), expect_tx for 9 rows"},
      {{64, 5, 5, 1}, {1, 2, 2, 1}, {0, -1, -1, 1}, 9, "elementStrides 2, box 5x5 at (-1,-1,n=1), expect_tx 9 rows"},
      {{64, 6, 6, 1}, {1, 2, 2, 1}, {0, 1, 1, 0}, 9, "elementStrides 2, box 6x6 at (1,1,n=0), expect_tx 9 rows"},
  };
  for (auto& t : tests) {
    CUtensorMap tm;
    cuuint64_t dims[4] = {C, W, H, N};
    cuuint64_t strides[3] = {C * 2, W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)t.box[0], (cuuint32_t)t.box[1], (cuuint32_t)t.box[2], (cuuint32_t)t.box[3]};
    cuuint32_t es[4] = {(cuuint32_t)t.es[0], (cuuint32_t)t.es[1], (cuuint32_t)t.es[2], (cuuint32_t)t.es[3]};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("\n== %s\n   encode rc=%d box=(%d,%d,%d,%d) estr=(%d,%d,%d,%d) crd=(%d,%d,%d,%d)\n", t.what, (int)r, t.box[0], t.box[1], t.box[2],
           t.box[3], t.es[0], t.es[1], t.es[2], t.es[3], t.crd[0], t.crd[1], t.crd[2], t.crd[3]);
    if (r != CUDA_SUCCESS) continue;
    const int dump = 48;
    cudaMemset(out, 0, 4096 * 4);
    probe<<<1, 128, 16 * 1024>>>(tm, t.crd[0], t.crd[1], t.crd[2], t.crd[3], t.rows, out, done);
    cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
    int hd = 0;
    std::vector<float> ho(dump);
    cudaMemcpy(&hd, done, 4, cudaMemcpyDeviceToHost);
    // re-run dumping more rows than expected to see how far TMA wrote
    probe<<<1, 128, 16 * 1024>>>(tm, t.crd[0], t.crd[1], t.crd[2], t.crd[3], dump, out, done);
    cudaDeviceSynchronize();
    cudaMemcpy(ho.data(), out, dump * 4, cudaMemcpyDeviceToHost);
    printf("   sync=%s barrier_completed_with_%d_rows=%d\n   rows:", cudaGetErrorString(e), t.rows, hd);
    for (int i = 0; i < dump; i++) printf(" %g", ho[i]);
    printf("\n");
  }
  return 0;
}
