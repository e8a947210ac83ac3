#include "common.cuh"
int p2pvg_gemm_tc_available() { return 0; }
int p2pvg_gemm_tc(const void*, int, long long, const void*, int, long long, void*, int, long long, int, int, int, int, const float*,
                  const void*, long long, void*, size_t, cudaStream_t) {
  p2pvg_set_error("gemm_tc: stub");
  return P2PVG_ERR_UNSUPPORTED;
}
