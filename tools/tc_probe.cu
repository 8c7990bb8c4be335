// Hardware probe for the tcgen05 building blocks used by the tensor-core actor path:
//   test 0: D[128x128] = A[128x64] * B[128x64]^T, fp16 operands in smem (K-major, no swizzle), fp32 accum
//   test 1: same with A read from TMEM (tcgen05.st -> tcgen05.mma "ts" form)
//   test 2: fp32 inputs split into fp16 hi/lo, 3 MMAs per k-block (hi*hi + lo*hi + hi*lo) -> fp32-level accuracy
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_probe tools/tc_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../multiagent_rl_b200/csrc/tc_common.cuh"

using namespace mpe;

constexpr int M = 128, N = 128, K = 64;

// smem operand layout: [K/8][rows][8 halves]  (core matrix = 8 rows x 16 B contiguous; LBO = rows*16, SBO = 128)
__device__ __host__ inline int op_index(int r, int k, int rows) { return (k / 8) * rows * 8 + r * 8 + (k % 8); }

__global__ void __launch_bounds__(128) probe(const float *A, const float *B, float *D, int test) {
  extern __shared__ __align__(128) unsigned char smem[];
  __half *sAh = reinterpret_cast<__half *>(smem);   // [K/8][128][8]
  __half *sAl = sAh + M * K;
  __half *sBh = sAl + M * K;
  __half *sBl = sBh + N * K;
  uint64_t *bar = reinterpret_cast<uint64_t *>(sBl + N * K);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  // operands: thread = row
  for (int k = 0; k < K; ++k) {
    const float a = A[tid * K + k], b = B[tid * K + k];
    const __half ah = __float2half_rn(a), bh = __float2half_rn(b);
    sAh[op_index(tid, k, M)] = ah; sAl[op_index(tid, k, M)] = __float2half_rn(a - __half2float(ah));
    sBh[op_index(tid, k, N)] = bh; sBl[op_index(tid, k, N)] = __float2half_rn(b - __half2float(bh));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t col_d = 0, col_ah = 128, col_al = 128 + K / 2;  // A in TMEM: 2 halves per 32-bit column

  if (test == 1) {  // stage A (hi and lo) into TMEM: lane = row, column c holds k = 2c, 2c+1
    uint32_t vh[K / 2], vl[K / 2];
    for (int c = 0; c < K / 2; ++c) {
      const __half2 h = __halves2half2(sAh[op_index(tid, 2 * c, M)], sAh[op_index(tid, 2 * c + 1, M)]);
      const __half2 l = __halves2half2(sAl[op_index(tid, 2 * c, M)], sAl[op_index(tid, 2 * c + 1, M)]);
      vh[c] = *reinterpret_cast<const uint32_t *>(&h); vl[c] = *reinterpret_cast<const uint32_t *>(&l);
    }
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + col_ah, vh);
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + col_al, vl);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(M, N);
    for (int kb = 0; kb < K / 16; ++kb) {
      const uint64_t bh = make_smem_desc(sBh + kb * 2 * N * 8, N * 16, 128);
      const uint64_t bl = make_smem_desc(sBl + kb * 2 * N * 8, N * 16, 128);
      if (test == 1) {
        mma_f16_ts(tmem + col_d, tmem + col_ah + kb * 8, bh, idesc, kb > 0);
        mma_f16_ts(tmem + col_d, tmem + col_al + kb * 8, bh, idesc, true);
        mma_f16_ts(tmem + col_d, tmem + col_ah + kb * 8, bl, idesc, true);
      } else {
        const uint64_t ah = make_smem_desc(sAh + kb * 2 * M * 8, M * 16, 128);
        const uint64_t al = make_smem_desc(sAl + kb * 2 * M * 8, M * 16, 128);
        mma_f16_ss(tmem + col_d, ah, bh, idesc, kb > 0);
        if (test == 2) {
          mma_f16_ss(tmem + col_d, al, bh, idesc, true);
          mma_f16_ss(tmem + col_d, ah, bl, idesc, true);
        }
      }
    }
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + col_d + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tmem, 512);
}

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1);
  for (auto &x : A) x = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  for (auto &x : B) x = ((rand() / (float)RAND_MAX) * 2.f - 1.f) * 0.18f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = (2 * M * K + 2 * N * K) * 2 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int bad = 0;
  for (int test = 0; test < 3; ++test) {
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 128, smem>>>(dA, dB, dD, test);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("test %d: CUDA error %s\n", test, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double max_err64 = 0, max_err16 = 0;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        double ref = 0, ref16 = 0;
        for (int k = 0; k < K; ++k) {
          ref += (double)A[i * K + k] * (double)B[j * K + k];
          ref16 += (double)__half2float(__float2half_rn(A[i * K + k])) * (double)__half2float(__float2half_rn(B[j * K + k]));
        }
        max_err64 = fmax(max_err64, fabs(D[i * N + j] - ref));
        max_err16 = fmax(max_err16, fabs(D[i * N + j] - ref16));
      }
    // tests 0: fp16-rounded inputs -> matches ref16 to fp32 accumulation error; tests 1, 2: matches the exact product
    const bool ok = (test == 0) ? max_err16 < 2e-6 : max_err64 < 2e-6;
    printf("test %d: max|D - exact| = %.3e   max|D - fp16-input product| = %.3e   %s\n", test, max_err64, max_err16,
           ok ? "OK" : "FAIL");
    bad += !ok;
  }
  return bad;
}
