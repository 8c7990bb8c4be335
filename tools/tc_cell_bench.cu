// Microbenchmark of the LSTM cell math (lstm_chunk) and dense1 epilogue in isolation: cycles per 8-unit chunk as a
// function of the number of warps per SM sub-partition.  Answers "how far is the cell phase of k_tc2 from what the
// MUFU / issue pipes can do, and how much thread-level parallelism does it need?".
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc_cell_bench tools/tc_cell_bench.cu
#include "../multiagent_rl_b200/csrc/tc_kernels.cu"

#include <cstdio>
using namespace mpe;

template <int MODE>
__global__ void cell_bench(const float *bias, const float *w2, float *out, long long *cycles, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  float *s_bg = reinterpret_cast<float *>(smem);          // 128
  float *s_w2 = s_bg + 128;                                // 64*16
  unsigned char *s_h = reinterpret_cast<unsigned char *>(s_w2 + 64 * 16);  // 16 KB
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_bg[i] = bias[i];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_w2[i] = w2[i];
  __syncthreads();
  const int row = threadIdx.x & 127;
  uint32_t v[32];
  for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(0.37f * (float)((threadIdx.x * 7 + i * 13) % 29 - 14));
  f2 c[4] = {0ull, 0ull, 0ull, 0ull};
  f2 pl[4] = {0ull, 0ull, 0ull, 0ull};
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      lstm_chunk<8>(v, s_bg + (it & 3) * 32, s_w2 + (it & 7) * 128, c, pl, s_h + (it & 3) * kChunkA, row, true);
    } else {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float a0 = fmaxf(fmaf(__uint_as_float(v[2 * j]), kWInv, s_bg[2 * j]), 0.0f);
        const float a1 = fmaxf(fmaf(__uint_as_float(v[2 * j + 1]), kWInv, s_bg[2 * j + 1]), 0.0f);
        __half h0, l0, h1, l1;
        split_f16(a0, h0, l0); split_f16(a1, h1, l1);
        hi[j] = pack_h2(h0, h1); lo[j] = pack_h2(l0, l1);
      }
      *reinterpret_cast<uint4 *>(s_h + row * 16) = make_uint4(hi[0] ^ hi[5], hi[1] ^ hi[6], lo[2] ^ lo[7], lo[3] ^ hi[15]);
    }
    // feed something back so iterations depend on each other only through c (like the real recurrence)
    float x0, x1; upk(c[it & 3], x0, x1);
    v[(it * 5) & 31] = __float_as_uint(x0 * 3.0f + x1);
  }
  const long long t1 = clock64();
  float x0, x1; upk(pl[0], x0, x1);
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  float *bias, *w2, *out; long long *cyc;
  cudaMalloc(&bias, 512); cudaMalloc(&w2, 4096); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  cudaMemset(bias, 0, 512); cudaMemset(w2, 0, 4096);
  const int iters = 400;
  const int smem = 512 + 4096 + 16384;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps_per_smsp = 1; warps_per_smsp <= 4; ++warps_per_smsp) {
      const int threads = warps_per_smsp * 128;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) cell_bench<0><<<148, threads, smem>>>(bias, w2, out, cyc, iters);
        else cell_bench<1><<<148, threads, smem>>>(bias, w2, out, cyc, iters);
        cudaDeviceSynchronize();
      }
      long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      printf("%s warps/SMSP=%d: %.0f cycles per call per warp (%.0f per SMSP-warp-slot); MUFU floor %d, per-SMSP rate %.0f cycles per chunk\n",
             mode == 0 ? "lstm_chunk (8 units)" : "dense1_half (32 cols)", warps_per_smsp, avg / iters, avg / iters,
             mode == 0 ? 64 * 8 : 0, avg / iters / warps_per_smsp);
    }
  return 0;
}
