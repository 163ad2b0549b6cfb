// Micro-benchmark: which issue pipe do the packed 16-bit integer and fp16x2 instructions use on sm_100a?
// Each kernel runs N independent chains of one (or two interleaved) instruction kinds; the result is warp
// instructions per clock per SM (4 sub-partitions). Two kinds on different pipes add up; on the same pipe they do not.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

#define CHAINS 8
#define ITERS 2048

__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

template <int KIND>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b)
{
  if (KIND == 0) return __vminu2(a, b);                                  // VIMNMX.U16x2
  if (KIND == 1) return h2u(__hmin2(u2h(a), u2h(b)));                    // HMNMX2
  if (KIND == 2) return h2u(__hgt2(u2h(a), u2h(b)));                     // HSET2.BF.GT
  if (KIND == 3) return h2u(__hfma2(u2h(a), u2h(b), u2h(b)));            // HFMA2
  if (KIND == 4) return (a & b) | (~a & 0x12345678U);                    // LOP3
  if (KIND == 5) return a * b + 0x00010000U;                             // IMAD
  if (KIND == 6) return __viaddmax_u16x2(a, b, 0x00010001U);             // VIADDMNMX
  if (KIND == 7) return h2u(__hadd2(u2h(a), u2h(b)));                    // HADD2
  if (KIND == 8) return __byte_perm(a, b, 0x6420);                       // PRMT
  if (KIND == 9) return h2u(__hmax2(__habs2(u2h(a)), u2h(b)));           // HMNMX2 with |a|
  if (KIND == 10) return 0x00010000U - a + (b & 1);                      // IADD3
  return a;
}

template <int K1, int K2>
__global__ void bench(uint32_t* out, long long* cycles, uint32_t seed)
{
  uint32_t x[CHAINS], y[CHAINS];
#pragma unroll
  for (int c = 0; c != CHAINS; ++c) {
    x[c] = seed * (c + 1) + threadIdx.x;
    y[c] = seed * (c + 7) + threadIdx.x * 3;
  }
  uint32_t  b  = seed | 0x3c003c00U;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i != ITERS; ++i) {
#pragma unroll
    for (int c = 0; c != CHAINS; ++c) {
      x[c] = op<K1>(x[c], b);
      if (K2 >= 0) y[c] = op<K2>(y[c], b);
    }
  }
  long long t1 = clock64();
  uint32_t  r  = 0;
#pragma unroll
  for (int c = 0; c != CHAINS; ++c) r ^= x[c] ^ y[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int K1, int K2>
void run(const char* name, uint32_t* out, long long* cyc)
{
  const int blocks = 148, threads = 1024;
  bench<K1, K2><<<blocks, threads>>>(out, cyc, 12345);
  bench<K1, K2><<<blocks, threads>>>(out, cyc, 12345);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i != blocks; ++i) mean += (double)h[i];
  mean /= blocks;
  double instr = (double)(threads / 32) * ITERS * CHAINS * (K2 >= 0 ? 2 : 1);
  printf("%-28s %.3f warp-instr/clk/SM (%.3f per sub-partition)\n", name, instr / mean, instr / mean / 4);
}

int main()
{
  uint32_t*  out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<0, -1>("VIMNMX.U16x2", out, cyc);
  run<1, -1>("HMNMX2", out, cyc);
  run<9, -1>("HMNMX2 |a|", out, cyc);
  run<2, -1>("HSET2.GT", out, cyc);
  run<3, -1>("HFMA2", out, cyc);
  run<7, -1>("HADD2", out, cyc);
  run<4, -1>("LOP3", out, cyc);
  run<5, -1>("IMAD", out, cyc);
  run<6, -1>("VIADDMNMX.U16x2", out, cyc);
  run<8, -1>("PRMT", out, cyc);
  run<10, -1>("IADD3", out, cyc);
  run<0, 3>("VIMNMX + HFMA2", out, cyc);
  run<0, 1>("VIMNMX + HMNMX2", out, cyc);
  run<1, 3>("HMNMX2 + HFMA2", out, cyc);
  run<2, 3>("HSET2 + HFMA2", out, cyc);
  run<2, 0>("HSET2 + VIMNMX", out, cyc);
  run<4, 3>("LOP3 + HFMA2", out, cyc);
  run<4, 5>("LOP3 + IMAD", out, cyc);
  run<5, 3>("IMAD + HFMA2", out, cyc);
  run<4, 0>("LOP3 + VIMNMX", out, cyc);
  run<6, 3>("VIADDMNMX + HFMA2", out, cyc);
  run<10, 5>("IADD3 + IMAD", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
