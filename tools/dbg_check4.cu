// Development tool: runs pk::check4<DEG> on the GPU and on the CPU on the same random inputs and reports mismatches.
#include "../srsran_projectvtlmo_b200/csrc/ldpc_packed_math.h"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace pusch_dec::pk;

template <int DEG>
__host__ __device__ void run_one(const uint32_t* s_in, const uint32_t* c_in, uint32_t* s_out, uint32_t* c_out, uint32_t mult)
{
  check4<DEG> ck;
  ck.begin();
#pragma unroll
  for (int e = 0; e != DEG; ++e) {
    ck.gather(e, s_in[2 * e], s_in[2 * e + 1], c_in[e]);
  }
  ck.reduce(mult);
#pragma unroll
  for (int e = 0; e != DEG; ++e) {
    uint32_t s0, s1;
    c_out[e]         = ck.scatter(e, s0, s1);
    s_out[2 * e]     = s0;
    s_out[2 * e + 1] = s1;
  }
}

template <int DEG>
__global__ void kern(const uint32_t* s_in, const uint32_t* c_in, uint32_t* s_out, uint32_t* c_out, uint32_t mult, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    run_one<DEG>(s_in + (size_t)i * 2 * DEG, c_in + (size_t)i * DEG, s_out + (size_t)i * 2 * DEG, c_out + (size_t)i * DEG, mult);
  }
}

/// One soft value in the decoder's storage format: the binary16 number S + 1152 (S in [-120, 120]), or an "infinite" value:
/// the loader's +-infinity, a grown one, or an IEEE infinity.
static uint32_t soft_lane(int mode)
{
  int r = rand() % 100;
  if (r < 10) {
    return H_1152; // S = 0
  }
  if (r < 15) {
    static const uint32_t inf[6] = {H_POS_INF, H_NEG_INF, 0x7c00U, 0xfc00U, 0x7a00U, 0xf9f0U};
    return inf[rand() % 6];
  }
  int v = (mode == 0) ? rand() % 241 - 120 : rand() % 21 - 10;
  return (H_1024 | (uint32_t)(v + 128)) & 0xffffU;
}

template <int DEG>
int test(int n, uint32_t mult)
{
  std::vector<uint32_t> s(n * 2 * DEG), c(n * DEG), so(n * 2 * DEG), co(n * DEG), sh(n * 2 * DEG), ch(n * DEG);
  for (int i = 0; i != n; ++i) {
    for (int e = 0; e != DEG; ++e) {
      for (int r = 0; r != 2; ++r) {
        s[(i * DEG + e) * 2 + r] = soft_lane(i & 1) | (soft_lane(i & 1) << 16);
      }
      uint32_t w = 0;
      for (int b = 0; b != 4; ++b) {
        int v = (i % 3 == 0) ? 0 : rand() % 191 - 95;
        w |= (uint32_t)(v + 128) << (8 * b);
      }
      c[i * DEG + e] = w;
    }
  }
  uint32_t *ds, *dc, *dso, *dco;
  cudaMalloc(&ds, s.size() * 4);
  cudaMalloc(&dc, c.size() * 4);
  cudaMalloc(&dso, s.size() * 4);
  cudaMalloc(&dco, c.size() * 4);
  cudaMemcpy(ds, s.data(), s.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dc, c.data(), c.size() * 4, cudaMemcpyHostToDevice);
  kern<DEG><<<(n + 127) / 128, 128>>>(ds, dc, dso, dco, mult, n);
  cudaMemcpy(so.data(), dso, s.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(co.data(), dco, c.size() * 4, cudaMemcpyDeviceToHost);
  cudaError_t err = cudaDeviceSynchronize();
  int bad = 0;
  for (int i = 0; i != n; ++i) {
    run_one<DEG>(&s[(size_t)i * 2 * DEG], &c[(size_t)i * DEG], &sh[(size_t)i * 2 * DEG], &ch[(size_t)i * DEG], mult);
    for (int e = 0; e != DEG; ++e) {
      bool m = ch[i * DEG + e] != co[i * DEG + e] || sh[(i * DEG + e) * 2] != so[(i * DEG + e) * 2] ||
               sh[(i * DEG + e) * 2 + 1] != so[(i * DEG + e) * 2 + 1];
      if (m && bad++ < 6) {
        printf("DEG %d check %d edge %d: in s=%08x %08x c=%08x | host c=%08x s=%08x %08x | gpu c=%08x s=%08x %08x\n", DEG, i, e,
               s[(i * DEG + e) * 2], s[(i * DEG + e) * 2 + 1], c[i * DEG + e], ch[i * DEG + e], sh[(i * DEG + e) * 2],
               sh[(i * DEG + e) * 2 + 1], co[i * DEG + e], so[(i * DEG + e) * 2], so[(i * DEG + e) * 2 + 1]);
      }
    }
  }
  printf("DEG %d mult %u: %d mismatching edges of %d (%s)\n", DEG, mult, bad, n * DEG, cudaGetErrorString(err));
  return bad;
}

int main()
{
  srand(1);
  int bad = 0;
  bad += test<3>(4000, 52428);
  bad += test<8>(4000, 52428);
  bad += test<10>(4000, 52428);
  bad += test<19>(4000, 52428);
  bad += test<19>(4000, 0);
  bad += test<5>(4000, 52428);
  return bad != 0;
}
