// Packed (four code blocks per thread) arithmetic of the layered normalized min-sum LDPC decoder.
//
// Four code blocks with the same base graph and lifting size are decoded together: thread j owns lifted check j of
// every layer for all four, in 2 registers of two 16-bit lanes (register 0 = code blocks 0 and 2, register 1 = code
// blocks 1 and 3). The lanes are IEEE binary16 numbers and the per-edge arithmetic is done with half2 instructions:
// every finite quantity of the decoder is an integer of magnitude <= 1392 (exact in binary16), so the results are the
// reference's integers bit for bit, while most of the work issues on the FMA pipe (HFMA2 / HADD2 with free |x|, -x and
// saturate-to-[0,1] modifiers) instead of the ALU pipe that bounds an integer formulation (measured on sm_100a,
// tools/ubench/pipes.cu: HFMA2 runs beside VIMNMX / LOP3 / PRMT; HMNMX2 and HSET2 share the ALU pipe):
//
//   soft value  S in [-120, 120] or "infinite"   stored as the half S + 1152           (shared memory, 2 bytes per lane)
//   c2v message c in [-120, 120]                 stored as the byte c + 128            (shared memory, 1 byte per lane)
//   a message byte b becomes the half 1024 + b = c + 1152 by placing 0x64 above it (one PRMT), so
//   v2c         q = S - c  =  half(S + 1152) - half(c + 1152)                          (exact, |q| <= 248)
//
// "Infinite" soft values (+-127 in the reference: fillers and promoted sums) are halves of magnitude >= 5800 (up to and
// including the IEEE infinities): they survive the subtraction of any message, never win a minimum (minima start at
// 120) and stay infinite under the update. No operation below can produce a NaN (no inf - inf, no 0 x inf).
//
// Reference arithmetic reproduced bit for bit (AVX2 / AVX-512 flavour):
//   v2c   : ldpc_decoder_avx512.cpp:81-121    clamp(soft - c2v, +-120), +-127 sticky
//   min   : ldpc_decoder_avx512.cpp:123-165   min1 / min2 start at 120, strict '<', sign product (0 counts as +)
//   c2v   : ldpc_decoder_avx512.cpp:167-216   (min-excluding-self * 52428) >> 16 with sign
//   soft  : ldpc_decoder_avx512.cpp:218-259   promotion sum: |sum| > 120 -> +-127, +-127 sticky
//
// This header is plain C++ (host emulation of the few intrinsics, halves through _Float16) unless compiled for the
// device, so that the same code is exercised on the CPU against the oracle (tests/test_packed_math_cpu.py via
// tests/host_emul/packed_math_harness.cpp).
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#else
#include <string.h>
#endif

#if defined(__CUDACC__)
#define PK_FN __host__ __device__ __forceinline__
#define PK_MFN __host__ __device__ __forceinline__
#if defined(__CUDA_ARCH__)
#define PK_UNROLL _Pragma("unroll")
#else
#define PK_UNROLL
#endif
#else
#define PK_FN static inline
#define PK_MFN inline
#define PK_UNROLL
#endif

namespace pusch_dec {
namespace pk {

/// The same 16-bit value in both lanes.
#define PK_REP2(v) ((((uint32_t)(v)) & 0xffffU) | (((uint32_t)(v)) << 16))

// binary16 constants (bit patterns).
constexpr uint32_t H_ONE     = 0x3c00U; ///< 1
constexpr uint32_t H_120     = 0x5780U; ///< 120
constexpr uint32_t H_128     = 0x5800U; ///< 128
constexpr uint32_t H_1024    = 0x6400U; ///< 1024
constexpr uint32_t H_1152    = 0x6480U; ///< 1152 = 1024 + 128
constexpr uint32_t H_8192    = 0x7000U; ///< 8192
constexpr uint32_t H_INV128  = 0x2000U; ///< 1 / 128
constexpr uint32_t H_M15_16  = 0xbb80U; ///< -120 / 128
constexpr uint32_t H_POS_INF = 0x7090U; ///< +9344: soft value of a +127 input
constexpr uint32_t H_NEG_INF = 0xeee0U; ///< -7040: soft value of a -127 input

constexpr uint32_t BS         = H_1152; ///< bit pattern of a zero soft value (patterns compare like SIGNED 16-bit integers)
constexpr uint32_t SOFT_ZERO2 = PK_REP2(H_1152);
constexpr uint32_t C2V_ZERO4  = 0x80808080U;

#define PK_LANES(expr_lo, expr_hi) ((uint32_t)(uint16_t)(expr_lo) | ((uint32_t)(uint16_t)(expr_hi) << 16))

// ---- integer helpers (loader, hard decision) -----------------------------------------------------------------------------
PK_FN uint32_t minu2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  return __vminu2(a, b);
#else
  uint16_t al = a, ah = a >> 16, bl = b, bh = b >> 16;
  return PK_LANES(al < bl ? al : bl, ah < bh ? ah : bh);
#endif
}
PK_FN uint32_t add2(uint32_t a, uint32_t b)
{
  return PK_LANES((uint16_t)a + (uint16_t)b, (uint16_t)(a >> 16) + (uint16_t)(b >> 16));
}
/// max(min(a + b, c), 0) on signed 16-bit lanes (VIADDMNMX.S16x2.RELU).
PK_FN uint32_t addmin_s2_relu(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
  return __viaddmin_s16x2_relu(a, b, c);
#else
  uint32_t s  = add2(a, b);
  int16_t  sl = (int16_t)(uint16_t)s, sh = (int16_t)(uint16_t)(s >> 16), cl = (int16_t)(uint16_t)c,
          ch = (int16_t)(uint16_t)(c >> 16);
  int16_t rl = sl < cl ? sl : cl, rh = sh < ch ? sh : ch;
  return PK_LANES(rl < 0 ? 0 : rl, rh < 0 ? 0 : rh);
#endif
}
/// 1 in every lane whose soft value is > 0.
PK_FN uint32_t positive_lanes(uint32_t s)
{
  return addmin_s2_relu(s, PK_REP2(0x10000U - BS), 0x00010001U);
}
/// Hard decision of one 16-bit soft pattern: the bit is 1 for soft <= 0 (log_likelihood_ratio.cpp:226-252).
PK_FN bool lane_hard_bit(uint32_t lane16)
{
  return (int16_t)(uint16_t)lane16 <= (int16_t)BS;
}

PK_FN uint32_t prmt2(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, b, sel);
#else
  uint64_t all = ((uint64_t)b << 32) | a;
  uint32_t r   = 0;
  for (int i = 0; i != 4; ++i) {
    r |= (uint32_t)((all >> (8 * ((sel >> (4 * i)) & 7U))) & 0xffU) << (8 * i);
  }
  return r;
#endif
}

/// Lane-wise (m * mult) >> 16 on integer magnitudes m <= 120 (mm512::scale_epi8, avx512_support.h:65-107); mult == 0:
/// identity.
PK_FN uint32_t scale2(uint32_t m, uint32_t mult)
{
  if (mult == 0) {
    return m;
  }
  uint32_t lo = ((m & 0xffffU) * mult) >> 16;
  uint32_t hi = ((m >> 16) * mult) & 0xffff0000U;
  return lo | hi;
}

// ---- half2 helpers: device = one SASS instruction each (modifiers folded), host = _Float16 emulation ----------------------
#if !defined(__CUDA_ARCH__)
namespace host {
inline double h2d(uint32_t bits16)
{
  uint16_t  u = (uint16_t)bits16;
  _Float16 f;
  memcpy(&f, &u, 2);
  return (double)f;
}
inline uint32_t d2h(double d)
{
  _Float16 f = (_Float16)d; // one rounding to nearest even; overflow gives an infinity
  uint16_t u;
  memcpy(&u, &f, 2);
  return u;
}
inline double sat01(double d)
{
  return d != d ? 0.0 : (d < 0.0 ? 0.0 : (d > 1.0 ? 1.0 : d));
}
inline double dabs(double d)
{
  return d < 0 ? -d : d;
}
template <typename F>
inline uint32_t lanes2(uint32_t a, uint32_t b, uint32_t c, F f)
{
  return PK_LANES(d2h(f(h2d(a), h2d(b), h2d(c))), d2h(f(h2d(a >> 16), h2d(b >> 16), h2d(c >> 16))));
}
} // namespace host
#else
__device__ __forceinline__ __half2 pk_h(uint32_t u)
{
  return *reinterpret_cast<__half2*>(&u);
}
__device__ __forceinline__ uint32_t pk_u(__half2 h)
{
  return *reinterpret_cast<uint32_t*>(&h);
}
#endif

/// a - b
PK_FN uint32_t hsub2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hsub2(pk_h(a), pk_h(b)));
#else
  return host::lanes2(a, b, 0, [](double x, double y, double) { return x - y; });
#endif
}
/// a + b
PK_FN uint32_t hadd2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hadd2(pk_h(a), pk_h(b)));
#else
  return host::lanes2(a, b, 0, [](double x, double y, double) { return x + y; });
#endif
}
/// a * b + c
PK_FN uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hfma2(pk_h(a), pk_h(b), pk_h(c)));
#else
  return host::lanes2(a, b, c, [](double x, double y, double z) { return x * y + z; });
#endif
}
/// min(a, b)
PK_FN uint32_t hmin2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hmin2(pk_h(a), pk_h(b)));
#else
  return host::lanes2(a, b, 0, [](double x, double y, double) { return x < y ? x : y; });
#endif
}
/// min(|q|, m)
PK_FN uint32_t hmin2_abs(uint32_t q, uint32_t m)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hmin2(__habs2(pk_h(q)), pk_h(m)));
#else
  return host::lanes2(q, m, 0, [](double x, double y, double) { return host::dabs(x) < y ? host::dabs(x) : y; });
#endif
}
/// max(|q|, m)
PK_FN uint32_t hmax2_abs(uint32_t q, uint32_t m)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hmax2(__habs2(pk_h(q)), pk_h(m)));
#else
  return host::lanes2(q, m, 0, [](double x, double y, double) { return host::dabs(x) > y ? host::dabs(x) : y; });
#endif
}
/// saturate(|q| - m): for integers, 1 where |q| > m and 0 elsewhere.
PK_FN uint32_t habs_gt(uint32_t q, uint32_t m)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hadd2_sat(__habs2(pk_h(q)), __hneg2(pk_h(m))));
#else
  return host::lanes2(q, m, 0, [](double x, double y, double) { return host::sat01(host::dabs(x) - y); });
#endif
}
/// saturate(a + b)
PK_FN uint32_t hadd2_sat(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hadd2_sat(pk_h(a), pk_h(b)));
#else
  return host::lanes2(a, b, 0, [](double x, double y, double) { return host::sat01(x + y); });
#endif
}
/// saturate(|q| * b + c)
PK_FN uint32_t habs_fma_sat(uint32_t q, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hfma2_sat(__habs2(pk_h(q)), pk_h(b), pk_h(c)));
#else
  return host::lanes2(q, b, c, [](double x, double y, double z) { return host::sat01(host::dabs(x) * y + z); });
#endif
}
/// a * b + |q|
PK_FN uint32_t hfma2_abs_c(uint32_t a, uint32_t b, uint32_t q)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hfma2(pk_h(a), pk_h(b), __habs2(pk_h(q))));
#else
  return host::lanes2(a, b, q, [](double x, double y, double z) { return x * y + host::dabs(z); });
#endif
}

/// max(m - |q|, 0)
PK_FN uint32_t habs_rsub_relu(uint32_t q, uint32_t m)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hfma2_relu(__habs2(pk_h(q)), pk_h(PK_REP2(H_ONE | 0x8000U)), pk_h(m)));
#else
  return host::lanes2(q, m, 0, [](double x, double y, double) { return y - host::dabs(x) > 0 ? y - host::dabs(x) : 0.0; });
#endif
}
/// |q| + b
PK_FN uint32_t habs_add(uint32_t q, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hadd2(__habs2(pk_h(q)), pk_h(b)));
#else
  return host::lanes2(q, b, 0, [](double x, double y, double) { return host::dabs(x) + y; });
#endif
}
/// 1.0 where |q| > m, else 0.0 (HSET2.BF.GT: ALU pipe)
PK_FN uint32_t habs_gt_set(uint32_t q, uint32_t m)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hgt2(__habs2(pk_h(q)), pk_h(m)));
#else
  return host::lanes2(q, m, 0, [](double x, double y, double) { return host::dabs(x) > y ? 1.0 : 0.0; });
#endif
}
/// 1.0 where a > b, else 0.0
PK_FN uint32_t hgt_set(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hgt2(pk_h(a), pk_h(b)));
#else
  return host::lanes2(a, b, 0, [](double x, double y, double) { return x > y ? 1.0 : 0.0; });
#endif
}
/// max(a, b)
PK_FN uint32_t hmax2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  return pk_u(__hmax2(pk_h(a), pk_h(b)));
#else
  return host::lanes2(a, b, 0, [](double x, double y, double) { return x > y ? x : y; });
#endif
}

// Which pipe the interchangeable steps use (all variants compute the same values; the split balances the ALU and FMA pipes
// inside each phase of a layer, because the warps of a CTA run the phases in lockstep between the layer barriers).
#ifndef PK_GATHER_FMA_REGS
#define PK_GATHER_FMA_REGS 1 ///< registers (of NR) whose min1 / min2 update runs on the FMA pipe (3 HFMA2/HADD2 for 2 HMNMX2)
#endif
#ifndef PK_SCATTER_T_ALU
#define PK_SCATTER_T_ALU 1 ///< |q| > min1 by HSET2 (ALU) instead of a saturating HADD2 (FMA)
#endif
#ifndef PK_SCATTER_T2_ALU
#define PK_SCATTER_T2_ALU 1 ///< v > 120 by HSET2 (ALU) instead of a saturating HADD2 (FMA)
#endif
#ifndef PK_SCATTER_CLAMP_ALU
#define PK_SCATTER_CLAMP_ALU 0 ///< min(|q|, 120) with infinite pass-through by 2 HMNMX2 + HADD2 instead of 2 HFMA2
#endif

#if defined(__CUDACC__)
/// Sign mask and 1.0 of both lanes, read from constant memory so that they stay register / constant-bank operands: as
/// immediates the and-or below needs two LOP3 (one immediate per instruction).
__constant__ uint32_t c_pk_sign_one[2] = {0x80008000U, PK_REP2(H_ONE)};
#endif
/// sign(q) as +-1 in both lanes (+1 for a zero).
PK_FN uint32_t sign_one(uint32_t q)
{
#if defined(__CUDA_ARCH__)
  return (q & c_pk_sign_one[0]) | c_pk_sign_one[1];
#else
  return (q & 0x80008000U) | PK_REP2(H_ONE);
#endif
}

/// Packed soft value (two lanes) of two int8 LLRs x0 (low lane) and x1 (high lane) given as biased bytes ub = x ^ 0x80
/// in bits 0-7 and 16-23 of `ub2`: the half x + 1152 = 0x6400 | ub; +-127 become infinite.
PK_FN uint32_t soft_from_biased_bytes(uint32_t ub2)
{
  uint32_t lane = ub2 | PK_REP2(H_1024);
  uint32_t tp   = addmin_s2_relu(lane, PK_REP2(0x10000U - 0x64feU), 0x00010001U); // 1 where ub == 255 (x == 127)
  uint32_t tn   = addmin_s2_relu(~lane, PK_REP2(0x6403U), 0x00010001U);           // 1 where ub <= 1   (x <= -127)
  return lane + tp * (H_POS_INF - 0x64ffU) + tn * (H_NEG_INF - 0x6401U);           // no carry between the lanes
}

/// Running state of one lifted check while a layer is processed, for 2 * NR code blocks (NR registers of two binary16
/// lanes): minima, sign product and the two candidate output magnitudes. The v2c values themselves are not kept here:
/// check_lanes (below) holds them in registers between the two passes; the TMEM-message kernel (ldpc_packed.cuh) parks
/// them in the soft array instead.
template <int NR>
struct check_acc {
  uint32_t m1[NR], m2[NR], x[NR];
  uint32_t pm2[NR], dpm[NR];

  PK_MFN void begin()
  {
PK_UNROLL
    for (int r = 0; r != NR; ++r) {
      m1[r] = m2[r] = PK_REP2(H_120);
      x[r]          = 0;
    }
  }

  /// One edge: soft words s[r] (half S + 1152 per lane) and messages of the previous iteration c[r] (half c + 1152 per
  /// lane); returns the v2c words in qq[r].
  PK_MFN void gather_q(const uint32_t* s, const uint32_t* c, uint32_t* qo)
  {
PK_UNROLL
    for (int r = 0; r != NR; ++r) {
      uint32_t qq = hsub2(s[r], c[r]);
      qo[r]       = qq;
      x[r] ^= qq; // bit 15 of every lane: parity of the negative v2c (a zero v2c is +0: it counts as positive)
      if (r < PK_GATHER_FMA_REGS) {
        uint32_t d = habs_rsub_relu(qq, m1[r]); // max(min1 - |q|, 0): exact, 0 for an infinite |q|
        uint32_t t = habs_add(qq, d);            // max(min1, |q|)
        m1[r]      = hsub2(m1[r], d);            // min(min1, |q|)
        m2[r]      = hmin2(m2[r], t);
      } else {
        uint32_t t = hmax2_abs(qq, m1[r]);
        m2[r]      = hmin2(m2[r], t);
        m1[r]      = hmin2_abs(qq, m1[r]);
      }
    }
  }

  PK_MFN void reduce(uint32_t mult)
  {
PK_UNROLL
    for (int r = 0; r != NR; ++r) {
      // The two minima are integers in [0, 120]: the half 1024 + m has m in its low mantissa bits, and back.
      uint32_t i1  = hadd2(m1[r], PK_REP2(H_1024)) & 0x00ff00ffU;
      uint32_t i2  = hadd2(m2[r], PK_REP2(H_1024)) & 0x00ff00ffU;
      uint32_t s1  = hsub2(scale2(i1, mult) | PK_REP2(H_1024), PK_REP2(H_1024));
      uint32_t s2  = hsub2(scale2(i2, mult) | PK_REP2(H_1024), PK_REP2(H_1024));
      uint32_t sgn = x[r] & 0x80008000U; // the sign product P of all edges
      // P * M for the two candidate magnitudes (min1: every edge but the minimum, min2: the minimum itself).
      uint32_t pm1 = s1 ^ sgn;
      pm2[r]       = s2 ^ sgn;
      dpm[r]       = hsub2(pm1, pm2[r]);
    }
  }

  /// New soft words sn[r] (half S + 1152) and new messages cn[r] (half c + 1152: the message byte is the low byte of
  /// every lane) of the edge whose v2c words are qi[r].
  PK_MFN void scatter_q(const uint32_t* qi, uint32_t* sn, uint32_t* cn)
  {
PK_UNROLL
    for (int r = 0; r != NR; ++r) {
      uint32_t qq = qi[r];
      // 1: |q| > min1 -> min1, 0: this edge is the minimum -> min2
      uint32_t t  = PK_SCATTER_T_ALU ? habs_gt_set(qq, m1[r]) : habs_gt(qq, m1[r]);
      uint32_t pm = hfma2(t, dpm[r], pm2[r]);                 // P * M
      uint32_t sq = sign_one(qq);
      // new message c = sign(q) * P * M (the sign product of the OTHER edges is P * sign(q))
      cn[r] = hfma2(sq, pm, PK_REP2(H_1152));
      // soft = sign(q) * promote(min(|q|, 120) + P * M); an infinite |q| stays infinite.
      uint32_t av;
      if (PK_SCATTER_CLAMP_ALU) {
        av = hmax2(hmin2_abs(qq, PK_REP2(H_120)), habs_add(qq, PK_REP2(0xec00U))); // max(min(|q|, 120), |q| - 4096)
      } else {
        // u = saturate((|q| - 120) / 128) is exact for |q| <= 248, so |q| - 128 u = min(|q|, 120) for every finite |q|.
        uint32_t u = habs_fma_sat(qq, PK_REP2(H_INV128), PK_REP2(H_M15_16));
        av         = hfma2_abs_c(u, PK_REP2(H_128 | 0x8000U), qq);
      }
      uint32_t v  = hadd2(av, pm);
      // 1 where v > 120 (v is an integer or infinite)
      uint32_t t2 = PK_SCATTER_T2_ALU ? hgt_set(v, PK_REP2(H_120)) : hadd2_sat(v, PK_REP2(H_120 | 0x8000U));
      uint32_t w  = hfma2(t2, PK_REP2(H_8192), v);
      sn[r]       = hfma2(sq, w, PK_REP2(H_1152));
    }
  }
};

/// State of one lifted check while a layer is processed, v2c values kept in registers between the passes.
template <int DEG, int NR>
struct check_lanes : check_acc<NR> {
  uint32_t q[DEG][NR]; ///< v2c

  /// Edge e: soft words s[r] (half S + 1152 per lane) and messages of the previous iteration c[r] (half c + 1152 per lane).
  PK_MFN void gather(int e, const uint32_t* s, const uint32_t* c)
  {
    check_acc<NR>::gather_q(s, c, q[e]);
  }

  /// New soft words sn[r] (half S + 1152) and new messages cn[r] (half c + 1152: the message byte is the low byte of
  /// every lane) of edge e.
  PK_MFN void scatter(int e, uint32_t* sn, uint32_t* cn)
  {
    check_acc<NR>::scatter_q(q[e], sn, cn);
  }
};

/// Four code blocks per thread, messages of one lifted edge packed in one 32-bit word (byte c = code block c): register 0
/// holds code blocks 0 and 2, register 1 code blocks 1 and 3.
template <int DEG>
struct check4 : check_lanes<DEG, 2> {
  PK_MFN void gather(int e, uint32_t s0, uint32_t s1, uint32_t cw)
  {
    uint32_t c[2] = {prmt2(cw, 0x64646464U, 0x4240U), prmt2(cw, 0x64646464U, 0x4341U)};
    uint32_t s[2] = {s0, s1};
    check_lanes<DEG, 2>::gather(e, s, c);
  }

  /// New packed messages (returned) and new soft words of edge e.
  PK_MFN uint32_t scatter(int e, uint32_t& s0, uint32_t& s1)
  {
    uint32_t cn[2], sn[2];
    check_lanes<DEG, 2>::scatter(e, sn, cn);
    s0 = sn[0];
    s1 = sn[1];
    return prmt2(cn[0], cn[1], 0x6240U);
  }
};

/// Two code blocks per thread (one register), messages of one lifted edge in one 16-bit word (low byte = lane 0).
template <int DEG>
struct check2 : check_lanes<DEG, 1> {
  PK_MFN void gather(int e, uint32_t s0, uint32_t cw16)
  {
    uint32_t c = prmt2(cw16, 0x64646464U, 0x4140U);
    check_lanes<DEG, 1>::gather(e, &s0, &c);
  }

  PK_MFN uint32_t scatter(int e, uint32_t& s0)
  {
    uint32_t cn, sn;
    check_lanes<DEG, 1>::scatter(e, &sn, &cn);
    s0 = sn;
    return prmt2(cn, 0U, 0x4420U); // low bytes of the two lanes -> one 16-bit word
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// Intra-code-block packing: the four lanes hold four lifted checks of the SAME code block, j, j + Z/4, j + Z/2 and
// j + 3Z/4 (Z % 4 == 0), so a single code block of any shape runs on the packed arithmetic with Z/4 threads.
// Soft values are stored per "base" index b in [0, Z/4): word(col, b) = { r0 = (S[b], S[b + Z/2]), r1 = (S[b + Z/4],
// S[b + 3Z/4]) } (lanes L0..L3 = quarters 0..3 of the column: r0 = (L0, L2), r1 = (L1, L3)). For an edge with circulant
// shift s, check j + m Z/4 reads S[(j + m Z/4 + s) mod Z]: with k = (j + s) mod Z = q Z/4 + b these are the lanes
// (q + m) mod 4 of word(col, b) - a rotation of the four lanes by q quarter turns, done with two PRMTs.
// ---------------------------------------------------------------------------------------------------------------------

/// Byte-permute selectors (for prmt(r0, r1, sel)) that rotate the lanes (L0, L1, L2, L3) left by q: .x gives the new r0 =
/// (L[q], L[q+2]), .y the new r1 = (L[q+1], L[q+3]).
PK_FN void rot4_selectors(uint32_t q, uint32_t& sel0, uint32_t& sel1)
{
  // q = 0: r0, r1 | 1: r1, swap(r0) | 2: swap(r0), swap(r1) | 3: swap(r1), r0
  const uint32_t t0[4] = {0x3210U, 0x7654U, 0x1032U, 0x5476U};
  const uint32_t t1[4] = {0x7654U, 0x1032U, 0x5476U, 0x3210U};
  sel0 = t0[q & 3U];
  sel1 = t1[q & 3U];
}

/// Rotates the four lanes of (r0, r1) left by q quarter turns (see above).
PK_FN void rot4(uint32_t& r0, uint32_t& r1, uint32_t q)
{
  uint32_t s0, s1;
  rot4_selectors(q, s0, s1);
  uint32_t n0 = prmt2(r0, r1, s0);
  uint32_t n1 = prmt2(r0, r1, s1);
  r0          = n0;
  r1          = n1;
}

} // namespace pk
} // namespace pusch_dec
