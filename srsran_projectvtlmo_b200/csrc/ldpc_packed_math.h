// Packed (four code blocks per thread) arithmetic of the layered normalized min-sum LDPC decoder.
//
// Four code blocks with the same base graph and lifting size are decoded together: thread j owns lifted check j of
// every layer for all four, in 2 x u16x2 registers (register 0 = code blocks 0 and 2, register 1 = code blocks 1 and 3).
// All quantities are kept as BIASED UNSIGNED 16-bit lanes so that negation and subtraction are plain 32-bit integer
// operations (no inter-lane borrow; they can issue on the FMA pipe as IMAD) and only min / max / select / permute
// use the ALU pipe (VIMNMX.U16x2, VIADDMNMX.[US]16x2, LOP3, PRMT - all single SASS instructions on sm_100a):
//
//   soft value  S in [-120, 120] or +-INF(8192)   stored as S + BS,  BS = 0x8080   (shared memory, 2 bytes per lane)
//   c2v message c in [-120, 120]                  stored as c + 128                 (shared memory, 1 byte per lane)
//   v2c         q = S - c                         held as   q + 0x8000  => bit 15 set <=> q >= 0
//   |q|                                           held as |q| + 0x8000
//
// "Infinite" soft values (+-127 in the reference: fillers and promoted sums) are held as +-8192 so that they survive
// the subtraction of any message, never win a minimum (minima start at 120) and stay infinite under the update.
//
// Reference arithmetic reproduced bit for bit (AVX2 / AVX-512 flavour):
//   v2c   : ldpc_decoder_avx512.cpp:81-121    clamp(soft - c2v, +-120), +-127 sticky
//   min   : ldpc_decoder_avx512.cpp:123-165   min1 / min2 start at 120, strict '<', sign product (0 counts as +)
//   c2v   : ldpc_decoder_avx512.cpp:167-216   (min-excluding-self * 52428) >> 16 with sign
//   soft  : ldpc_decoder_avx512.cpp:218-259   promotion sum: |sum| > 120 -> +-127, +-127 sticky
//
// This header is plain C++ (host emulation of the few intrinsics) unless compiled for the device, so that the same
// code is exercised on the CPU against the oracle (tests/test_packed_math_cpu.py via tools/packed_math_harness.cpp).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PK_FN __host__ __device__ __forceinline__
#define PK_MFN __host__ __device__ __forceinline__
#else
#define PK_FN static inline
#define PK_MFN inline
#endif

namespace pusch_dec {
namespace pk {

/// The same 16-bit value in both lanes.
#define PK_REP2(v) ((((uint32_t)(v)) & 0xffffU) | (((uint32_t)(v)) << 16))

constexpr uint32_t BS      = 0x8080U; ///< bias of stored soft values
constexpr uint32_t BQ      = 0x8000U; ///< bias of v2c values and magnitudes
constexpr uint32_t INF     = 8192U;   ///< magnitude standing for the reference's +-127
constexpr uint32_t INF_CUT = 4096U;
constexpr uint32_t SOFT_ZERO2 = PK_REP2(0x8080U);
constexpr uint32_t C2V_ZERO4  = 0x80808080U;

#define PK_LANES(expr_lo, expr_hi) ((uint32_t)(uint16_t)(expr_lo) | ((uint32_t)(uint16_t)(expr_hi) << 16))
PK_FN uint32_t maxu2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  return __vmaxu2(a, b);
#else
  uint16_t al = a, ah = a >> 16, bl = b, bh = b >> 16;
  return PK_LANES(al > bl ? al : bl, ah > bh ? ah : bh);
#endif
}
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
/// min(a + b, c) on unsigned lanes, the sum wrapping modulo 2^16 (add.u16x2 + min.u16x2 = VIADDMNMX.U16x2).
PK_FN uint32_t addmin_u2(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
  return __viaddmin_u16x2(a, b, c);
#else
  return minu2(add2(a, b), c);
#endif
}
PK_FN uint32_t addmax_u2(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
  return __viaddmax_u16x2(a, b, c);
#else
  return maxu2(add2(a, b), c);
#endif
}
PK_FN uint32_t addmax_s2(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
  return __viaddmax_s16x2(a, b, c);
#else
  uint32_t s  = add2(a, b);
  int16_t  sl = (int16_t)(uint16_t)s, sh = (int16_t)(uint16_t)(s >> 16), cl = (int16_t)(uint16_t)c,
          ch = (int16_t)(uint16_t)(c >> 16);
  return PK_LANES(sl > cl ? sl : cl, sh > ch ? sh : ch);
#endif
}
/// max(min(a + b, c), 0) on signed lanes.
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
/// 0xffff in every lane whose bit 15 is set.
PK_FN uint32_t lane_mask(uint32_t a)
{
#if defined(__CUDA_ARCH__)
  // prmt's sign-replicate mode (selector nibble bit 3). NOT __byte_perm: that intrinsic ignores bit 3 of the nibbles.
  uint32_t r;
  asm("prmt.b32 %0, %1, 0, 0xbb99;" : "=r"(r) : "r"(a));
  return r;
#else
  return ((a & 0x8000U) ? 0xffffU : 0U) | ((a & 0x80000000U) ? 0xffff0000U : 0U);
#endif
}

PK_FN uint32_t sel(uint32_t mask, uint32_t a, uint32_t b)
{
  return (a & mask) | (b & ~mask);
}

/// Lane-wise (m * mult) >> 16 on true magnitudes m <= 120 (mm512::scale_epi8, avx512_support.h:65-107); mult == 0: identity.
PK_FN uint32_t scale2(uint32_t m, uint32_t mult)
{
  if (mult == 0) {
    return m;
  }
  uint32_t lo = ((m & 0xffffU) * mult) >> 16;
  uint32_t hi = ((m >> 16) * mult) & 0xffff0000U;
  return lo | hi;
}

/// Packed soft value (two lanes) of two int8 LLRs x0 (low lane) and x1 (high lane) given as biased bytes ub = x ^ 0x80
/// in bits 0-7 and 16-23 of `ub2`: +-127 become +-INF.
PK_FN uint32_t soft_from_biased_bytes(uint32_t ub2)
{
  uint32_t lane = ub2 | 0x80008000U;                                            // x + 0x8080 (x + 128 + 0x8000)
  uint32_t tp   = addmin_s2_relu(lane, PK_REP2(0x10000U - 0x80feU), 0x00010001U);  // 1 where ub == 255 (x == 127)
  uint32_t nl   = 0x01010100U - lane;                                           // 2 * 0x8080 - lane, lane-wise
  uint32_t tn   = addmin_s2_relu(nl, PK_REP2(0x10000U - 0x80feU), 0x00010001U);    // 1 where x == -127
  return lane + tp * (INF - 127U) - tn * (INF - 127U);
}

/// State of one lifted check while a layer is processed, for 2 * NR code blocks (NR registers of two 16-bit lanes).
/// KEEP_A: keep |v2c| of every edge in registers between the two passes (else it is recomputed: 2 instructions).
template <int DEG, int NR, bool KEEP_A = true>
struct check_lanes {
  uint32_t q[DEG][NR];                  ///< v2c + BQ
  uint32_t a[KEEP_A ? DEG : 1][NR];     ///< |v2c| + BQ
  uint32_t m1[NR], m2[NR], x[NR];
  uint32_t nm1[NR], pm1[NR], pm2[NR];

  PK_MFN void begin()
  {
#pragma unroll
    for (int r = 0; r != NR; ++r) {
      m1[r] = m2[r] = PK_REP2(BQ + 120U);
      x[r]          = 0;
    }
  }

  /// Edge e: soft words s[r] (S + BS per lane) and messages of the previous iteration c[r] (c + 128 per lane).
  PK_MFN void gather(int e, const uint32_t* s, const uint32_t* c)
  {
#pragma unroll
    for (int r = 0; r != NR; ++r) {
      uint32_t qq = s[r] - c[r];          // lanes never borrow: S + BS > c + 128
      uint32_t nq = 0x00010000U - qq;     // lane-wise 0x10000 - q
      uint32_t aa = maxu2(qq, nq);
      q[e][r]     = qq;
      if (KEEP_A) {
        a[e][r] = aa;
      }
      x[r] ^= qq;
      uint32_t t = maxu2(m1[r], aa);
      m2[r]      = minu2(m2[r], t);
      m1[r]      = minu2(m1[r], aa);
    }
  }

  PK_MFN void reduce(uint32_t mult)
  {
#pragma unroll
    for (int r = 0; r != NR; ++r) {
      uint32_t s1 = scale2(m1[r] & 0x7fff7fffU, mult);
      uint32_t s2 = scale2(m2[r] & 0x7fff7fffU, mult);
      nm1[r]      = 0x00010000U - m1[r];
      // bit 15 of x = parity of the non-negative v2c; the sign product P is negative iff (#negative) is odd.
      uint32_t pmk = lane_mask((DEG & 1) ? ~x[r] : x[r]);
      // 128 + P * M for the two candidate magnitudes (min1: every edge but the minimum, min2: the minimum itself).
      pm1[r] = sel(pmk, 0x00800080U - s1, 0x00800080U + s1);
      pm2[r] = sel(pmk, 0x00800080U - s2, 0x00800080U + s2);
    }
  }

  /// New soft words sn[r] and new messages cn[r] (128 + c per lane) of edge e.
  PK_MFN void scatter(int e, uint32_t* sn, uint32_t* cn)
  {
#pragma unroll
    for (int r = 0; r != NR; ++r) {
      uint32_t qq = q[e][r];
      uint32_t aa = KEEP_A ? a[e][r] : maxu2(qq, 0x00010000U - qq);
      uint32_t t    = addmin_u2(aa, nm1[r], 0x00010001U); // 1: |q| > min1 -> min1, 0: this edge is the minimum -> min2
      uint32_t mask = t * 0xffffU;
      uint32_t pm   = sel(mask, pm1[r], pm2[r]);          // 128 + P * M
      uint32_t sq   = lane_mask(qq);                      // lanes with q >= 0
      // new message c = sign(q) * P * M (the sign product of the OTHER edges is P * sign(q)), stored as 128 + c
      cn[r] = sel(sq, pm, 0x01000100U - pm);
      // soft = sign(q) * promote(min(|q|, 120) + P * M); infinite |q| (>= INF - 120) passes the clamp.
      uint32_t av = addmax_u2(aa, PK_REP2(0x10000U - INF_CUT), minu2(aa, PK_REP2(BQ + 120U)));
      uint32_t v  = av + pm;                               // bias BQ + 128 = BS
      uint32_t t2 = addmin_s2_relu(v, PK_REP2(0x10000U - (BS + 120U)), 0x00010001U);
      uint32_t w  = minu2(v + t2 * INF, PK_REP2(BS + INF));
      uint32_t nw = 0x01010100U - w;                       // lane-wise 2 * BS - w
      sn[r]       = sel(sq, w, nw);
    }
  }
};

/// Four code blocks per thread, messages of one lifted edge packed in one 32-bit word (byte c = code block c): register 0
/// holds code blocks 0 and 2, register 1 code blocks 1 and 3.
template <int DEG>
struct check4 : check_lanes<DEG, 2> {
  PK_MFN void gather(int e, uint32_t s0, uint32_t s1, uint32_t cw)
  {
#if defined(__CUDA_ARCH__)
    uint32_t c[2] = {cw & 0x00ff00ffU, __byte_perm(cw, 0, 0x4341)};
#else
    uint32_t c[2] = {cw & 0x00ff00ffU, (cw >> 8) & 0x00ff00ffU};
#endif
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
    return cn[0] + (cn[1] << 8);
  }
};

/// Two code blocks per thread (one register), messages of one lifted edge in one 16-bit word (low byte = lane 0).
template <int DEG>
struct check2 : check_lanes<DEG, 1, false> {
  PK_MFN void gather(int e, uint32_t s0, uint32_t cw16)
  {
    uint32_t c = (cw16 & 0xffU) | ((cw16 & 0xff00U) << 8);
    check_lanes<DEG, 1, false>::gather(e, &s0, &c);
  }

  PK_MFN uint32_t scatter(int e, uint32_t& s0)
  {
    uint32_t cn, sn;
    check_lanes<DEG, 1, false>::scatter(e, &sn, &cn);
    s0 = sn;
    return (cn & 0xffU) | (cn >> 8); // bits 0-7 and 16-23 -> one 16-bit word
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
