/* TEST INFRASTRUCTURE (oracle): plain-C restatement of the reference's PUSCH soft-demodulation chain, x86 (AVX2 / AVX-512)
 * flavour, for SURVEY.md 8(f) row 2: soft demodulation -> descrambling -> UL-SCH demultiplexing (data only). Pinned
 * against the compiled reference (oracle/_ref, ref_demod_harness.cpp) by tests/test_oracle_demod_cpu.py and against golden
 * fixtures generated from it (tests/golden/demod.npz). Only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.
 *
 * The reference computes in binary32. Every rounding step below is the one the reference's object code performs when it is
 * built the way the reference builds it (g++ -O3 with FMA available, GNU fp-contract=fast): a product feeding a sum inside
 * one expression is ONE fused operation there (written fmaf() here); this file is compiled with -ffp-contract=off so that
 * nothing else is fused. A demodulation_mapper call treats the first floor(n / W) * W symbols of its block with SIMD code
 * and the rest with scalar code whose arithmetic differs (division instead of reciprocal multiplication, round-half-away
 * instead of round-half-even, a per-symbol instead of a per-component near-zero rule): both are restated, W = 16 / 8 / 16 /
 * 4 symbols for QPSK / 16QAM / 64QAM / 256QAM. */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define NEAR_ZERO 1e-9f

/* ---- scrambling sequence, TS 38.211 5.2.1 (pseudo_random_generator_impl.cpp: x1 from 1, x2 from c_init, Nc = 1600) ------ */
void oracle_scrambling_sequence(uint32_t c_init, uint8_t* bits, uint32_t n)
{
  uint32_t x1 = 1, x2 = c_init & 0x7fffffffU; /* bit i = x(n + i) */
  for (uint32_t i = 0; i != 1600; ++i) {
    uint32_t f1 = ((x1 >> 3) ^ x1) & 1U;
    uint32_t f2 = ((x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1) ^ x2) & 1U;
    x1          = (x1 >> 1) | (f1 << 30);
    x2          = (x2 >> 1) | (f2 << 30);
  }
  for (uint32_t i = 0; i != n; ++i) {
    bits[i]     = (uint8_t)((x1 ^ x2) & 1U);
    uint32_t f1 = ((x1 >> 3) ^ x1) & 1U;
    uint32_t f2 = ((x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1) ^ x2) & 1U;
    x1          = (x1 >> 1) | (f1 << 30);
    x2          = (x2 >> 1) | (f2 << 30);
  }
}

/* ---- quantisers ------------------------------------------------------------------------------------------------------------ */
/* mm256::quantize_ps / mm512::quantize_ps (avx2_helpers.h:120-160, avx512_helpers.h:138-176): scale, clip, round to nearest
 * even, NaN -> 0. */
static int8_t quantize_simd(float v, float range_limit)
{
  float s = v * (120.0f / range_limit);
  if (s > 120.0f) {
    s = 120.0f;
  }
  if (s < -120.0f) {
    s = -120.0f;
  }
  if (s != s) {
    return 0;
  }
  return (int8_t)rintf(s); /* default rounding mode: nearest even */
}

/* log_likelihood_ratio::quantize (log_likelihood_ratio.cpp:88-97): clip, round half away from zero. */
static int8_t quantize_scalar(float v, float range_limit)
{
  float c = v;
  if (fabsf(v) > range_limit) {
    c = copysignf(range_limit, v);
  }
  return (int8_t)roundf(c / range_limit * 120);
}

/* ---- piecewise-linear max-log tables (demodulation_mapper_qam64.cpp:42-84, demodulation_mapper_qam256.cpp:42-172) ---------- */
typedef struct {
  float width;    /* interval width */
  float inv;      /* 1.0F / width, as the SIMD code computes it */
  int   n;        /* number of intervals */
  float slope[16], icpt[16];
} pw_table;

static pw_table T64[3], T256[4];
static float    SQ10, SQ2G; /* 1 / sqrt(10); 2 sqrt(2) */
static int      tables_ready;

static void fill(pw_table* t, float unit, float width_mult, int n, const int* slope_mult, const float* icpt_num, float icpt_den)
{
  t->width = width_mult * unit;
  t->inv   = 1.0f / t->width;
  t->n     = n;
  for (int i = 0; i != n; ++i) {
    t->slope[i] = (float)slope_mult[i] * unit;
    t->icpt[i]  = icpt_num[i] / icpt_den;
  }
}

static void init_tables(void)
{
  if (tables_ready) {
    return;
  }
  const float s42 = 1.0f / sqrtf(42.0f), s170 = 1.0f / sqrtf(170.0f);
  SQ10 = 1.0f / sqrtf(10.0f);
  SQ2G = 2.0f * 1.41421356237309504880f;
  {
    static const int   s01[8] = {16, 12, 8, 4, 4, 8, 12, 16};
    static const float i01[8] = {24, 12, 4, 0, 0, -4, -12, -24};
    static const int   s23[8] = {8, 4, 4, 8, -8, -4, -4, -8};
    static const float i23[8] = {20, 8, 8, 12, 12, 8, 8, 20};
    static const int   s45[4] = {4, -4, 4, -4};
    static const float i45[4] = {12, -4, -4, 12};
    fill(&T64[0], s42, 2, 8, s01, i01, 21);
    fill(&T64[1], s42, 2, 8, s23, i23, 21);
    fill(&T64[2], s42, 4, 4, s45, i45, 21);
  }
  {
    static const int   s01[16] = {32, 28, 24, 20, 16, 12, 8, 4, 4, 8, 12, 16, 20, 24, 28, 32};
    static const float i01[16] = {112, 84, 60, 40, 24, 12, 4, 0, 0, -4, -12, -24, -40, -60, -84, -112};
    static const int   s23[16] = {16, 12, 8, 4, 4, 8, 12, 16, -16, -12, -8, -4, -4, -8, -12, -16};
    static const float i23[16] = {88, 60, 36, 16, 16, 28, 36, 40, 40, 36, 28, 16, 16, 36, 60, 88};
    static const int   s45[16] = {8, 4, 4, 8, -8, -4, -4, -8, 8, 4, 4, 8, -8, -4, -4, -8};
    static const float i45[16] = {52, 24, 24, 44, -20, -8, -8, -12, -12, -8, -8, -20, 44, 24, 24, 52};
    static const int   s67[8]  = {4, -4, 4, -4, 4, -4, 4, -4};
    static const float i67[8]  = {28, -20, 12, -4, -4, 12, -20, 28};
    fill(&T256[0], s170, 2, 16, s01, i01, 85);
    fill(&T256[1], s170, 2, 16, s23, i23, 85);
    fill(&T256[2], s170, 2, 16, s45, i45, 85);
    fill(&T256[3], s170, 4, 8, s67, i67, 85);
  }
  tables_ready = 1;
}

static int clamp_idx(float f, int n)
{
  /* cvtps_epi32 of an integer-valued float; out-of-range and NaN give INT_MIN, which clips to 0. */
  int idx = (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : (int)0x80000000;
  if (idx == (int)0x80000000) {
    return 0;
  }
  idx += n / 2;
  return idx < 0 ? 0 : (idx > n - 1 ? n - 1 : idx);
}

/* mm256::interval_function (avx2_helpers.h:236-254) = mm512::interval_function (avx512_helpers.h:295-313). */
static float interval_simd(const pw_table* t, float x, float rcp)
{
  int   idx = clamp_idx(floorf(x * t->inv), t->n);
  float r   = fmaf(t->slope[idx], x, t->icpt[idx]) * rcp;
  return (fabsf(x) <= NEAR_ZERO) ? 0.0f : r;
}

/* interval_function (demodulation_mapper_intervals.h:34-64). */
static float interval_scalar(const pw_table* t, float x, float rcp)
{
  int   idx = clamp_idx(floorf(x / t->width), t->n);
  float l   = fmaf(t->slope[idx], x, t->icpt[idx]);
  return l * rcp;
}

static float safe_rcp(float nv)
{
  return (nv > 0) ? 1.0f / nv : 0.0f;
}

/* ---- demodulation_mapper::demodulate_soft for ONE block (demodulation_mapper_impl.cpp:76-106) ------------------------------- */
void oracle_demodulate_soft(int qm, int pi2, int8_t* llr, const float* sym, const float* nv, uint32_t n)
{
  init_tables();
  const uint32_t W     = (qm == 2) ? 16 : (qm == 4) ? 8 : (qm == 6) ? 16 : (qm == 8) ? 4 : 1;
  const uint32_t nsimd = (qm == 1) ? 0 : (n / W) * W;
  for (uint32_t i = 0; i != n; ++i) {
    const float re = sym[2 * i], im = sym[2 * i + 1], v = nv[i];
    int8_t*     o  = llr + (size_t)i * qm;
    const float xy[2] = {re, im};
    if (qm == 1) {
      /* demod_BPSK_symbol (demodulation_mapper_impl.cpp:33-41); pi/2-BPSK rotates the odd symbols (:55-73). */
      float a = re, b = im;
      if (pi2 && (i & 1U)) {
        a = im;
        b = -re;
      }
      o[0] = !(v > 0) ? 0 : quantize_scalar(SQ2G * (a + b) / v, 24);
      continue;
    }
    const int simd = i < nsimd;
    if (qm == 2) {
      for (int c = 0; c != 2; ++c) {
        if (simd) {
          o[c] = quantize_simd((SQ2G * xy[c]) * safe_rcp(v), 24); /* demodulation_mapper_qpsk.cpp:68-76 */
        } else {
          o[c] = !(v > 0) ? 0 : quantize_scalar(SQ2G * xy[c] / v, 24); /* :121-129 */
        }
      }
      continue;
    }
    const int zero_sym = (fmaf(re, re, im * im) < NEAR_ZERO); /* is_near_zero(cf_t), math_utils.h:91-94 (scalar tails only) */
    if (qm == 4) {
      const float g = 4.0f * SQ10, thr = 2 * SQ10;
      for (int c = 0; c != 2; ++c) {
        const float x = xy[c];
        if (simd) {
          /* demodulation_mapper_qam16.cpp:66-105 */
          const float rcp    = safe_rcp(v);
          const float first  = g * x;
          const float second = 2.0f * first - copysignf(0.8f, x);
          float       l01    = ((fabsf(x) > thr) ? second : first) * rcp;
          float       l23    = (0.8f - fabsf(first)) * rcp;
          if (fabsf(x) <= NEAR_ZERO) {
            l01 = l23 = 0.0f;
          }
          o[c]     = quantize_simd(l01, 20);
          o[2 + c] = quantize_simd(l23, 20);
        } else if (zero_sym || !(v > 0)) {
          o[c] = o[2 + c] = 0;
        } else {
          /* demod_16QAM_symbol_01 / _23 (:196-224) */
          float l = g * x;
          if (fabsf(x) > thr) {
            l = 2 * l - copysignf(0.8f, x);
          }
          o[c]     = quantize_scalar(l / v, 20);
          o[2 + c] = quantize_scalar(fmaf(-g, fabsf(x), 0.8f) / v, 20);
        }
      }
      continue;
    }
    const pw_table* T  = (qm == 6) ? T64 : T256;
    const int       np = qm / 2;
    if (!simd && zero_sym) {
      memset(o, 0, (size_t)qm);
      continue;
    }
    const float rcp = safe_rcp(v);
    for (int p = 0; p != np; ++p) {
      for (int c = 0; c != 2; ++c) {
        o[2 * p + c] = simd ? quantize_simd(interval_simd(&T[p], xy[c], rcp), 20)
                            : quantize_scalar(interval_scalar(&T[p], xy[c], rcp), 20);
      }
    }
  }
}

/* ---- pusch_demodulator_impl::demodulate (pusch_demodulator_impl.cpp:129-301) minus the equalizer, followed by
 * ulsch_demultiplex_impl without UCI (the bypass of ulsch_demultiplex_impl.cpp:253-275: soft bits pass through) ---------------
 * re_per_symbol[s]: data REs (per layer) of OFDM symbol s of the allocation (0: none); sym / nv: equalized symbols and noise
 * variances in the order [OFDM symbol][RE][layer]. Blocks of at most MAX_BLOCK_SIZE / (layers * qm) subcarriers, each one
 * demodulation_mapper call; revert_scrambling (:38-127) negates where the sequence bit is 1. Returns the number of LLRs. */
int oracle_pusch_demodulate(int qm, int pi2, uint32_t rnti, uint32_t n_id, uint32_t nof_layers, const uint32_t* re_per_symbol,
                            uint32_t nof_ofdm_symbols, const float* sym, const float* nv, int8_t* llr_out)
{
  const uint32_t max_block_subc = 4096 / (nof_layers * (uint32_t)qm);
  size_t         pos            = 0; /* equalized symbols consumed */
  size_t         total          = 0;
  for (uint32_t s = 0; s != nof_ofdm_symbols; ++s) {
    total += (size_t)re_per_symbol[s] * nof_layers;
  }
  static uint8_t* seq     = 0;
  static size_t   seq_cap = 0;
  if (seq_cap < total * (size_t)qm) {
    seq_cap = total * (size_t)qm;
    seq     = (uint8_t*)__builtin_realloc(seq, seq_cap);
  }
  oracle_scrambling_sequence((rnti << 15) + n_id, seq, (uint32_t)(total * (size_t)qm));
  for (uint32_t s = 0; s != nof_ofdm_symbols; ++s) {
    uint32_t left = re_per_symbol[s];
    while (left != 0) {
      uint32_t nsubc = left < max_block_subc ? left : max_block_subc;
      uint32_t nblk  = nsubc * nof_layers;
      oracle_demodulate_soft(qm, pi2, llr_out + pos * (size_t)qm, sym + 2 * pos, nv + pos, nblk);
      pos += nblk;
      left -= nsubc;
    }
  }
  for (size_t i = 0; i != total * (size_t)qm; ++i) {
    if (seq[i]) {
      llr_out[i] = (int8_t)(-llr_out[i]);
    }
  }
  return (int)(total * (size_t)qm);
}
