/* TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT PATH. See oracle_port.h for scope and parity status (PINNED).
 *
 * Scalar restatement of the reference's PUSCH channel-decoding arithmetic, AVX2/AVX-512 flavour. Every function cites
 * the reference file:line it follows. Written from the behaviour of the reference, not from its text: scalar loops,
 * one lifted check at a time, compact per-edge message storage.
 */
#include "oracle_port.h"
#include "nr_ldpc_bg_tables.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------------------------------
 * CRC (crc_calculator_lut_impl.cpp:33-38 polynomials; :66-152 + crc_calculator_lut_impl.h:92-103: equals the plain
 * bitwise CRC of exactly nbits bits).
 * ---------------------------------------------------------------------------------------------------------------- */
static void crc_params(int poly, uint32_t* gen, unsigned* order)
{
  switch (poly) {
    case ORACLE_CRC24A:
      *gen   = 0x1864CFB;
      *order = 24;
      break;
    case ORACLE_CRC24B:
      *gen   = 0x1800063;
      *order = 24;
      break;
    default:
      *gen   = 0x11021;
      *order = 16;
      break;
  }
}

uint32_t oracle_crc(int poly, const uint8_t* packed, uint32_t nbits)
{
  uint32_t gen;
  unsigned order;
  crc_params(poly, &gen, &order);
  uint32_t top = 1U << order, reg = 0;
  for (uint32_t i = 0; i != nbits; ++i) {
    uint32_t bit = (packed[i >> 3] >> (7 - (i & 7))) & 1U;
    reg          = (reg << 1) ^ (bit << order);
    if (reg & top) {
      reg ^= gen;
    }
  }
  /* The loop above divides by x^0; the CRC is the remainder of data * x^order, folded in through "bit << order". */
  return reg & (top - 1);
}

/* ------------------------------------------------------------------------------------------------------------------
 * LLR algebra (log_likelihood_ratio.cpp:39-86).
 * ---------------------------------------------------------------------------------------------------------------- */
static int is_inf(int8_t v)
{
  return (v == 127) || (v == -127);
}

int8_t oracle_llr_add(int8_t a, int8_t b)
{
  if (a == -b) {
    return 0;
  }
  if (is_inf(a)) {
    return a;
  }
  if (is_inf(b)) {
    return b;
  }
  int t = (int)a + (int)b;
  if (t > 120) {
    return 120;
  }
  if (t < -120) {
    return -120;
  }
  return (int8_t)t;
}

int8_t oracle_llr_promotion_sum(int8_t a, int8_t b)
{
  if (a == -b) {
    return 0;
  }
  if (is_inf(a)) {
    return a;
  }
  if (is_inf(b)) {
    return b;
  }
  int t = (int)a + (int)b;
  if (t > 120) {
    return 127;
  }
  if (t < -120) {
    return -127;
  }
  return (int8_t)t;
}

int oracle_hard_decision(uint8_t* out_packed, const int8_t* llrs, uint32_t n)
{
  int no_zero = 1;
  for (uint32_t i = 0; i != n; ++i) {
    uint8_t mask = (uint8_t)(0x80U >> (i & 7));
    if (llrs[i] <= 0) {
      out_packed[i >> 3] |= mask;
    } else {
      out_packed[i >> 3] &= (uint8_t)~mask;
    }
    if (llrs[i] == 0) {
      no_zero = 0;
    }
  }
  return no_zero;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Rate dematching (ldpc_rate_dematcher_impl.cpp:46-201, ldpc_rate_dematcher_avx512_impl.cpp:29-64).
 * ---------------------------------------------------------------------------------------------------------------- */
static int sat8(int v)
{
  return v > 127 ? 127 : (v < -128 ? -128 : v);
}

/* AVX-512 combine: the first floor(n/64)*64 elements are clamp(adds_epi8(a,b), +-120); the remaining n%64 elements
 * use the scalar LLR sum (avx512_impl.cpp:45-63). The two only differ on non-finite inputs. */
static void combine_avx512(int8_t* out, const int8_t* in, uint32_t n)
{
  uint32_t vec = (n / 64) * 64;
  for (uint32_t i = 0; i != vec; ++i) {
    int s  = sat8((int)out[i] + (int)in[i]);
    out[i] = (int8_t)(s > 120 ? 120 : (s < -120 ? -120 : s));
  }
  for (uint32_t i = vec; i != n; ++i) {
    out[i] = oracle_llr_add(in[i], out[i]); /* in0 + in1 == (in1 += in0) : rhs=in1 is tested first (operator+ :107-111) */
  }
}

static const double K0_FACTOR[2][4] = {{0, 17, 33, 56}, {0, 13, 25, 43}};

int oracle_dematch(int8_t* out, uint32_t N, const int8_t* in, uint32_t E, int new_data, int rv, int Qm, uint32_t Nref,
                   uint32_t F)
{
  int bgi;
  if (N % 66 == 0) {
    bgi = 0;
  } else if (N % 50 == 0) {
    bgi = 1;
  } else {
    return -1;
  }
  uint32_t Ns  = bgi ? 50 : 66;
  uint32_t Kb  = bgi ? 10 : 22;
  uint32_t Z   = N / Ns;
  uint32_t Ncb = Nref ? (Nref < N ? Nref : N) : N;
  uint32_t sys = (Kb - 2) * Z;
  if (F >= sys || (Qm > 1 && E % (uint32_t)Qm != 0)) {
    return -1;
  }
  uint32_t info = sys - F;
  uint32_t k0   = (uint32_t)((uint16_t)floor(K0_FACTOR[bgi][rv] * (double)Ncb / (double)N)) * Z;

  /* De-interleaving (ldpc_rate_dematcher_impl.cpp:203-257): out[(E/Qm) * j + i] = in[i * Qm + j]. */
  int8_t* d = (int8_t*)malloc(E ? E : 1);
  if (Qm > 1) {
    uint32_t S = E / (uint32_t)Qm;
    for (uint32_t i = 0; i != S; ++i) {
      for (uint32_t j = 0; j != (uint32_t)Qm; ++j) {
        d[S * j + i] = in[i * (uint32_t)Qm + j];
      }
    }
  } else {
    memcpy(d, in, E);
  }

  /* allot_llrs (:128-201). */
  int      copy = new_data;
  uint32_t idx = k0, pos = 0;
  while (pos != E) {
    if (idx < info) {
      uint32_t n = info - idx;
      if (n > E - pos) {
        n = E - pos;
      }
      if (copy) {
        memset(out, 0, idx);
        memcpy(out + idx, d + pos, n);
      } else {
        combine_avx512(out + idx, d + pos, n);
      }
      idx += n;
      pos += n;
    } else if (copy) {
      memset(out, 0, info);
    }
    if (copy) {
      memset(out + info, 127, F);
    }
    if (idx < sys) {
      idx = sys;
    }
    uint32_t n = Ncb - idx;
    if (n > E - pos) {
      n = E - pos;
    }
    if (copy) {
      memcpy(out + idx, d + pos, n);
    } else {
      combine_avx512(out + idx, d + pos, n);
    }
    idx = (idx + n) % Ncb;
    pos += n;
    if (pos != E) {
      copy = 0;
    }
  }
  if (copy && idx != 0) {
    /* NB: the tail of the N-byte span, not of the Ncb circular buffer (:197-200). */
    memset(out + N - (Ncb - idx), 0, Ncb - idx);
  }
  free(d);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * LDPC decoding (ldpc_decoder_impl.cpp:60-308, ldpc_decoder_avx512.cpp:81-290, avx512_support.h:65-107).
 * ---------------------------------------------------------------------------------------------------------------- */
static int ls_index(int Z)
{
  /* TS 38.212 Table 5.3.2-1: Z = a * 2^j, a in {2,3,5,7,9,11,13,15} -> set index 0..7 (ldpc_luts_impl.cpp:57). */
  static const int a[8] = {2, 3, 5, 7, 9, 11, 13, 15};
  for (int i = 0; i != 8; ++i) {
    for (int z = a[i]; z <= 384; z *= 2) {
      if (z == Z) {
        return i;
      }
    }
  }
  return -1;
}

/* mm512::scale_epi8 on one byte (avx512_support.h:65-107): bytes within +-max are multiplied as UNSIGNED bytes by
 * (uint16)(sf * 65536) and shifted right by 16; other bytes pass through. */
static int8_t scale_byte(int8_t a, float sf)
{
  if (sf >= .9999) {
    return a;
  }
  if (a > 120 || a < -120) {
    return a;
  }
  uint32_t m = (uint16_t)(sf * 65536.0F);
  return (int8_t)(uint8_t)((((uint32_t)(uint8_t)a) * m) >> 16);
}

int oracle_ldpc_decode(uint8_t* out_packed, const int8_t* in, uint32_t n_in, int bg, int Z, uint32_t F, int crc_poly,
                       int max_it, float scaling, uint32_t* nof_layers_out)
{
  const uint16_t* row_ptr = (bg == 1) ? NR_BG1_ROW_PTR : NR_BG2_ROW_PTR;
  const uint8_t*  col     = (bg == 1) ? NR_BG1_COL : NR_BG2_COL;
  int             ils     = ls_index(Z);
  if (ils < 0) {
    return -3;
  }
  const uint16_t* shift_tab = (bg == 1) ? NR_BG1_SHIFT[ils] : NR_BG2_SHIFT[ils];
  uint32_t        Kb        = (bg == 1) ? 22 : 10;
  uint32_t        Ns        = (bg == 1) ? 66 : 50;
  uint32_t        nof_edges = (bg == 1) ? NR_BG1_NOF_EDGES : NR_BG2_NOF_EDGES;
  uint32_t        K         = Kb * (uint32_t)Z;
  uint32_t        zz        = (uint32_t)Z;
  if (n_in > Ns * zz || n_in < K + 2 * zz) {
    return -3;
  }
  if (nof_layers_out) {
    *nof_layers_out = 0;
  }

  /* Trim trailing zeros (:86-99). */
  uint32_t last = n_in;
  while (last > 0 && in[last - 1] == 0) {
    --last;
  }
  if (last == 0) {
    if (crc_poly == ORACLE_CRC_NONE) {
      for (uint32_t i = 0; i != K; ++i) {
        out_packed[i >> 3] |= (uint8_t)(0x80U >> (i & 7));
      }
    }
    return -1;
  }
  uint32_t cbl = last + 2 * zz;
  if (cbl < K + 4 * zz) {
    cbl = K + 4 * zz;
  }
  if (cbl % zz != 0) {
    cbl = (cbl / zz + 1) * zz;
  }
  uint32_t L = cbl / zz - Kb;
  if (nof_layers_out) {
    *nof_layers_out = L;
  }

  /* load_soft_bits (:149-174): two punctured nodes, then the input. Everything past the input is "whatever the decoder
   * object held before" in the reference; those nodes are only reachable for layers beyond the trimmed length, which are
   * never processed, except for the tail of the last partially filled node, which the input covers with explicit zeros
   * (the trimmed LLRs are zeros and are still copied). */
  int8_t* soft = (int8_t*)calloc((Ns + 2) * zz, 1);
  memcpy(soft + 2 * zz, in, n_in);

  int8_t*  c2v      = (int8_t*)calloc((size_t)nof_edges * zz, 1); /* per edge, indexed by variable position k */
  int8_t*  v2c      = (int8_t*)malloc(19 * (size_t)zz);
  uint8_t* row_init = (uint8_t*)calloc(64, 1);
  int      result   = -1;

  for (int it = 0; it != max_it; ++it) {
    for (uint32_t l = 0; l != L; ++l) {
      uint32_t e0 = row_ptr[l], deg = row_ptr[l + 1] - e0;
      /* Variable-to-check (:176-219 + avx512.cpp:81-121). */
      for (uint32_t e = 0; e != deg; ++e) {
        const int8_t* s = soft + (uint32_t)col[e0 + e] * zz;
        const int8_t* c = c2v + (size_t)(e0 + e) * zz;
        for (uint32_t k = 0; k != zz; ++k) {
          int8_t v;
          if (!row_init[l]) {
            v = s[k];
          } else {
            int dlt = sat8((int)s[k] - (int)c[k]);
            dlt     = dlt > 120 ? 120 : dlt;
            dlt     = dlt < -120 ? -120 : dlt;
            if (!(127 > s[k])) {
              dlt = 127;
            }
            if (!(s[k] > -127)) {
              dlt = -127;
            }
            v = (int8_t)dlt;
          }
          v2c[e * zz + k] = v;
        }
      }
      /* Check-to-variable (:236-308 + avx512.cpp:123-216). */
      for (uint32_t j = 0; j != zz; ++j) {
        int8_t  m1 = 120, m2 = 120;
        uint8_t sg = 0, amin = 0;
        for (uint32_t e = 0; e != deg; ++e) {
          uint32_t sh  = shift_tab[e0 + e] % zz;
          int8_t   v   = v2c[e * zz + (j + sh) % zz];
          int8_t   a   = (int8_t)(v < 0 ? -v : v); /* abs_epi8: -128 stays -128 */
          int      lt1 = m1 > a;
          int8_t   hlp = lt1 ? m1 : a;
          sg ^= (uint8_t)v;
          if (lt1) {
            m1   = a;
            amin = (uint8_t)e;
          }
          if (m2 > a) {
            m2 = hlp;
          }
        }
        for (uint32_t e = 0; e != deg; ++e) {
          uint32_t sh  = shift_tab[e0 + e] % zz;
          uint32_t k   = (j + sh) % zz;
          int8_t   v   = v2c[e * zz + k];
          int8_t   mag = scale_byte((e == amin) ? m2 : m1, scaling);
          int8_t   fs  = (int8_t)((uint8_t)v ^ sg);
          c2v[(size_t)(e0 + e) * zz + k] = (fs >= 0) ? mag : (int8_t)(uint8_t)(((uint8_t)mag ^ 0xFFU) + 1U);
        }
      }
      /* Soft bits (:221-234 + avx512.cpp:218-259). */
      for (uint32_t e = 0; e != deg; ++e) {
        int8_t*       s = soft + (uint32_t)col[e0 + e] * zz;
        const int8_t* c = c2v + (size_t)(e0 + e) * zz;
        for (uint32_t k = 0; k != zz; ++k) {
          int8_t cc = c[k], vv = v2c[e * zz + k];
          int    cp = cc > 120, cm = cc < -120, vp = vv > 120, vm = vv < -120;
          int    sum = sat8((int)cc + (int)vv);
          if ((sum > 120) || (cp && !vm) || (vp && !cm)) {
            sum = 127;
          }
          if ((sum < -120) || (cm && !vp) || (vm && !cp)) {
            sum = -127;
          }
          s[k] = (int8_t)sum;
        }
      }
      row_init[l] = 1;
    }
    if (crc_poly != ORACLE_CRC_NONE) {
      int ok = oracle_hard_decision(out_packed, soft, K);
      if (ok && oracle_crc(crc_poly, out_packed, K - F) == 0) {
        result = it + 1;
        break;
      }
    }
  }
  if (crc_poly == ORACLE_CRC_NONE) {
    oracle_hard_decision(out_packed, soft, K);
  }
  free(soft);
  free(c2v);
  free(v2c);
  free(row_init);
  return result;
}

/* pusch_codeblock_decoder::decode (pusch_codeblock_decoder.cpp:35-71). */
int oracle_cb_decode(uint8_t* cb_data, int8_t* softbuf, uint32_t N, const int8_t* llrs, uint32_t E, int new_data,
                     int rv, int Qm, uint32_t Nref, uint32_t F, int bg, int Z, int crc_poly, int early_stop, int max_it)
{
  oracle_dematch(softbuf, N, llrs, E, new_data, rv, Qm, Nref, F);
  if (early_stop) {
    return oracle_ldpc_decode(cb_data, softbuf, N, bg, Z, F, crc_poly, max_it, 0.8F, NULL);
  }
  oracle_ldpc_decode(cb_data, softbuf, N, bg, Z, F, ORACLE_CRC_NONE, max_it, 0.8F, NULL);
  uint32_t K = (uint32_t)((bg == 1) ? 22 : 10) * (uint32_t)Z;
  return (oracle_crc(crc_poly, cb_data, K - F) == 0) ? max_it : -1;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Segmentation, Rx side (ldpc_segmenter_impl.cpp:58-68,254-331; ldpc.h:128-228).
 * ---------------------------------------------------------------------------------------------------------------- */
static const uint16_t ALL_Z[51] = {2,  3,  4,  5,  6,  7,  8,  9,   10,  11,  12,  13,  14,  15,  16,  18,  20,
                                   22, 24, 26, 28, 30, 32, 36, 40,  44,  48,  52,  56,  60,  64,  72,  80,  88,
                                   96, 104, 112, 120, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288, 320, 352, 384};

int oracle_segment_rx(uint32_t tbs, int bg, int Qm, int nof_layers, uint32_t nof_llrs, oracle_cb_meta* out)
{
  uint32_t tb_crc  = (tbs <= 3824) ? 16 : 24;
  uint32_t B       = tbs + tb_crc;
  uint32_t max_seg = (bg == 1) ? 8448 : 3840;
  uint32_t C       = (B <= max_seg) ? 1 : (B + (max_seg - 24) - 1) / (max_seg - 24);
  uint32_t Bp      = B + ((C > 1) ? 24 * C : 0);
  uint32_t ref_len = 22;
  if (bg == 2) {
    ref_len = (B > 640) ? 10 : (B > 560) ? 9 : (B > 192) ? 8 : 6;
  }
  uint32_t Z = 0;
  for (int i = 0; i != 51; ++i) {
    if (ALL_Z[i] * C * ref_len >= Bp) {
      Z = ALL_Z[i];
      break;
    }
  }
  if (Z == 0) {
    return -1;
  }
  uint32_t K        = ((bg == 1) ? 22 : 10) * Z;
  uint32_t cb_crc   = (C > 1) ? 24 : 0;
  uint32_t max_info = (Bp + C - 1) / C - cb_crc;
  uint32_t nsl      = (nof_llrs / (uint32_t)Qm) / (uint32_t)nof_layers; /* symbols per layer */
  uint32_t nshort   = C - (nsl % C);
  uint32_t off      = 0;
  for (uint32_t i = 0; i != C; ++i) {
    uint32_t per = (i < nshort) ? nsl / C : (nsl + C - 1) / C;
    uint32_t E   = per * (uint32_t)nof_layers * (uint32_t)Qm;
    out[i].bg              = (uint32_t)bg;
    out[i].Z               = Z;
    out[i].full_length     = K * ((bg == 1) ? 3 : 5);
    out[i].rm_length       = E;
    out[i].nof_filler_bits = K - (max_info + cb_crc);
    out[i].cw_offset       = off;
    out[i].nof_crc_bits    = (C == 1) ? tb_crc : cb_crc;
    off += E;
  }
  return (off == nof_llrs) ? (int)C : -1;
}

/* ------------------------------------------------------------------------------------------------------------------
 * TB-level decoder (pusch_decoder_impl.cpp:35-46 CRC selection, :309-382 per-CB task, :384-450 join, :452-497 TB assembly).
 * ---------------------------------------------------------------------------------------------------------------- */
#define DATA_STRIDE (ORACLE_MAX_CB_LEN / 8 + 8)

oracle_harq_buffer* oracle_harq_create(uint32_t nof_cbs)
{
  oracle_harq_buffer* b = (oracle_harq_buffer*)calloc(1, sizeof(*b));
  b->nof_cbs            = nof_cbs;
  b->soft               = (int8_t*)calloc((size_t)nof_cbs * ORACLE_MAX_CB_LEN, 1);
  b->data               = (uint8_t*)calloc((size_t)nof_cbs * DATA_STRIDE, 1);
  return b;
}

void oracle_harq_destroy(oracle_harq_buffer* b)
{
  if (b) {
    free(b->soft);
    free(b->data);
    free(b);
  }
}

static unsigned get_bit(const uint8_t* p, uint32_t i)
{
  return (p[i >> 3] >> (7 - (i & 7))) & 1U;
}

static void put_bit(uint8_t* p, uint32_t i, unsigned b)
{
  uint8_t mask = (uint8_t)(0x80U >> (i & 7));
  p[i >> 3]    = (uint8_t)(b ? (p[i >> 3] | mask) : (p[i >> 3] & ~mask));
}

int oracle_pusch_decode(oracle_harq_buffer* harq, uint8_t* tb_out, uint32_t tb_bytes, const int8_t* llrs,
                        uint32_t nof_llrs, int bg, int rv, int Qm, uint32_t Nref, int nof_layers, int max_it,
                        int early_stop, int new_data, oracle_tb_result* result)
{
  oracle_cb_meta meta[ORACLE_MAX_CB];
  uint32_t       tbs = tb_bytes * 8;
  int            C   = oracle_segment_rx(tbs, bg, Qm, nof_layers, nof_llrs, meta);
  if (C < 0 || (uint32_t)C != harq->nof_cbs) {
    return -1;
  }
  int crc_poly = (C > 1) ? ORACLE_CRC24B : ((tbs > 3824) ? ORACLE_CRC24A : ORACLE_CRC16);
  if (new_data) {
    memset(harq->crc, 0, sizeof(harq->crc));
  }
  uint32_t nobs = 0, imin = 0xffffffffU, imax = 0;
  double   isum = 0;
  for (int cb = 0; cb != C; ++cb) {
    int8_t*       soft = harq->soft + (size_t)cb * ORACLE_MAX_CB_LEN;
    uint8_t*      data = harq->data + (size_t)cb * DATA_STRIDE;
    const int8_t* in   = llrs + meta[cb].cw_offset;
    if (harq->crc[cb]) {
      oracle_dematch(soft, meta[cb].full_length, in, meta[cb].rm_length, new_data, rv, Qm, Nref,
                     meta[cb].nof_filler_bits);
      continue;
    }
    int it = oracle_cb_decode(data, soft, meta[cb].full_length, in, meta[cb].rm_length, new_data, rv, Qm, Nref,
                              meta[cb].nof_filler_bits, bg, (int)meta[cb].Z, crc_poly, early_stop, max_it);
    uint32_t obs;
    if (it >= 0) {
      harq->crc[cb] = 1;
      obs           = (uint32_t)it;
    } else {
      obs = (uint32_t)max_it;
    }
    ++nobs;
    isum += obs;
    imin = obs < imin ? obs : imin;
    imax = obs > imax ? obs : imax;
  }

  int tb_ok = 0, all_ok = 1;
  for (int cb = 0; cb != C; ++cb) {
    all_ok &= harq->crc[cb];
  }
  if (C == 1) {
    tb_ok = harq->crc[0];
    if (tb_ok) {
      memcpy(tb_out, harq->data, tb_bytes);
    }
  } else if (all_ok) {
    uint32_t off = 0, checksum = 0;
    for (int cb = 0; cb != C; ++cb) {
      const uint8_t* data     = harq->data + (size_t)cb * DATA_STRIDE;
      uint32_t       K        = meta[cb].full_length / ((bg == 1) ? 3 : 5);
      uint32_t       nof_data = K - meta[cb].nof_crc_bits - meta[cb].nof_filler_bits;
      uint32_t       n        = tbs - off < nof_data ? tbs - off : nof_data;
      for (uint32_t i = 0; i != n; ++i) {
        put_bit(tb_out, off + i, get_bit(data, i));
      }
      if (cb == C - 1) {
        for (uint32_t i = 0; i != 24; ++i) {
          checksum = (checksum << 1) | get_bit(data, n + i);
        }
      }
      off += n;
    }
    if (oracle_crc(ORACLE_CRC24A, tb_out, tbs) == checksum) {
      tb_ok = 1;
    } else {
      memset(harq->crc, 0, sizeof(harq->crc));
    }
  }
  result->tb_crc_ok        = tb_ok;
  result->nof_codeblocks   = (uint32_t)C;
  result->nof_observations = nobs;
  result->iter_min         = nobs ? imin : 0;
  result->iter_max         = nobs ? imax : 0;
  result->iter_mean        = nobs ? (float)(isum / nobs) : 0;
  return 0;
}

double oracle_pusch_bench(uint32_t tb_bytes, const int8_t* llrs, uint32_t nof_llrs, int bg, int Qm, uint32_t Nref,
                          int nof_layers, int max_it, int early_stop, int reps, int* nof_crc_ok)
{
  oracle_cb_meta meta[ORACLE_MAX_CB];
  int            C = oracle_segment_rx(tb_bytes * 8, bg, Qm, nof_layers, nof_llrs, meta);
  if (C < 0) {
    return -1;
  }
  oracle_harq_buffer* h  = oracle_harq_create((uint32_t)C);
  uint8_t*            tb = (uint8_t*)malloc(tb_bytes);
  oracle_tb_result    res;
  int                 ok = 0;
  struct timespec     t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int r = 0; r != reps; ++r) {
    oracle_pusch_decode(h, tb, tb_bytes, llrs, nof_llrs, bg, 0, Qm, Nref, nof_layers, max_it, early_stop, 1, &res);
    ok += res.tb_crc_ok;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (nof_crc_ok) {
    *nof_crc_ok = ok;
  }
  free(tb);
  oracle_harq_destroy(h);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
