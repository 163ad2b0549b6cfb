// TEST INFRASTRUCTURE. Runs the packed four-code-block LDPC arithmetic of the CUDA decoder
// (srsran_projectvtlmo_b200/csrc/ldpc_packed_math.h) on the CPU, one lifted check after the other, so that the exact
// same expressions the kernel executes are compared with the oracle without a GPU (tests/test_packed_math_cpu.py).
// Built by tests/host_emul/Makefile into tests/host_emul/libpacked_math_host.so.
#include "../../srsran_projectvtlmo_b200/csrc/ldpc_packed_math.h"
#include "../../srsran_projectvtlmo_b200/csrc/nr_ldpc_bg_tables.h"
#include <cstring>
#include <vector>

using namespace pusch_dec::pk;

namespace {

int ls_index(uint32_t Z)
{
  static const uint32_t a[8] = {2, 3, 5, 7, 9, 11, 13, 15};
  for (int i = 0; i != 8; ++i) {
    for (uint32_t z = a[i]; z <= 384; z *= 2) {
      if (z == Z) {
        return i;
      }
    }
  }
  return -1;
}

uint32_t crc_bits(const uint8_t* packed, uint32_t nbits, uint32_t gen, uint32_t order)
{
  uint32_t top = 1U << order, reg = 0;
  for (uint32_t i = 0; i != nbits; ++i) {
    uint32_t bit = (packed[i >> 3] >> (7 - (i & 7))) & 1U;
    reg          = (reg << 1) ^ (bit << order);
    if (reg & top) {
      reg ^= gen;
    }
  }
  return reg;
}

template <int DEG>
void run_check(uint32_t* soft, uint32_t* c2v_row, const uint32_t* tab_row, int j, int Z, uint32_t mult)
{
  check4<DEG> ck;
  int         addr[DEG];
  ck.begin();
  for (int e = 0; e != DEG; ++e) {
    uint32_t te = tab_row[e];
    int      k  = j + (int)(te >> 16);
    k           = (k >= Z) ? k - Z : k;
    addr[e]     = (int)(te & 0xffffU) + k;
    ck.gather(e, soft[2 * addr[e]], soft[2 * addr[e] + 1], c2v_row[e * Z + j]);
  }
  ck.reduce(mult);
  for (int e = 0; e != DEG; ++e) {
    uint32_t s0, s1;
    c2v_row[e * Z + j]    = ck.scatter(e, s0, s1);
    soft[2 * addr[e]]     = s0;
    soft[2 * addr[e] + 1] = s1;
  }
}

/// Same check with the two-code-blocks-per-thread variant (check2): two "threads" (r = 0, 1) share the lifted check.
template <int DEG>
void run_check_halves(uint32_t* soft, uint32_t* c2v_row, const uint32_t* tab_row, int j, int Z, uint32_t mult)
{
  for (int r = 0; r != 2; ++r) {
    check2<DEG> ck;
    int         addr[DEG];
    ck.begin();
    for (int e = 0; e != DEG; ++e) {
      uint32_t te = tab_row[e];
      int      k  = j + (int)(te >> 16);
      k           = (k >= Z) ? k - Z : k;
      addr[e]     = (int)(te & 0xffffU) + k;
      uint32_t cw = c2v_row[e * Z + j];
      uint32_t c16 = ((cw >> (8 * r)) & 0xffU) | (((cw >> (16 + 8 * r)) & 0xffU) << 8);
      ck.gather(e, soft[2 * addr[e] + r], c16);
    }
    ck.reduce(mult);
    for (int e = 0; e != DEG; ++e) {
      uint32_t s0;
      uint32_t c16 = ck.scatter(e, s0);
      uint32_t cw  = c2v_row[e * Z + j];
      cw &= ~((0xffU << (8 * r)) | (0xffU << (16 + 8 * r)));
      cw |= ((c16 & 0xffU) << (8 * r)) | ((c16 >> 8) << (16 + 8 * r));
      c2v_row[e * Z + j]    = cw;
      soft[2 * addr[e] + r] = s0;
    }
  }
}

/// Intra-code-block packing: "thread" j owns the lifted checks j, j + Z/4, j + Z/2, j + 3Z/4 of ONE code block.
template <int DEG>
void run_check_q4(uint32_t* soft, uint32_t* c2v_row, const uint32_t* tab_row, int j, int Z, uint32_t mult)
{
  const int   Z4 = Z / 4;
  check4<DEG> ck;
  int         addr[DEG];
  uint32_t    qs[DEG];
  ck.begin();
  for (int e = 0; e != DEG; ++e) {
    uint32_t te  = tab_row[e];
    int      col = (int)(te & 0xffffU) / Z;
    int      k   = (j + (int)(te >> 16)) % Z;
    qs[e]        = (uint32_t)(k / Z4);
    addr[e]      = col * Z4 + k % Z4;
    uint32_t r0 = soft[2 * addr[e]], r1 = soft[2 * addr[e] + 1];
    rot4(r0, r1, qs[e]);
    ck.gather(e, r0, r1, c2v_row[e * Z4 + j]);
  }
  ck.reduce(mult);
  for (int e = 0; e != DEG; ++e) {
    uint32_t s0, s1;
    c2v_row[e * Z4 + j] = ck.scatter(e, s0, s1);
    rot4(s0, s1, (4 - qs[e]) & 3U);
    soft[2 * addr[e]]     = s0;
    soft[2 * addr[e] + 1] = s1;
  }
}

} // namespace

extern "C" {

/// One code block decoded with the intra-code-block packing (Z % 4 == 0). Same conventions as pk_host_decode_group.
int pk_host_decode_q4(uint8_t* bits_out, const int8_t* llrs, uint32_t n_in, uint32_t bg, uint32_t Z, uint32_t nof_filler,
                      uint32_t crc_poly, uint32_t max_it, uint32_t mode, uint32_t mult, uint32_t nof_layers, int* iters_out)
{
  const uint16_t* row_ptr = (bg == 1) ? NR_BG1_ROW_PTR : NR_BG2_ROW_PTR;
  const uint8_t*  col     = (bg == 1) ? NR_BG1_COL : NR_BG2_COL;
  int             ils     = ls_index(Z);
  if (ils < 0 || Z % 4 != 0) {
    return -1;
  }
  const uint16_t* shift = (bg == 1) ? NR_BG1_SHIFT[ils] : NR_BG2_SHIFT[ils];
  const uint32_t  Kb = (bg == 1) ? 22 : 10, K = Kb * Z, Z4 = Z / 4;
  const uint32_t  ncols = Kb + nof_layers, nedges = row_ptr[nof_layers];
  const uint32_t  gen   = crc_poly == 1 ? 0x1864CFBU : (crc_poly == 2 ? 0x1800063U : 0x11021U);
  const uint32_t  order = crc_poly == 3 ? 16 : 24;
  std::vector<uint32_t> tab(nedges), soft(2 * (size_t)ncols * Z4), c2v((size_t)nedges * Z4, C2V_ZERO4);
  for (uint32_t e = 0; e != nedges; ++e) {
    tab[e] = (uint32_t)col[e] * Z | ((uint32_t)(shift[e] % Z) << 16);
  }
  bool allzero = true;
  auto llr_at  = [&](uint32_t c, uint32_t k) -> uint32_t {
    uint32_t i = c * Z + k;
    int8_t   x = (i >= 2 * Z && i - 2 * Z < n_in) ? llrs[i - 2 * Z] : 0;
    allzero    = allzero && (x == 0);
    return (uint32_t)(uint8_t)(x ^ 0x80);
  };
  for (uint32_t c = 0; c != ncols; ++c) {
    for (uint32_t b = 0; b != Z4; ++b) {
      soft[2 * (c * Z4 + b)]     = soft_from_biased_bytes(llr_at(c, b) | (llr_at(c, b + 2 * Z4) << 16));
      soft[2 * (c * Z4 + b) + 1] = soft_from_biased_bytes(llr_at(c, b + Z4) | (llr_at(c, b + 3 * Z4) << 16));
    }
  }
  *iters_out = -1;
  if (allzero && mode == 1) {
    return 0;
  }
  std::vector<uint8_t> hb((K + 7) / 8);
  for (uint32_t it = 0; it != max_it; ++it) {
    for (uint32_t l = 0; l != nof_layers; ++l) {
      uint32_t e0  = row_ptr[l];
      int      deg = row_ptr[l + 1] - e0;
      for (int j = 0; j != (int)Z4; ++j) {
        uint32_t* cr = c2v.data() + (size_t)e0 * Z4;
        switch (deg) {
#define CASE(D)                                                                                                        \
  case D:                                                                                                              \
    run_check_q4<D>(soft.data(), cr, tab.data() + e0, j, (int)Z, mult);                                                \
    break;
          CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(19)
#undef CASE
          default:
            return -2;
        }
      }
    }
    bool last_it = (it + 1 == max_it);
    if (mode == 1 || last_it) {
      std::fill(hb.begin(), hb.end(), 0);
      bool any_zero = false;
      for (uint32_t i = 0; i != K; ++i) {
        uint32_t c = i / Z, k = i % Z, m = k / Z4, b = k % Z4;
        uint32_t w    = soft[2 * (c * Z4 + b) + (m & 1)];
        uint32_t lane = (m & 2) ? (w >> 16) : (w & 0xffffU);
        if (lane_hard_bit(lane)) {
          hb[i >> 3] |= (uint8_t)(0x80U >> (i & 7));
        }
        any_zero |= (lane == BS);
      }
      std::memcpy(bits_out, hb.data(), K / 8);
      if (K % 8) {
        uint8_t mask     = (uint8_t)(0xff00U >> (K % 8));
        bits_out[K / 8] = (uint8_t)((bits_out[K / 8] & ~mask) | (hb[K / 8] & mask));
      }
      if ((crc_bits(hb.data(), K - nof_filler, gen, order) == 0) && (mode == 2 || !any_zero)) {
        *iters_out = (int)(mode == 1 ? it + 1 : max_it);
        break;
      }
    }
  }
  return 0;
}

static int g_lanes_per_thread = 4;
/// Selects the per-thread packing the emulation runs: 4 (check4) or 2 (check2).
void pk_host_set_lanes_per_thread(int n)
{
  g_lanes_per_thread = n;
}

/// Decodes up to four code blocks (lanes) packed together. llrs[c]: n_in[c] int8 LLRs (decoder input, natural order).
/// mode 1: early stop with CRC after every iteration; mode 2: CRC after max_it iterations. bits_out[c]: K/8 bytes,
/// written like the kernel does. iters_out[c]: iteration count or -1.
int pk_host_decode_group(uint8_t* const* bits_out, const int8_t* const* llrs, const uint32_t* n_in, uint32_t nof_lanes,
                         uint32_t bg, uint32_t Z, const uint32_t* nof_filler, uint32_t crc_poly, uint32_t max_it,
                         uint32_t mode, uint32_t mult, uint32_t nof_layers, int* iters_out)
{
  const uint16_t* row_ptr = (bg == 1) ? NR_BG1_ROW_PTR : NR_BG2_ROW_PTR;
  const uint8_t*  col     = (bg == 1) ? NR_BG1_COL : NR_BG2_COL;
  int             ils     = ls_index(Z);
  if (ils < 0 || nof_lanes == 0 || nof_lanes > 4) {
    return -1;
  }
  const uint16_t* shift = (bg == 1) ? NR_BG1_SHIFT[ils] : NR_BG2_SHIFT[ils];
  uint32_t        Kb = (bg == 1) ? 22 : 10, K = Kb * Z;
  uint32_t        nvar   = (Kb + nof_layers) * Z;
  uint32_t        nedges = row_ptr[nof_layers];
  uint32_t        gen    = crc_poly == 1 ? 0x1864CFBU : (crc_poly == 2 ? 0x1800063U : 0x11021U);
  uint32_t        order  = crc_poly == 3 ? 16 : 24;

  std::vector<uint32_t> tab(nedges), soft(2 * (size_t)nvar), c2v((size_t)nedges * Z, C2V_ZERO4);
  for (uint32_t e = 0; e != nedges; ++e) {
    tab[e] = (uint32_t)col[e] * Z | ((uint32_t)(shift[e] % Z) << 16);
  }
  bool allzero[4] = {true, true, true, true};
  for (uint32_t i = 0; i != nvar; ++i) {
    uint32_t ub[4];
    for (uint32_t c = 0; c != 4; ++c) {
      int8_t x = 0;
      if (c < nof_lanes && i >= 2 * Z && i - 2 * Z < n_in[c]) {
        x = llrs[c][i - 2 * Z];
      }
      if (x != 0) {
        allzero[c] = false;
      }
      ub[c] = (uint32_t)(uint8_t)(x ^ 0x80);
    }
    soft[2 * i]     = soft_from_biased_bytes(ub[0] | (ub[2] << 16));
    soft[2 * i + 1] = soft_from_biased_bytes(ub[1] | (ub[3] << 16));
  }
  bool done[4] = {false, false, false, false};
  for (uint32_t c = 0; c != 4; ++c) {
    iters_out[c] = -1;
    if (c >= nof_lanes) {
      done[c] = true;
    }
  }
  std::vector<uint8_t> hb((K + 7) / 8);
  auto hard = [&](uint32_t c, bool& any_zero) {
    std::fill(hb.begin(), hb.end(), 0);
    any_zero = false;
    for (uint32_t i = 0; i != K; ++i) {
      uint32_t w    = soft[2 * i + (c & 1)];
      uint32_t lane = (c & 2) ? (w >> 16) : (w & 0xffffU);
      if (lane_hard_bit(lane)) {
        hb[i >> 3] |= (uint8_t)(0x80U >> (i & 7));
      }
      any_zero |= (lane == BS);
    }
  };
  for (uint32_t it = 0; it != max_it; ++it) {
    for (uint32_t l = 0; l != nof_layers; ++l) {
      uint32_t e0  = row_ptr[l];
      int      deg = row_ptr[l + 1] - e0;
      for (int j = 0; j != (int)Z; ++j) {
        uint32_t* cr = c2v.data() + (size_t)e0 * Z;
        switch (deg) {
#define CASE(D)                                                                                                        \
  case D:                                                                                                              \
    if (g_lanes_per_thread == 4) {                                                                                     \
      run_check<D>(soft.data(), cr, tab.data() + e0, j, (int)Z, mult);                                                 \
    } else {                                                                                                           \
      run_check_halves<D>(soft.data(), cr, tab.data() + e0, j, (int)Z, mult);                                          \
    }                                                                                                                  \
    break;
          CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(19)
#undef CASE
          default:
            return -2;
        }
      }
    }
    bool last_it = (it + 1 == max_it);
    if (mode == 1 || last_it) {
      bool all_done = true;
      for (uint32_t c = 0; c != nof_lanes; ++c) {
        if (done[c]) {
          continue;
        }
        if (mode == 1 && allzero[c]) {
          continue; // the reference returns before touching the output
        }
        bool any_zero;
        hard(c, any_zero);
        std::memcpy(bits_out[c], hb.data(), K / 8);
        if (K % 8) {
          uint8_t mask        = (uint8_t)(0xff00U >> (K % 8));
          bits_out[c][K / 8] = (uint8_t)((bits_out[c][K / 8] & ~mask) | (hb[K / 8] & mask));
        }
        bool ok = (crc_bits(hb.data(), K - nof_filler[c], gen, order) == 0) && (mode == 2 || !any_zero);
        if (ok) {
          done[c]      = true;
          iters_out[c] = (int)(mode == 1 ? it + 1 : max_it);
        } else {
          all_done = false;
        }
      }
      if (all_done && mode == 1) {
        break;
      }
    }
  }
  return 0;
}
}
