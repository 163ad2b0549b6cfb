// PDSCH mirror of the decoding path: code-block CRC attachment, LDPC encoding, rate matching (bit
// selection + bit interleaving) on the device, behind hal::hw_accelerator_pdsch_enc
// (include/srsran/hal/phy/upper/channel_processors/hw_accelerator_pdsch_enc.h:36-97). Reference arithmetic reproduced bit
// for bit:
//   segmentation + CRCs  ldpc_segmenter_tx_impl.cpp (TB CRC from the caller, CRC24B per code block when C > 1, zero padding
//                        of the last segment, filler bits)
//   encoding             ldpc_encoder_impl.cpp / TS 38.212 5.3.2: systematic, the 4 core parity blocks through the
//                        dual-diagonal structure, one independent parity block per extension row; fillers encode as 0
//   rate matching        ldpc_rate_matcher_impl.cpp:37-147: circular buffer of Ncb = min(Nref, N) bits read from k0, filler
//                        bits skipped, then f[i Qm + j] = e[j E/Qm + i]
//
// One CTA per code block. The lifted variable nodes live in shared memory as ONE BYTE PER BIT ((N_b + 2) Z <= 26112
// bytes): a circulant shift is an index rotation, exactly as in the decoder, and every parity bit is an XOR of up to 19
// byte loads. The kernel is bound by its output (E bytes of unpacked + E/8 bytes of packed bits per code block written to
// HBM), the encoding itself is ~316 Z byte operations per code block.
#pragma once
#include "pusch_dec_kernels.cuh"

namespace pusch_dec {

/// One code-block encoding operation. 64 bytes.
struct enc_cb_desc {
  const uint8_t* src;        ///< CB mode: the code block's K' = K - F message bits, packed MSB first; TB mode: the packed TB
  uint8_t*       out_bits;   ///< rate-matched bits, one per byte (E of them); may be null
  uint8_t*       out_packed; ///< rate-matched bits packed MSB first (ceil(E / 8) bytes); may be null
  uint32_t       E;          ///< rate-matched length
  uint32_t       Ncb;        ///< circular buffer length: Nref ? min(Nref, N) : N
  uint32_t       k0;         ///< starting position (ldpc_rate_matcher_impl.cpp:89-90)
  uint32_t       nof_filler;
  uint32_t       src_bit0;   ///< TB mode: first bit of this code block in the stream TB | TB CRC | zero padding
  uint32_t       info_bits;  ///< bits taken from `src` (TB mode: without the code-block CRC; CB mode: K')
  uint32_t       tbs_bits;   ///< TB mode: payload bits of the transport block
  uint32_t       tb_crc;     ///< TB mode: checksum of the transport block (computed by the caller, as the reference's
                             ///< pdsch_encoder_hw_impl does: pdsch_encoder_hw_impl.cpp:249-261)
  uint16_t       Z;
  uint8_t        bg;
  uint8_t        ils;
  uint8_t        Qm;
  uint8_t        tb_mode;    ///< 0: code-block input, 1: TB input with `tb_crc` = checksum, 2: with `tb_crc` = CRC job index
  uint8_t        tb_crc_len; ///< 16 or 24
  uint8_t        cb_crc_len; ///< TB mode: 24 when the TB has several code blocks, else 0
};
static_assert(sizeof(enc_cb_desc) == 64, "enc_cb_desc must stay 64 bytes");

/// Structure of the first core parity column per (base graph, lifting-set index): it has three entries in rows 0..3, two
/// with equal shifts; summing the four rows leaves P^y p1 = sum of the lambdas. ent[r] = its shift in row r (-1: none).
struct enc_core_desc {
  int16_t y;
  int16_t ent[4];
};
__constant__ enc_core_desc c_enc_core[2][8];

#ifndef PDSCH_ENC_THREADS
#define PDSCH_ENC_THREADS 256
#endif
constexpr int ENC_THREADS = PDSCH_ENC_THREADS;
static_assert(ENC_THREADS % 32 == 0 && ENC_THREADS >= 192,
              "the core-row phase of the packed encoder gives four threads to each of the 4 * 12 (row, word) pairs of Z = 384");

/// Shared memory of the encoder for a code block of base graph `bg` and lifting size Z.
__host__ __device__ inline uint32_t enc_smem_bytes(uint32_t bg, uint32_t Z)
{
  const uint32_t nodes = (((bg == 1) ? 68U : 52U) * Z + 15U) & ~15U; // one byte per lifted variable node
  const uint32_t lam   = (4U * Z + 15U) & ~15U;                       // the four core-row sums
  const uint32_t tab   = MAX_EDGES * 8U;                              // (column offset, lifted shift) per edge
  const uint32_t words = 272U * 4U;                                   // packed message words (code-block CRC)
  return nodes + lam + tab + words + 4U * 1024U;                      // + CRC byte tables
}

__global__ void __launch_bounds__(ENC_THREADS) pdsch_encode_kernel(const enc_cb_desc* __restrict__ descs,
                                                                   const uint32_t* __restrict__ tb_crcs)
{
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const enc_cb_desc& d    = descs[blockIdx.x];
  if ((d.Z & 31U) == 0) {
    return; // handled by pdsch_encode_packed_kernel
  }
  const int          t    = threadIdx.x;
  const int          lane = t & 31;
  const uint32_t     Z = d.Z, bg = d.bg, Kb = (bg == 1) ? 22 : 10, Nb = (bg == 1) ? 68 : 52, K = Kb * Z;
  const uint32_t     nrows = Nb - Kb;
  uint8_t*           nodes = smem_raw;
  uint8_t*           lam   = nodes + ((Nb * Z + 15U) & ~15U);
  uint2*             tab   = reinterpret_cast<uint2*>(lam + ((4U * Z + 15U) & ~15U));
  uint32_t*          msgw  = reinterpret_cast<uint32_t*>(tab + MAX_EDGES);
  uint32_t*          tabs  = msgw + 272;

  // ---- rate-matching geometry: which part of the code word is read at all -----------------------------------------------------
  const uint32_t sys = (Kb - 2) * Z, F = d.nof_filler, Ncb = d.Ncb, E = d.E;
  const uint32_t fs = min(sys - F, Ncb), fe = min(sys, Ncb), Fc = fe - fs; // filler positions inside the circular buffer
  const uint32_t V  = Ncb - Fc;                                            // bits per lap
  const uint32_t v0 = (d.k0 < fs) ? d.k0 : ((d.k0 < fe) ? fs : d.k0 - Fc); // valid index of the first selected bit
  // Highest code-word position read: without a wrap only [k0, k0 + E + Fc) is needed - parity rows beyond it are skipped.
  const uint32_t top   = (v0 + E <= V) ? min(Ncb, ((v0 + E <= fs) ? v0 + E : v0 + E + Fc)) : Ncb;
  const uint32_t nodes_needed = (top + 2 * Z + Z - 1) / Z; // variable-node columns incl. the two punctured ones
  const uint32_t rows_needed  = (nodes_needed > Kb) ? min(nrows, nodes_needed - Kb) : 0U;
  const uint32_t rows_done    = max(rows_needed, 4U);

  const uint32_t nedges = c_row_ptr[bg - 1][rows_done];
  for (uint32_t e = t; e < nedges; e += ENC_THREADS) {
    tab[e] = make_uint2((uint32_t)c_col[bg - 1][e] * Z, c_shift[bg - 1][d.ils][e] % Z);
  }
  if (d.tb_mode && d.cb_crc_len != 0) {
    build_crc_tables(tabs, 2, t, ENC_THREADS);
  }

  // ---- message bits: one byte per bit; fillers and everything beyond K' are zero -----------------------------------------------
  const uint8_t* __restrict__ src = d.src;
  const uint32_t info = d.info_bits;
  if (d.tb_mode) {
    // tb_mode 2: the checksum was computed on the device by the CRC job of that index (crc_kernel, launched in front)
    const uint32_t tbs = d.tbs_bits, crc_len = d.tb_crc_len, tb_crc = (d.tb_mode == 2) ? tb_crcs[d.tb_crc] : d.tb_crc;
    for (uint32_t i = t; i < K; i += ENC_THREADS) {
      uint32_t bit = 0;
      if (i < info) {
        const uint32_t s = d.src_bit0 + i;
        if (s < tbs) {
          bit = (__ldg(src + (s >> 3)) >> (7 - (s & 7))) & 1U;
        } else if (s < tbs + crc_len) {
          bit = (tb_crc >> (crc_len - 1 - (s - tbs))) & 1U;
        }
      }
      nodes[i] = (uint8_t)bit;
    }
    if (d.cb_crc_len != 0) {
      // CRC24B of the info bits (ldpc_segmenter_tx_impl.cpp: every segment of a multi-segment TB): pack them MSB first,
      // one warp folds the words, the 24 checksum bits follow the info bits.
      __syncthreads();
      const uint32_t nw = (info + 31) / 32;
      for (uint32_t w = t >> 5; w < nw; w += ENC_THREADS / 32) {
        const uint32_t i   = w * 32 + 31 - lane; // ballot bit l <-> bit 31 - l of the word: MSB first
        const uint32_t b   = (i < info) ? nodes[i] : 0U;
        const uint32_t wd  = __ballot_sync(0xffffffffU, b != 0);
        if (lane == 0) {
          msgw[w] = wd;
        }
      }
      __syncthreads();
      if (t < 32) {
        const uint32_t crc = warp_crc_words<false>(msgw, info, 2, tabs, lane);
        if (lane < 24) {
          nodes[info + lane] = (uint8_t)((crc >> (23 - lane)) & 1U);
        }
      }
    }
  } else {
    for (uint32_t i = t; i < K; i += ENC_THREADS) {
      nodes[i] = (i < info) ? (uint8_t)((__ldg(src + (i >> 3)) >> (7 - (i & 7))) & 1U) : (uint8_t)0;
    }
  }
  __syncthreads();

  auto rot_xor = [&](uint32_t e, uint32_t j) -> uint32_t {
    const uint2 te = tab[e];
    uint32_t    k  = j + te.y;
    k              = (k >= Z) ? k - Z : k;
    return nodes[te.x + k];
  };

  // ---- core rows: lambda_r = sum over the systematic columns ---------------------------------------------------------------------
  for (uint32_t i = t; i < 4 * Z; i += ENC_THREADS) {
    const uint32_t r = i / Z, j = i - r * Z;
    uint32_t       acc = 0;
    for (uint32_t e = c_row_ptr[bg - 1][r], e1 = c_row_ptr[bg - 1][r + 1]; e != e1; ++e) {
      if (tab[e].x < K) {
        acc ^= rot_xor(e, j);
      }
    }
    lam[i] = (uint8_t)acc;
  }
  __syncthreads();
  const enc_core_desc core = c_enc_core[bg - 1][d.ils];
  const uint32_t      y    = (uint32_t)core.y % Z;
  // p1 = P^-y (lambda_0 + lambda_1 + lambda_2 + lambda_3)
  for (uint32_t j = t; j < Z; j += ENC_THREADS) {
    uint32_t k = j + Z - y;
    k          = (k >= Z) ? k - Z : k;
    nodes[K + j] = lam[k] ^ lam[Z + k] ^ lam[2 * Z + k] ^ lam[3 * Z + k];
  }
  __syncthreads();
  // Dual diagonal: row r (0..2) introduces column K_b + 1 + r with shift 0.
  for (uint32_t j = t; j < Z; j += ENC_THREADS) {
    uint32_t prev = 0;
#pragma unroll
    for (uint32_t r = 0; r != 3; ++r) {
      uint32_t acc = lam[r * Z + j] ^ prev;
      if (core.ent[r] >= 0) {
        uint32_t k = j + (uint32_t)core.ent[r] % Z;
        k          = (k >= Z) ? k - Z : k;
        acc ^= nodes[K + k];
      }
      nodes[K + (1 + r) * Z + j] = (uint8_t)acc;
      prev                       = acc;
    }
  }
  __syncthreads();

  // ---- extension rows: one new column each, independent of one another ---------------------------------------------------------
  if (rows_needed > 4) {
    const uint32_t ntask = (rows_needed - 4) * Z;
    for (uint32_t i = t; i < ntask; i += ENC_THREADS) {
      const uint32_t r = 4 + i / Z, j = i - (r - 4) * Z;
      uint32_t       acc = 0;
      // (the last edge of an extension row is its own new column)
      for (uint32_t e = c_row_ptr[bg - 1][r], e1 = c_row_ptr[bg - 1][r + 1] - 1; e != e1; ++e) {
        acc ^= rot_xor(e, j);
      }
      nodes[(Kb + r) * Z + j] = (uint8_t)acc;
    }
  }
  __syncthreads();

  // ---- rate matching: bit selection from k0 (fillers skipped, wrap at Ncb) + interleaving, eight output bits per thread ------------
  const uint32_t Qm = d.Qm, EQ = E / Qm;
  const uint8_t* cw = nodes + 2 * Z; // the two punctured columns are not part of the code word
  auto out_bit = [&](uint32_t n) -> uint32_t {
    // f[i Qm + j] = e[j E/Qm + i] (ldpc_rate_matcher_impl.cpp:149-270)
    const uint32_t i = n / Qm, jq = n - i * Qm;
    uint32_t       u = v0 + jq * EQ + i;
    u                = (u >= V) ? u % V : u;
    return cw[(u < fs) ? u : u + Fc];
  };
  uint8_t* __restrict__ ob = d.out_bits;
  uint8_t* __restrict__ op = d.out_packed;
  for (uint32_t b = t; b * 8 < E; b += ENC_THREADS) {
    const uint32_t n0 = b * 8, cnt = min(8U, E - n0);
    uint32_t       lo = 0, hi = 0, byte = 0;
#pragma unroll
    for (uint32_t k = 0; k != 8; ++k) {
      const uint32_t bit = (k < cnt) ? out_bit(n0 + k) : 0U;
      byte |= bit << (7 - k);
      if (k < 4) {
        lo |= bit << (8 * k);
      } else {
        hi |= bit << (8 * (k - 4));
      }
    }
    if (op != nullptr) {
      op[b] = (uint8_t)byte;
    }
    if (ob != nullptr) {
      if (cnt == 8 && ((reinterpret_cast<uintptr_t>(ob) & 7U) == 0)) {
        *reinterpret_cast<uint2*>(ob + n0) = make_uint2(lo, hi);
      } else {
        for (uint32_t k = 0; k != cnt; ++k) {
          ob[n0 + k] = (uint8_t)((byte >> (7 - k)) & 1U);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Packed form for Z % 32 == 0 (every transport block large enough to matter: Z = 384 for full-size code blocks): a lifted
// variable node is Z / 32 words of 32 bits (bit j of the column = bit j % 32 of word j / 32), a circulant shift is a word
// rotation plus one funnel shift, a parity word is the XOR of up to 19 of them. The encoding proper then costs ~4
// instructions per (edge, word) - 15 k thread instructions per full-size code block instead of 1.8 M with one byte per
// bit - and the kernel is bound by what it has to write: E bytes of unpacked and E / 8 bytes of packed bits.
// ---------------------------------------------------------------------------------------------------------------------
/// TWO_LAPS: the selection covers at most two laps of the circular buffer (every practical code rate), so the wrap is one
/// conditional subtraction.
template <int QM, bool TWO_LAPS>
__device__ __forceinline__ void enc_rate_match_packed(const uint32_t* __restrict__ cw, uint32_t v0, uint32_t V, uint32_t fs,
                                                      uint32_t Fc, uint32_t E, uint8_t* __restrict__ ob, uint8_t* __restrict__ op,
                                                      int t)
{
  const uint32_t EQ = E / QM;
  const bool     aligned = (reinterpret_cast<uintptr_t>(ob) & 7U) == 0;
  // Code-word bit behind output bit n: f[i QM + j] = e[j E/QM + i]; e[m] = the m-th non-filler position from k0, modulo the
  // buffer (ldpc_rate_matcher_impl.cpp:101-147, 149-270).
  auto src_bit = [&](uint32_t n) -> uint32_t {
    const uint32_t i = n / QM, jq = n - i * QM;
    uint32_t       u = v0 + jq * EQ + i;
    if constexpr (TWO_LAPS) {
      u = (u >= V) ? u - V : u;
    } else {
      u = (u >= V) ? u % V : u;
    }
    const uint32_t p = (u < fs) ? u : u + Fc;
    return cw[p >> 5] >> (p & 31U);
  };
  for (uint32_t b = t; b * 8 < E; b += ENC_THREADS) {
    const uint32_t n0 = b * 8;
    if (n0 + 8 <= E) {
      // Eight bits: shifted in from the top, so that bit k of the group ends in bit 24 + k.
      uint32_t acc = 0;
#pragma unroll
      for (uint32_t k = 0; k != 8; ++k) {
        acc = __funnelshift_r(acc, src_bit(n0 + k), 1);
      }
      const uint32_t r = acc >> 24; // bit k = output bit n0 + k
      if (op != nullptr) {
        op[b] = (uint8_t)(__brev(r) >> 24);
      }
      if (ob != nullptr) {
        // one byte per bit: bits 0..3 of r into the four bytes of a word by one multiplication (distinct powers, no carries)
        const uint32_t lo = ((r & 0xfU) * 0x00204081U) & 0x01010101U, hi = ((r >> 4) * 0x00204081U) & 0x01010101U;
        if (aligned) {
          *reinterpret_cast<uint2*>(ob + n0) = make_uint2(lo, hi);
        } else {
          for (uint32_t k = 0; k != 8; ++k) {
            ob[n0 + k] = (uint8_t)((r >> k) & 1U);
          }
        }
      }
    } else {
      uint32_t byte = 0;
      for (uint32_t k = 0; n0 + k < E; ++k) {
        const uint32_t bit = src_bit(n0 + k) & 1U;
        byte |= bit << (7 - k);
        if (ob != nullptr) {
          ob[n0 + k] = (uint8_t)bit;
        }
      }
      if (op != nullptr) {
        op[b] = (uint8_t)byte;
      }
    }
  }
}

/// Rate matching 32 output bits per thread for QM in {2, 4, 8} (QM divides 32): the 32 / QM symbols of a thread take NS =
/// 32 / QM CONSECUTIVE bits from each of the QM streams e[j E/QM + i], i.e. one funnel-shifted word load per stream instead
/// of one shared-memory bit extraction per output bit; the bits are then interleaved in registers. Runs that cross the end of
/// the circular buffer or the filler gap (a handful per code block) are gathered bit by bit. Output: one packed word
/// (four bytes, MSB first) and 32 unpacked bytes (two 16-byte stores) per thread, contiguous across the warp.
template <int QM>
__device__ __forceinline__ void enc_rate_match_words(const uint32_t* __restrict__ cw, uint32_t v0, uint32_t V, uint32_t fs,
                                                     uint32_t Fc, uint32_t E, uint8_t* __restrict__ ob, uint8_t* __restrict__ op,
                                                     int t)
{
  static_assert(QM == 2 || QM == 4 || QM == 8, "QM must divide 32");
  constexpr uint32_t NS = 32 / QM;
  const uint32_t     EQ = E / QM;
  const uint32_t     nfull = E / 32; // whole 32-bit groups; the tail (< 32 bits) is done bit by bit below
  auto bit_at = [&](uint32_t m) -> uint32_t { // e[m]
    uint32_t u = v0 + m;
    u          = (u >= V) ? u % V : u;
    const uint32_t p = (u < fs) ? u : u + Fc;
    return (cw[p >> 5] >> (p & 31U)) & 1U;
  };
  for (uint32_t g = t; g < nfull; g += ENC_THREADS) {
    const uint32_t i0 = g * NS; // first symbol of the group
    uint32_t       chunk[QM];   // NS bits of stream j, bit s = symbol i0 + s
#pragma unroll
    for (int j = 0; j != QM; ++j) {
      const uint32_t m0 = (uint32_t)j * EQ + i0;
      uint32_t       u  = v0 + m0;
      u                 = (u >= V) ? u - V : u;
      // fast: at most one lap behind, the NS positions neither wrap nor straddle the filler gap
      if (u < V && u + NS <= V && (u >= fs || u + NS <= fs)) {
        const uint32_t p = (u < fs) ? u : u + Fc;
        chunk[j]         = __funnelshift_r(cw[p >> 5], cw[(p >> 5) + 1], p & 31U);
      } else {
        uint32_t c = 0;
        for (uint32_t s2 = 0; s2 != NS; ++s2) {
          c |= bit_at(m0 + s2) << s2;
        }
        chunk[j] = c;
      }
    }
    // Output bit n = 32 g + s QM + j  <-  chunk[j] bit s. Packed MSB first: bit n of the group at position 31 - (s QM + j).
    uint32_t word = 0;
#pragma unroll
    for (int s2 = 0; s2 != (int)NS; ++s2) {
#pragma unroll
      for (int j = 0; j != QM; ++j) {
        word |= ((chunk[j] >> s2) & 1U) << (31 - (s2 * QM + j));
      }
    }
    if (op != nullptr) {
      // bytes in memory order: the first eight bits of the group are the first byte
      reinterpret_cast<uint32_t*>(op)[g] = __byte_perm(word, 0, 0x0123);
    }
    if (ob != nullptr) {
      uint32_t       w8[8];
      const uint32_t rw = __brev(word); // bit n of the group at position n
#pragma unroll
      for (int k = 0; k != 8; ++k) {
        // output bits 4k .. 4k + 3 as one byte each, by one multiplication (distinct powers, no carries)
        w8[k] = (((rw >> (4 * k)) & 0xfU) * 0x00204081U) & 0x01010101U;
      }
      uint4* dst = reinterpret_cast<uint4*>(ob + (size_t)g * 32);
      dst[0]     = make_uint4(w8[0], w8[1], w8[2], w8[3]);
      dst[1]     = make_uint4(w8[4], w8[5], w8[6], w8[7]);
    }
  }
  // Tail: the last E % 32 bits.
  for (uint32_t n = nfull * 32 + t; n < E; n += ENC_THREADS) {
    const uint32_t i = n / QM, j = n - i * QM;
    const uint32_t bit = bit_at(j * EQ + i);
    if (ob != nullptr) {
      ob[n] = (uint8_t)bit;
    }
  }
  if (op != nullptr) {
    for (uint32_t b = nfull * 4 + t; b * 8 < E; b += ENC_THREADS) {
      uint32_t byte = 0;
      for (uint32_t k = 0; k != 8 && b * 8 + k < E; ++k) {
        const uint32_t n = b * 8 + k, i = n / QM, j = n - i * QM;
        byte |= bit_at(j * EQ + i) << (7 - k);
      }
      op[b] = (uint8_t)byte;
    }
  }
}

/// Which descriptors a launch handles: the packed kernel those with Z % 32 == 0, the byte kernel the others.
__device__ __forceinline__ bool enc_is_packed(const enc_cb_desc& d)
{
  return (d.Z & 31U) == 0;
}

__global__ void __launch_bounds__(ENC_THREADS) pdsch_encode_packed_kernel(const enc_cb_desc* __restrict__ descs,
                                                                          const uint32_t* __restrict__ tb_crcs)
{
  // 68 columns x 12 words, the four core-row sums, the edge table, the CRC byte tables
  __shared__ uint32_t colw[68 * 12 + 4]; // (+ one word: the rate matcher reads word pairs)
  __shared__ uint32_t lamw[4 * 12];
  __shared__ uint2    tab[MAX_EDGES];
  __shared__ uint32_t tabs[1024];
  __shared__ uint32_t msgw[272];
  const enc_cb_desc&  d = descs[blockIdx.x];
  if (!enc_is_packed(d)) {
    return;
  }
  const int      t    = threadIdx.x;
  const int      lane = t & 31;
  const uint32_t Z = d.Z, W = Z / 32, bg = d.bg, Kb = (bg == 1) ? 22 : 10, Nb = (bg == 1) ? 68 : 52, K = Kb * Z;
  const uint32_t nrows = Nb - Kb;

  const uint32_t sys = (Kb - 2) * Z, F = d.nof_filler, Ncb = d.Ncb, E = d.E;
  const uint32_t fs = min(sys - F, Ncb), fe = min(sys, Ncb), Fc = fe - fs;
  const uint32_t V  = Ncb - Fc;
  const uint32_t v0 = (d.k0 < fs) ? d.k0 : ((d.k0 < fe) ? fs : d.k0 - Fc);
  const uint32_t top = (v0 + E <= V) ? min(Ncb, ((v0 + E <= fs) ? v0 + E : v0 + E + Fc)) : Ncb;
  const uint32_t nodes_needed = (top + 2 * Z + Z - 1) / Z;
  const uint32_t rows_needed  = (nodes_needed > Kb) ? min(nrows, nodes_needed - Kb) : 0U;
  const uint32_t rows_done    = max(rows_needed, 4U);

  const uint32_t nedges = c_row_ptr[bg - 1][rows_done];
  for (uint32_t e = t; e < nedges; e += ENC_THREADS) {
    const uint32_t sh = c_shift[bg - 1][d.ils][e] % Z;
    tab[e]            = make_uint2((uint32_t)c_col[bg - 1][e] * W, sh); // first word of the column, lifted shift
  }
  if (d.tb_mode && d.cb_crc_len != 0) {
    build_crc_tables(tabs, 2, t, ENC_THREADS);
  }

  // ---- message words: bit i of the code block = bit i % 32 of word i / 32 -----------------------------------------------------
  const uint8_t* __restrict__ src = d.src;
  const uint32_t info = d.info_bits;
  const uint32_t tbs = d.tbs_bits, crc_len = d.tb_crc_len;
  const uint32_t tb_crc = (d.tb_mode == 2) ? tb_crcs[d.tb_crc] : d.tb_crc;
  for (uint32_t w = t; w < K / 32; w += ENC_THREADS) {
    const uint32_t i0 = w * 32;
    uint32_t       msb = 0; // the 32 bits MSB first (bit i0 in bit 31)
    if (i0 < info) {
      const uint32_t s0 = d.tb_mode ? d.src_bit0 + i0 : i0;
      if (!d.tb_mode || s0 + 32 <= tbs) {
        // Two aligned words (the staged inputs are 16-byte aligned and padded), big-endian, funnel-shifted to the bit.
        const uint32_t* __restrict__ w32 = reinterpret_cast<const uint32_t*>(src) + (s0 >> 5);
        const uint32_t a = __byte_perm(__ldg(w32), 0, 0x0123), b = __byte_perm(__ldg(w32 + 1), 0, 0x0123);
        msb              = __funnelshift_l(b, a, s0 & 31U);
      } else {
        for (uint32_t k = 0; k != 32; ++k) {
          const uint32_t s = s0 + k;
          uint32_t       bit = 0;
          if (s < tbs) {
            bit = (__ldg(src + (s >> 3)) >> (7 - (s & 7))) & 1U;
          } else if (s < tbs + crc_len) {
            bit = (tb_crc >> (crc_len - 1 - (s - tbs))) & 1U;
          }
          msb |= bit << (31 - k);
        }
      }
      if (i0 + 32 > info) {
        msb &= 0xffffffffU << (i0 + 32 - info); // only the first info - i0 bits
      }
    }
    msgw[w] = msb;
    colw[w] = __brev(msb);
  }
  __syncthreads();
  if (d.tb_mode && d.cb_crc_len != 0) {
    // CRC24B of the info bits, appended behind them (ldpc_segmenter_tx_impl.cpp).
    if (t < 32) {
      const uint32_t crc = warp_crc_words<false>(msgw, info, 2, tabs, lane);
      if (lane < 24) {
        const uint32_t i = info + lane;
        if ((crc >> (23 - lane)) & 1U) {
          atomicOr(&colw[i >> 5], 1U << (i & 31U));
        }
      }
    }
    __syncthreads();
  }

  // Word w of P^s x (x = column starting at word c0): bits [(32 w + s) mod Z, ...) of x.
  auto rot_word = [&](uint32_t c0, uint32_t s, uint32_t w) -> uint32_t {
    uint32_t a = w + (s >> 5);
    a          = (a >= W) ? a - W : a;
    uint32_t b = a + 1;
    b          = (b >= W) ? b - W : b;
    return __funnelshift_r(colw[c0 + a], colw[c0 + b], s & 31U);
  };

  // ---- core rows: four threads share the (up to 19) edges of a (row, word), combined by two shuffles -----------------------------
  {
    const uint32_t task = (uint32_t)t >> 2, q = (uint32_t)t & 3U;
    uint32_t       acc = 0;
    if (task < 4 * W) {
      const uint32_t r = task / W, w = task - r * W;
      for (uint32_t e = c_row_ptr[bg - 1][r] + q, e1 = c_row_ptr[bg - 1][r + 1]; e < e1; e += 4) {
        const uint2 te = tab[e];
        if (te.x < Kb * W) {
          acc ^= rot_word(te.x, te.y, w);
        }
      }
    }
    acc ^= __shfl_xor_sync(0xffffffffU, acc, 1);
    acc ^= __shfl_xor_sync(0xffffffffU, acc, 2);
    if (task < 4 * W && q == 0) {
      lamw[task] = acc;
    }
  }
  __syncthreads();
  const enc_core_desc core = c_enc_core[bg - 1][d.ils];
  uint32_t*           p1   = colw + Kb * W;
  if (t < (int)W) {
    // p1 = P^-y (sum of the lambdas): word t of the rotation by Z - y
    const uint32_t y = (uint32_t)core.y % Z, s = (y == 0) ? 0U : Z - y;
    uint32_t       a = t + (s >> 5);
    a                = (a >= W) ? a - W : a;
    uint32_t b       = a + 1;
    b                = (b >= W) ? b - W : b;
    const uint32_t ta = lamw[a] ^ lamw[W + a] ^ lamw[2 * W + a] ^ lamw[3 * W + a];
    const uint32_t tb = lamw[b] ^ lamw[W + b] ^ lamw[2 * W + b] ^ lamw[3 * W + b];
    p1[t]             = __funnelshift_r(ta, tb, s & 31U);
  }
  __syncthreads();
  if (t < (int)W) {
    uint32_t prev = 0;
#pragma unroll
    for (uint32_t r = 0; r != 3; ++r) {
      uint32_t acc = lamw[r * W + t] ^ prev;
      if (core.ent[r] >= 0) {
        acc ^= rot_word(Kb * W, (uint32_t)core.ent[r] % Z, t);
      }
      colw[(Kb + 1 + r) * W + t] = acc;
      prev                       = acc;
    }
  }
  __syncthreads();

  // ---- extension rows ------------------------------------------------------------------------------------------------------------
  if (rows_needed > 4) {
    // (same split: four threads per (row, word); the loop bound is uniform, every thread reaches the shuffles)
    const uint32_t ntask = (rows_needed - 4) * W, q = (uint32_t)t & 3U;
    for (uint32_t i0 = 0; i0 < ntask; i0 += ENC_THREADS / 4) {
      const uint32_t i = i0 + ((uint32_t)t >> 2);
      uint32_t       acc = 0, r = 0, w = 0;
      if (i < ntask) {
        r = 4 + i / W;
        w = i - (r - 4) * W;
        for (uint32_t e = c_row_ptr[bg - 1][r] + q, e1 = c_row_ptr[bg - 1][r + 1] - 1; e < e1; e += 4) {
          const uint2 te = tab[e];
          acc ^= rot_word(te.x, te.y, w);
        }
      }
      acc ^= __shfl_xor_sync(0xffffffffU, acc, 1);
      acc ^= __shfl_xor_sync(0xffffffffU, acc, 2);
      if (i < ntask && q == 0) {
        colw[(Kb + r) * W + w] = acc;
      }
    }
  }
  __syncthreads();

  // ---- rate matching ---------------------------------------------------------------------------------------------------------------
  const uint32_t* cw = colw + 2 * W;
  uint8_t* const  ob = d.out_bits;
  uint8_t* const  op = d.out_packed;
  // Word-granular form where the outputs allow vector stores (buffers of a batch are 8-byte aligned; 16 needed here).
  const bool      vec = (reinterpret_cast<uintptr_t>(ob) & 15U) == 0 && (reinterpret_cast<uintptr_t>(op) & 3U) == 0;
  if (vec && d.Qm == 8) {
    enc_rate_match_words<8>(cw, v0, V, fs, Fc, E, ob, op, t);
  } else if (vec && d.Qm == 4) {
    enc_rate_match_words<4>(cw, v0, V, fs, Fc, E, ob, op, t);
  } else if (vec && d.Qm == 2) {
    enc_rate_match_words<2>(cw, v0, V, fs, Fc, E, ob, op, t);
  } else if (v0 + E <= 2 * V) {
    switch (d.Qm) {
      case 8:
        enc_rate_match_packed<8, true>(cw, v0, V, fs, Fc, E, ob, op, t);
        break;
      case 6:
        enc_rate_match_packed<6, true>(cw, v0, V, fs, Fc, E, ob, op, t);
        break;
      case 4:
        enc_rate_match_packed<4, true>(cw, v0, V, fs, Fc, E, ob, op, t);
        break;
      case 2:
        enc_rate_match_packed<2, true>(cw, v0, V, fs, Fc, E, ob, op, t);
        break;
      default:
        enc_rate_match_packed<1, true>(cw, v0, V, fs, Fc, E, ob, op, t);
        break;
    }
  } else {
    switch (d.Qm) {
      case 8:
        enc_rate_match_packed<8, false>(cw, v0, V, fs, Fc, E, ob, op, t);
        break;
      case 6:
        enc_rate_match_packed<6, false>(cw, v0, V, fs, Fc, E, ob, op, t);
        break;
      case 4:
        enc_rate_match_packed<4, false>(cw, v0, V, fs, Fc, E, ob, op, t);
        break;
      case 2:
        enc_rate_match_packed<2, false>(cw, v0, V, fs, Fc, E, ob, op, t);
        break;
      default:
        enc_rate_match_packed<1, false>(cw, v0, V, fs, Fc, E, ob, op, t);
        break;
    }
  }
}

} // namespace pusch_dec
