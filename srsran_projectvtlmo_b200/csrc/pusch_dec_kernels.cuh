// Hand-written sm_100a kernels of the PUSCH channel-decoding path:
//   1. rate_dematch_kernel   - fused de-interleave + circular-buffer placement + HARQ soft combining, HBM-resident buffers
//   2. ldpc_decode_kernel    - batched layered normalized min-sum LDPC decoder with in-kernel CRC early stop
//   3. tb_assemble_crc_kernel / crc_kernel - transport-block assembly and CRC24A/B/16 by carry-less folding
//
// Arithmetic follows the reference's AVX2/AVX-512 flavour bit for bit (see DESIGN.md for the derivations):
//   rate dematcher : lib/phy/upper/channel_coding/ldpc/ldpc_rate_dematcher_impl.cpp:46-201,
//                    ldpc_rate_dematcher_avx512_impl.cpp:29-64
//   LDPC decoder   : lib/phy/upper/channel_coding/ldpc/ldpc_decoder_impl.cpp:60-308, ldpc_decoder_avx512.cpp:81-290,
//                    avx512_support.h:65-107
//   CRC            : lib/phy/upper/channel_coding/crc_calculator_lut_impl.cpp:33-38,66-152
#pragma once
#include "nr_ldpc_bg_tables.h"
#include "pusch_dec_types.h"
#include <cuda_runtime.h>

namespace pusch_dec {

// ---------------------------------------------------------------------------------------------------------------------
// Constant tables (filled once per device by upload_tables()).
// ---------------------------------------------------------------------------------------------------------------------
__constant__ uint16_t c_row_ptr[2][48];
__constant__ uint8_t  c_col[2][MAX_EDGES];
__constant__ uint16_t c_shift[2][8][MAX_EDGES];
/// x^(32 i) mod g, i = 0..271, for CRC24A, CRC24B, CRC16 (index poly - 1).
__constant__ uint32_t c_xpow32[3][272];
/// The same table in global memory, for lookups whose index differs from lane to lane (a constant-bank access with
/// divergent addresses is replayed once per distinct address).
__device__ uint32_t g_xpow32[3][272];
/// x^(128 i) mod g, i = 0..1023.
__constant__ uint32_t c_xpow128[3][1024];
/// x^(2^i) mod g, i = 0..31 (square-and-multiply ladders).
__constant__ uint32_t c_xpow2[3][32];
/// Byte tables of the three CRCs: g_crc_tabs[poly - 1][k * 256 + b] = (b(x) x^(8k) x^order) mod g, k = 0..3.
__device__ uint32_t g_crc_tabs[3][1024];

__host__ __device__ __forceinline__ uint32_t crc_gen(int poly)
{
  return poly == 1 ? 0x1864CFBU : (poly == 2 ? 0x1800063U : 0x11021U);
}
__host__ __device__ __forceinline__ uint32_t crc_order(int poly)
{
  return poly == 3 ? 16U : 24U;
}

/// (a(x) * b(x)) mod g over GF(2); a, b of degree < order.
__host__ __device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b, uint32_t gen, uint32_t order)
{
  uint32_t top = 1U << order;
  uint32_t acc = 0;
  for (int i = (int)order - 1; i >= 0; --i) {
    acc <<= 1;
    if (acc & top) {
      acc ^= gen;
    }
    if ((b >> i) & 1U) {
      acc ^= a;
    }
  }
  return acc;
}

/// gf2_mulmod with the degree of the generator known at compile time (fully unrolled, branch-free).
template <int ORDER>
__device__ __forceinline__ uint32_t gf2_mulmod_fixed(uint32_t a, uint32_t b, uint32_t gen)
{
  uint32_t acc = 0;
#pragma unroll
  for (int i = ORDER - 1; i >= 0; --i) {
    acc <<= 1;
    acc ^= gen & (0U - ((acc >> ORDER) & 1U));
    acc ^= a & (0U - ((b >> i) & 1U));
  }
  return acc;
}

/// Remainder update with `nbits` message bits taken from the top of `word` (MSB first): reg <- (reg * x^nbits + m * x^order) mod g.
__host__ __device__ __forceinline__ uint32_t crc_push_bits(uint32_t reg, uint32_t word, uint32_t nbits, uint32_t gen,
                                                           uint32_t order)
{
  uint32_t top = 1U << order;
  for (uint32_t i = 0; i != nbits; ++i) {
    uint32_t bit = (word >> (31 - i)) & 1U;
    reg          = (reg << 1) ^ (bit << order);
    if (reg & top) {
      reg ^= gen;
    }
  }
  return reg;
}

// ---------------------------------------------------------------------------------------------------------------------
// Kernel 1: rate dematching + HARQ combining.
// One CTA per code block. Stage 1 streams the E rate-matched LLRs from HBM with 16-byte loads and scatters them into
// shared memory in DE-INTERLEAVED order (d[j * E/Qm + i] = in[i * Qm + j], ldpc_rate_dematcher_impl.cpp:203-257), so
// that stage 2 reads rate-matched position p at smem[p]. Stage 2: every thread owns aligned 16-byte chunks of the
// soft buffer; the chunk is classified in O(1) against the break points of the reference's write set (info / filler /
// parity / wrap / end of data / zeroed head and tail). Uniform chunks (the bulk) take a vector path; chunks that
// contain a break point, multi-visit wraps (E > circular data length) and HARQ combining take the per-byte path, which
// derives, in closed form, the ordered list of rate-matched positions that land on a byte. The result is independent
// of scheduling and reproduces the reference's write set exactly (untouched bytes stay untouched).
// ---------------------------------------------------------------------------------------------------------------------
constexpr uint32_t DM_STAGE_CAP = 64 * 1024; ///< largest E staged through shared memory

struct dm_geom {
  uint32_t N, Ncb, E, F, info, sys, D, s0, first_len, S, Qm, new_data, zero_head, tail_start, copy_at_end, blk;
  uint32_t i_wrap, i_end; ///< break points: soft-buffer index of data slot s0, and of the first unvisited slot (E < D)
};

__device__ __forceinline__ int llr_add_scalar(int a_new, int b_old)
{
  // log_likelihood_ratio::operator+ (lib/phy/upper/log_likelihood_ratio.cpp:39-75), new value tested first.
  if (a_new == -b_old) {
    return 0;
  }
  if (a_new == 127 || a_new == -127) {
    return a_new;
  }
  if (b_old == 127 || b_old == -127) {
    return b_old;
  }
  return max(-120, min(120, a_new + b_old));
}

__device__ __forceinline__ int llr_add_simd(int a_new, int b_old)
{
  // clamp(adds_epi8(a, b), +-120) (ldpc_rate_dematcher_avx512_impl.cpp:45-58).
  int s = max(-128, min(127, a_new + b_old));
  return max(-120, min(120, s));
}

/// True if position p of the rate-matched stream is combined by the SIMD loop of the reference (the first
/// floor(n / blk) * blk elements of each contiguous combine call), false if by its scalar tail.
__device__ __forceinline__ bool dm_simd_rule(const dm_geom& g, uint32_t p)
{
  if (g.blk == 0) {
    return false;
  }
  uint32_t start, maxlen;
  if (p < g.first_len) {
    uint32_t info_part = (g.s0 < g.info) ? g.info - g.s0 : 0;
    if (p < info_part) {
      start  = 0;
      maxlen = info_part;
    } else {
      start  = info_part;
      maxlen = g.first_len - info_part;
    }
  } else {
    uint32_t rel = p - g.first_len;
    uint32_t lap = rel / g.D;
    uint32_t lp  = rel - lap * g.D;
    if (lp < g.info) {
      start  = g.first_len + lap * g.D;
      maxlen = g.info;
    } else {
      start  = g.first_len + lap * g.D + g.info;
      maxlen = g.D - g.info;
    }
  }
  uint32_t len = min(maxlen, g.E - start);
  return (p - start) < (len / g.blk) * g.blk;
}

/// Rate-matched (de-interleaved) position p -> LLR.
template <bool STAGED>
__device__ __forceinline__ int dm_fetch(const dm_geom& g, const int8_t* __restrict__ in, const int8_t* sm, uint32_t p)
{
  if (STAGED) {
    return sm[p];
  }
  if (g.Qm > 1) {
    uint32_t q = p / g.S;
    uint32_t r = p - q * g.S;
    return __ldg(in + (size_t)r * g.Qm + q);
  }
  return __ldg(in + p);
}

/// New value of soft-buffer byte i given its old value (ldpc_rate_dematcher_impl.cpp:128-201, all passes).
template <bool STAGED>
__device__ __forceinline__ int dm_byte(const dm_geom& g, const int8_t* __restrict__ in, const int8_t* sm, uint32_t i, int val)
{
  bool     is_data = (i < g.info) || (i >= g.sys && i < g.Ncb);
  uint32_t slot    = (i < g.info) ? i : g.info + (i - g.sys);
  uint32_t p0      = (slot >= g.s0) ? slot - g.s0 : slot + g.D - g.s0;
  if (g.new_data) {
    if (i < g.zero_head || i >= g.tail_start) {
      val = 0;
    } else if (i >= g.info && i < g.sys) {
      val = 127;
    }
  }
  if (is_data) {
    for (uint32_t p = p0; p < g.E; p += g.D) {
      int x = dm_fetch<STAGED>(g, in, sm, p);
      if (g.new_data && p < g.first_len) {
        val = x;
      } else {
        val = dm_simd_rule(g, p) ? llr_add_simd(x, val) : llr_add_scalar(x, val);
      }
    }
  }
  return val;
}

#ifndef DM_UNR
#define DM_UNR 4
#endif
#ifndef DM_MINB
#define DM_MINB 6 // caps the registers at 42: sixteen 128-thread CTAs per SM
#endif
template <bool STAGED>
__global__ void __launch_bounds__(256, DM_MINB) rate_dematch_kernel(const cb_desc* __restrict__ descs, int8_t* __restrict__ soft_base,
                                                           uint32_t combine_block)
{
  extern __shared__ __align__(16) int8_t dm_sm[];
  // The descriptor is fetched once, by 16 threads: for a small batch it lives in page-locked HOST memory (the dematcher
  // starts before the descriptors' device copy has arrived), where every separate load would be a trip over the link.
  __shared__ __align__(16) cb_desc sd;
  static_assert(sizeof(cb_desc) % 4 == 0 && sizeof(cb_desc) / 4 <= 32, "descriptor fetch");
  if (threadIdx.x < sizeof(cb_desc) / 4) {
    reinterpret_cast<uint32_t*>(&sd)[threadIdx.x] = reinterpret_cast<const uint32_t*>(descs + blockIdx.x)[threadIdx.x];
  }
  __syncthreads();
  const cb_desc& d = sd;
  if (!(d.flags & FLAG_DEMATCH) || ((d.E <= DM_STAGE_CAP) != STAGED)) {
    return;
  }
  dm_geom g;
  g.N         = d.N;
  g.Ncb       = d.Ncb;
  g.E         = d.E;
  g.F         = d.nof_filler;
  uint32_t Kb = (d.bg == 1) ? 22 : 10;
  g.sys       = (Kb - 2) * d.Z;
  g.info      = g.sys - g.F;
  g.D         = g.info + (g.Ncb - g.sys);
  uint32_t k0 = d.k0;
  g.s0        = (k0 < g.info) ? k0 : ((k0 < g.sys) ? g.info : g.info + (k0 - g.sys));
  g.first_len = g.D - g.s0;
  g.Qm        = d.Qm ? d.Qm : 1;
  g.S         = g.E / g.Qm;
  g.new_data  = d.new_data;
  g.zero_head = min(k0, g.info);
  g.copy_at_end = (g.new_data && g.E <= g.first_len) ? 1 : 0;
  g.blk       = combine_block;
  g.tail_start = g.N; // none
  if (g.copy_at_end) {
    uint32_t end_slot = g.s0 + g.E;
    uint32_t idx_end  = (end_slot <= g.info) ? g.sys : g.sys + (end_slot - g.info);
    if (idx_end >= g.Ncb) {
      idx_end -= g.Ncb;
    }
    if (idx_end != 0) {
      g.tail_start = g.N - (g.Ncb - idx_end);
    }
  }
  g.i_wrap = (g.s0 < g.info) ? g.s0 : g.sys + (g.s0 - g.info);
  {
    uint32_t se = g.s0 + g.E; // first unvisited slot (only meaningful when E < D)
    se          = (se >= g.D) ? se - g.D : se;
    g.i_end     = (se < g.info) ? se : g.sys + (se - g.info);
  }

  const int8_t* __restrict__ in  = d.llr;
  int8_t*                    out = soft_base + (size_t)d.slot * SOFT_STRIDE;

  // ---- stage 1: HBM -> shared memory, de-interleaving on the way -----------------------------------------------------
  if (STAGED) {
    const uint32_t mis  = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 15U); // bytes before `in` in its 16-byte line
    const uint32_t nvec = (mis + g.E + 15) / 16;
    const int8_t*  base = in - mis;
    // Vector v of this thread: `whole` = entirely inside the code block (all but the first / last one), then `w` holds it.
    auto scatter = [&](uint32_t v, bool whole, const uint4& w) {
      int      first = (int)(v * 16) - (int)mis; // index into `in` of byte 0 of this vector
      uint32_t lo    = first < 0 ? (uint32_t)(-first) : 0U;
      uint32_t hi    = min(16U, (uint32_t)((int)g.E - first));
      if (whole) {
        const uint32_t idx = (uint32_t)first;
        if (g.Qm == 8 && (idx & 15U) == 0 && (g.S & 1U) == 0) {
          // Two modulation symbols (rows i, i + 1; i even) x 8 bits: bit j of both rows lands on d[j S + i], d[j S + i + 1].
          const uint32_t i = idx >> 3;
          uint16_t*      d = reinterpret_cast<uint16_t*>(dm_sm + i);
          const uint32_t h = g.S >> 1; // stride between columns in 16-bit units
          d[0 * h] = (uint16_t)__byte_perm(w.x, w.z, 0x0040);
          d[1 * h] = (uint16_t)__byte_perm(w.x, w.z, 0x0051);
          d[2 * h] = (uint16_t)__byte_perm(w.x, w.z, 0x0062);
          d[3 * h] = (uint16_t)__byte_perm(w.x, w.z, 0x0073);
          d[4 * h] = (uint16_t)__byte_perm(w.y, w.w, 0x0040);
          d[5 * h] = (uint16_t)__byte_perm(w.y, w.w, 0x0051);
          d[6 * h] = (uint16_t)__byte_perm(w.y, w.w, 0x0062);
          d[7 * h] = (uint16_t)__byte_perm(w.y, w.w, 0x0073);
        } else if (g.Qm == 1 && (idx & 15U) == 0) {
          *reinterpret_cast<uint4*>(dm_sm + idx) = w;
        } else {
          const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
          uint32_t       i     = idx / g.Qm;
          uint32_t       j     = idx - i * g.Qm;
          uint32_t       p     = (g.Qm > 1) ? j * g.S + i : idx;
#pragma unroll
          for (uint32_t b = 0; b != 16; ++b) {
            dm_sm[p] = (int8_t)(ww[b >> 2] >> (8 * (b & 3)));
            p += g.S;
            if (++j == g.Qm) {
              j = 0;
              ++i;
              p = i;
            }
          }
        }
        return;
      }
      // Partial vector at either end of the code block: byte by byte.
      for (uint32_t b = lo; b < hi; ++b) {
        uint32_t idx = (uint32_t)(first + (int)b);
        uint32_t p   = idx;
        if (g.Qm > 1) {
          uint32_t i = idx / g.Qm;
          p          = (idx - i * g.Qm) * g.S + i;
        }
        dm_sm[p] = __ldg(in + idx);
      }
    };
    // Four vectors per thread and round: their loads are all issued before the first one is consumed, so that a code block
    // costs about one HBM round trip instead of one per vector (the kernel is latency-bound: one dependent chain per CTA).
    constexpr uint32_t UNR = DM_UNR;
    for (uint32_t v0 = threadIdx.x; v0 < nvec; v0 += UNR * blockDim.x) {
      uint4 w[UNR];
      bool  whole[UNR];
#pragma unroll
      for (uint32_t u = 0; u != UNR; ++u) {
        const uint32_t v     = v0 + u * blockDim.x;
        const int      first = (int)(v * 16) - (int)mis;
        whole[u]             = v < nvec && first >= 0 && first + 16 <= (int)g.E;
        w[u]                 = whole[u] ? __ldg(reinterpret_cast<const uint4*>(base) + v) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (uint32_t u = 0; u != UNR; ++u) {
        const uint32_t v = v0 + u * blockDim.x;
        if (v < nvec) {
          scatter(v, whole[u], w[u]);
        }
      }
    }
    __syncthreads();
  }

  // ---- stage 2: soft buffer chunks -------------------------------------------------------------------------------------
  // Only [0, Ncb) and the zeroed tail [tail_start, N) can change.
  const uint32_t nchunks    = (g.N + 15) / 16;
  const uint32_t head_end   = (g.Ncb + 15) / 16;          // chunks [0, head_end) cover [0, Ncb)
  const uint32_t tail_begin = max(head_end, g.tail_start / 16); // chunks [tail_begin, nchunks) cover the zeroed tail
  const uint32_t nwork      = head_end + (nchunks - tail_begin);
  const bool     single     = g.E <= g.D;                 // every data byte is visited at most once

  for (uint32_t w = threadIdx.x; w < nwork; w += blockDim.x) {
    uint32_t c  = (w < head_end) ? w : tail_begin + (w - head_end);
    uint32_t i0 = c * 16;
    uint32_t i1 = i0 + 16;
    bool     full = (i1 <= g.N);
    // A chunk is uniform if no break point lies strictly inside it.
    bool mixed = !full || !single;
#define DM_BP(x) mixed |= ((x) > i0 && (x) < i1)
    DM_BP(g.info);
    DM_BP(g.sys);
    DM_BP(g.Ncb);
    DM_BP(g.zero_head);
    DM_BP(g.tail_start);
    DM_BP(g.i_wrap);
    DM_BP(g.i_end);
#undef DM_BP
    if (!mixed) {
      // Uniform chunk: byte i0 decides for all 16.
      uint32_t i       = i0;
      bool     is_data = (i < g.info) || (i >= g.sys && i < g.Ncb);
      uint32_t slot    = (i < g.info) ? i : g.info + (i - g.sys);
      uint32_t p0      = (slot >= g.s0) ? slot - g.s0 : slot + g.D - g.s0;
      bool     visited = is_data && p0 < g.E;
      if (visited && g.new_data && p0 < g.first_len) {
        // plain copy of 16 consecutive rate-matched positions
        uint4 nv;
        if (STAGED) {
          const uint32_t  sh = (p0 & 3U) * 8U;
          const uint32_t* ws = reinterpret_cast<const uint32_t*>(dm_sm + (p0 & ~3U));
          uint32_t        a0 = ws[0], a1 = ws[1], a2 = ws[2], a3 = ws[3];
          uint32_t        a4 = sh ? ws[4] : 0U;
          nv.x = __funnelshift_r(a0, a1, sh);
          nv.y = __funnelshift_r(a1, a2, sh);
          nv.z = __funnelshift_r(a2, a3, sh);
          nv.w = __funnelshift_r(a3, a4, sh);
        } else {
          int8_t* nb = reinterpret_cast<int8_t*>(&nv);
#pragma unroll
          for (uint32_t b = 0; b != 16; ++b) {
            nb[b] = (int8_t)dm_fetch<false>(g, in, dm_sm, p0 + b);
          }
        }
        *reinterpret_cast<uint4*>(out + i0) = nv;
        continue;
      }
      if (STAGED && visited && dm_simd_rule(g, p0 + 15)) {
        // HARQ combining of 16 consecutive positions inside the SIMD body of one combine_softbits call
        // (ldpc_rate_dematcher_avx512_impl.cpp:29-64): clamp(adds_epi8(new, old), +-120), four bytes per instruction group.
        const uint32_t  sh = (p0 & 3U) * 8U;
        const uint32_t* ws = reinterpret_cast<const uint32_t*>(dm_sm + (p0 & ~3U));
        uint32_t        a0 = ws[0], a1 = ws[1], a2 = ws[2], a3 = ws[3];
        uint32_t        a4 = sh ? ws[4] : 0U;
        uint4           nw = make_uint4(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(a2, a3, sh),
                              __funnelshift_r(a3, a4, sh));
        // A new transmission zeroes the head of the buffer before its wrapped part is combined onto it.
        uint4 ov = (g.new_data && i < g.zero_head) ? make_uint4(0, 0, 0, 0) : *reinterpret_cast<const uint4*>(out + i0);
        auto  comb = [](uint32_t x, uint32_t y) {
          return __vmins4(__vmaxs4(__vaddss4(x, y), 0x88888888U), 0x78787878U);
        };
        *reinterpret_cast<uint4*>(out + i0) = make_uint4(comb(nw.x, ov.x), comb(nw.y, ov.y), comb(nw.z, ov.z), comb(nw.w, ov.w));
        continue;
      }
      if (!visited) {
        if (g.new_data && (i < g.zero_head || i >= g.tail_start)) {
          *reinterpret_cast<uint4*>(out + i0) = make_uint4(0, 0, 0, 0);
        } else if (g.new_data && i >= g.info && i < g.sys) {
          *reinterpret_cast<uint4*>(out + i0) = make_uint4(0x7f7f7f7fU, 0x7f7f7f7fU, 0x7f7f7f7fU, 0x7f7f7f7fU);
        }
        continue; // untouched otherwise
      }
      // visited + combine: per-byte path below
    }
    // Per-byte path.
    __align__(16) int8_t vals[16];
    if (full) {
      *reinterpret_cast<uint4*>(vals) = *reinterpret_cast<const uint4*>(out + i0);
    } else {
      for (uint32_t b = 0; i0 + b < g.N; ++b) {
        vals[b] = out[i0 + b];
      }
    }
    bool changed = false;
#pragma unroll 4
    for (uint32_t b = 0; b != 16; ++b) {
      uint32_t i = i0 + b;
      if (i >= g.N) {
        break;
      }
      int nv  = dm_byte<STAGED>(g, in, dm_sm, i, vals[b]);
      changed |= (nv != vals[b]);
      vals[b] = (int8_t)nv;
    }
    if (!changed) {
      continue;
    }
    if (full) {
      *reinterpret_cast<uint4*>(out + i0) = *reinterpret_cast<const uint4*>(vals);
    } else {
      for (uint32_t b = 0; i0 + b < g.N; ++b) {
        out[i0 + b] = vals[b];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Kernel 2: batched layered normalized min-sum LDPC decoder.
//
// One group of TPC threads per code block (TPC = 32, 64, 128, 256 or 384 >= Z), CBS groups per CTA. Thread j owns lifted
// check j of every layer. Shared memory per code block:
//   tab  : per edge of the processed layers, node base (col * Z) and shift (V mod Z)
//   soft : (K_b + L) * Z int8 a-posteriori LLRs in natural order; the circulant shift is an index rotation
//   c2v  : one int8 per lifted edge, stored in CHECK order (thread-private column), zero before the first visit so that
//          v2c = f(soft, 0) = soft reproduces the reference's "row not initialised" copy (ldpc_decoder_impl.cpp:196-200)
// ---------------------------------------------------------------------------------------------------------------------
struct dec_smem_layout {
  uint32_t tab_off, soft_off, c2v_off, misc_off, total;
};

__host__ __device__ inline dec_smem_layout dec_layout(uint32_t bg, uint32_t Z, uint32_t layer_cap)
{
#ifdef __CUDA_ARCH__
  uint32_t nedges = c_row_ptr[bg - 1][layer_cap];
#else
  uint32_t nedges = ((bg == 1) ? NR_BG1_ROW_PTR : NR_BG2_ROW_PTR)[layer_cap];
#endif
  uint32_t        Kb = (bg == 1) ? 22 : 10;
  dec_smem_layout l;
  l.tab_off  = 0;
  l.soft_off = (nedges * 4 + 15) & ~15U;
  l.c2v_off  = l.soft_off + (((Kb + layer_cap) * Z + 15) & ~15U);
  l.misc_off = l.c2v_off + ((nedges * Z + 15) & ~15U);
  l.total    = l.misc_off + 64;
  return l;
}

template <int TPC>
__device__ __forceinline__ void group_sync(int group)
{
  if (TPC == 32) {
    __syncwarp();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(TPC) : "memory");
  }
}

/// One layer for one lifted check, degree known at compile time so the v2c messages stay in registers.
template <int DEG>
__device__ __forceinline__ void
process_check(int8_t* __restrict__ soft, int8_t* __restrict__ c2v, const uint32_t* __restrict__ tab, int j, int Z, uint32_t mult)
{
  int      pk[DEG];
  int      m1 = 120, m2 = 120, amin = 0;
  unsigned sg = 0;
#pragma unroll
  for (int e = 0; e != DEG; ++e) {
    uint32_t te = tab[e];
    int      k  = j + (int)(te >> 16);
    k           = (k >= Z) ? k - Z : k;
    int addr    = (int)(te & 0xffffU) + k;
    int s       = soft[addr];
    int c       = c2v[e * Z + j];
    // compute_var_to_check_msgs (ldpc_decoder_avx512.cpp:81-121): clamp(soft - c2v, +-120); +-infinity is sticky.
    int x = max(-120, min(120, s - c));
    x     = (s >= 127) ? 127 : x;
    x     = (s <= -127) ? -127 : x;
    pk[e] = (addr << 8) | (x & 0xff);
    // analyze_var_to_check_msgs (:123-165): min1 / min2 / argmin with strict '<', both minima start at 120.
    int a = abs(x);
    sg ^= (unsigned)x;
    m2   = min(m2, max(m1, a));
    amin = (a < m1) ? e : amin;
    m1   = min(m1, a);
  }
  // mm512::scale_epi8 (avx512_support.h:65-107): (m * (uint16)(sf * 65536)) >> 16 on magnitudes in [0, 120].
  int M1 = mult ? (int)(((uint32_t)m1 * mult) >> 16) : m1;
  int M2 = mult ? (int)(((uint32_t)m2 * mult) >> 16) : m2;
#pragma unroll
  for (int e = 0; e != DEG; ++e) {
    int x    = (int)(int8_t)(pk[e] & 0xff);
    int addr = pk[e] >> 8;
    // compute_check_to_var_msgs (:167-216): magnitude excluding self, sign = parity xor own sign (0 counts as +).
    int mag = (e == amin) ? M2 : M1;
    int c   = ((int)(sg ^ (unsigned)x) < 0) ? -mag : mag;
    c2v[e * Z + j] = (int8_t)c;
    // compute_soft_bits (:218-259): promotion sum; |c| <= 95 is never infinite, an infinite v2c is sticky.
    int sum = c + x;
    int r   = (sum > 120) ? 127 : ((sum < -120) ? -127 : sum);
    r       = (x >= 127) ? 127 : r;
    r       = (x <= -127) ? -127 : r;
    soft[addr] = (int8_t)r;
  }
}

/// Hard decision of the first K soft bits (MSB-first, bit = (llr <= 0)), "no zero LLR" test and CRC of the first
/// K - F bits, cooperatively by the TPC threads of a group. Returns (to every thread of the group) 1 if the CRC is zero
/// and no LLR is zero. get_hard_bits + hard_decision + crc->calculate (ldpc_decoder_impl.cpp:126-134).
template <int TPC>
__device__ __forceinline__ uint32_t hard_decision_crc(const int8_t* soft,
                                                      uint8_t*      bits_out,
                                                      uint32_t      K,
                                                      uint32_t      nb,
                                                      int           poly,
                                                      uint32_t*     misc,
                                                      int           t,
                                                      int           group,
                                                      bool          check_zero = true)
{
  const int      lane   = t & 31;
  const int      warp   = t >> 5;
  constexpr int  NW     = TPC / 32;
  const uint32_t nwords = (K + 31) / 32;
  const uint32_t gen = crc_gen(poly), order = crc_order(poly);
  const uint32_t nfull = nb / 32, rem = nb % 32;

  if (t == 0) {
    misc[0] = 0; // xor of CRC contributions
    misc[1] = 0; // any zero LLR
    misc[2] = 0; // partial word
  }
  group_sync<TPC>(group);

  uint32_t my_word = 0, my_idx = 0xffffffffU, zero_any = 0;
  uint32_t slot_in_warp = 0;
  uint32_t acc = 0;
  for (uint32_t w = warp; w < nwords; w += NW) {
    uint32_t idx = w * 32 + lane;
    int      s   = (idx < K) ? (int)soft[idx] : 1;
    uint32_t b   = __ballot_sync(0xffffffffU, s <= 0);
    uint32_t z   = __ballot_sync(0xffffffffU, s == 0);
    uint32_t word = __brev(b);
    zero_any |= z;
    if (lane == 0) {
      // MSB-first packed bytes, written as one 32-bit store (bits_out is 4-byte aligned, slot stride 1056).
      reinterpret_cast<uint32_t*>(bits_out)[w] = __byte_perm(word, 0, 0x0123);
    }
    if ((uint32_t)lane == (slot_in_warp & 31)) {
      my_word = word;
      my_idx  = w;
    }
    ++slot_in_warp;
    if ((slot_in_warp & 31) == 0 || w + NW >= nwords) {
      // Every lane that holds a word folds it: contribution = (word(x) * x^order mod g) * x^(32 * (nfull - 1 - idx)).
      if (poly != 0 && my_idx != 0xffffffffU) {
        if (my_idx < nfull) {
          uint32_t rr = crc_push_bits(0, my_word, 32, gen, order);
          acc ^= gf2_mulmod(rr, c_xpow32[poly - 1][nfull - 1 - my_idx], gen, order);
        } else if (my_idx == nfull && rem != 0) {
          misc[2] = my_word;
        }
      }
      my_idx = 0xffffffffU;
    }
  }
  if (poly != 0) {
    acc = __reduce_xor_sync(0xffffffffU, acc);
    if (lane == 0 && acc != 0) {
      atomicXor(&misc[0], acc);
    }
  }
  if (lane == 0 && zero_any != 0) {
    atomicOr(&misc[1], 1U);
  }
  group_sync<TPC>(group);
  uint32_t ok = 0;
  if (poly != 0) {
    uint32_t total = misc[0];
    if (rem != 0) {
      // total * x^rem + partial bits: continue the division over the last `rem` bits.
      total = crc_push_bits(total, misc[2], rem, gen, order);
    }
    ok = (total == 0 && (!check_zero || misc[1] == 0)) ? 1U : 0U;
  }
  group_sync<TPC>(group);
  return ok;
}

template <int TPC, int CBS>
__global__ void __launch_bounds__(TPC* CBS, (TPC * CBS >= 256) ? 2 : 4) ldpc_decode_kernel(const cb_desc* __restrict__ descs,
                                                               const uint32_t* __restrict__ order,
                                                               cb_result* __restrict__ results,
                                                               const int8_t* __restrict__ soft_base,
                                                               uint8_t* __restrict__ bits_base,
                                                               uint32_t* __restrict__ crc_flags,
                                                               uint32_t nof_cbs,
                                                               uint32_t smem_per_cb)
{
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int      group = threadIdx.x / TPC;
  const int      t     = threadIdx.x % TPC;
  if (blockIdx.x * CBS + group >= nof_cbs) {
    return;
  }
  const uint32_t cb = order[blockIdx.x * CBS + group];
  const cb_desc& d  = descs[cb];
  if (!(d.flags & FLAG_DECODE)) {
    return;
  }
  uint8_t* bits_slot = bits_base + (size_t)d.slot * BITS_STRIDE;
  const uint32_t Z  = d.Z;
  const uint32_t bg = d.bg;
  const uint32_t Kb = (bg == 1) ? 22 : 10;
  const uint32_t K  = Kb * Z;

  // Code blocks whose CRC is already ok are only dematched (pusch_decoder_impl.cpp:335-345).
  if ((d.flags & FLAG_TRACK_CRC) && !d.new_data && crc_flags[d.slot] != 0) {
    if (t == 0) {
      results[cb] = {0, 1U, 0U, 2U};
    }
    if (d.bits_out != nullptr) {
      for (uint32_t i = t; i < (K + 31) / 32; i += TPC) {
        reinterpret_cast<uint32_t*>(d.bits_out)[i] = reinterpret_cast<const uint32_t*>(bits_slot)[i];
      }
    }
    return;
  }

  uint8_t*              smem = smem_raw + (size_t)group * smem_per_cb;
  const dec_smem_layout lay  = dec_layout(bg, Z, d.layer_cap);
  uint32_t*             tab  = reinterpret_cast<uint32_t*>(smem + lay.tab_off);
  int8_t*               soft = reinterpret_cast<int8_t*>(smem + lay.soft_off);
  int8_t*               c2v  = reinterpret_cast<int8_t*>(smem + lay.c2v_off);
  uint32_t*             misc = reinterpret_cast<uint32_t*>(smem + lay.misc_off);

  // ---- prologue: clear soft + c2v, load the decoder input, find the last non-zero LLR ------------------------------
  {
    uint4*   z4 = reinterpret_cast<uint4*>(smem + lay.soft_off);
    uint32_t n4 = (lay.misc_off - lay.soft_off) / 16;
    for (uint32_t i = t; i < n4; i += TPC) {
      z4[i] = make_uint4(0, 0, 0, 0);
    }
    if (t < 16) {
      misc[t] = 0;
    }
  }
  group_sync<TPC>(group);

  const int8_t* src    = (d.flags & FLAG_USE_HARQ) ? soft_base + (size_t)d.slot * SOFT_STRIDE : d.llr;
  uint32_t      cap_in = (Kb + d.layer_cap) * Z - 2 * Z;
  uint32_t      n_load = min(min(d.n_in, d.scan_len), cap_in);
  uint32_t      last   = 0;
  {
    int8_t*  dst = soft + 2 * Z;
    uint32_t nv  = n_load / 16;
    bool     al  = ((2 * Z) % 4) == 0;
    for (uint32_t i = t; i < nv; i += TPC) {
      uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
      if (v.x | v.y | v.z | v.w) {
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 3; q >= 0; --q) {
          if (w[q] != 0) {
            // index of the last non-zero byte (little endian: highest byte = highest index)
            uint32_t hb = 3 - (__clz(w[q]) >> 3);
            last        = max(last, i * 16 + q * 4 + hb + 1);
            break;
          }
        }
        if (al) {
          uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + i * 16);
          d32[0] = v.x;
          d32[1] = v.y;
          d32[2] = v.z;
          d32[3] = v.w;
        } else {
          const int8_t* vb = reinterpret_cast<const int8_t*>(&v);
#pragma unroll
          for (int q = 0; q != 16; ++q) {
            dst[i * 16 + q] = vb[q];
          }
        }
      }
    }
    for (uint32_t i = nv * 16 + t; i < n_load; i += TPC) {
      int8_t x = src[i];
      if (x != 0) {
        dst[i] = x;
        last   = max(last, i + 1);
      }
    }
    last = __reduce_max_sync(0xffffffffU, last);
    if ((t & 31) == 0 && last != 0) {
      atomicMax(&misc[8], last);
    }
  }
  group_sync<TPC>(group);
  last = misc[8];

  const uint32_t nb      = K - d.nof_filler;
  const int      poly    = d.crc_poly;
  const uint32_t mode    = d.mode;
  int            iters   = -1;
  uint32_t       crc_ok  = 0;
  uint32_t       L       = 0;
  uint32_t       status  = 0;

  if (last == 0) {
    // All-zero input (ldpc_decoder_impl.cpp:88-94): without a CRC calculator the output becomes all ones.
    if (mode != MODE_EARLY_STOP) {
      for (uint32_t i = t; i < K; i += TPC) {
        soft[i] = -1;
      }
      group_sync<TPC>(group);
      crc_ok = hard_decision_crc<TPC>(soft, bits_slot, K, nb, (mode == MODE_CRC_AT_END) ? poly : 0, misc, t, group, false);
      if (mode == MODE_CRC_AT_END && crc_ok) {
        iters = d.max_it;
      }
    }
  } else {
    uint32_t cbl = max(last + 2 * Z, K + 4 * Z);
    cbl          = ((cbl + Z - 1) / Z) * Z;
    L            = cbl / Z - Kb;
    if (L > d.layer_cap) {
      status = 1;
      L      = 0;
    }
    // Edge table of the processed layers.
    const uint32_t nedges = c_row_ptr[bg - 1][L];
    for (uint32_t e = t; e < nedges; e += TPC) {
      uint32_t sh = c_shift[bg - 1][d.ils][e] % Z;
      tab[e]      = ((uint32_t)c_col[bg - 1][e] * Z) | (sh << 16);
    }
    group_sync<TPC>(group);

    const uint32_t mult = d.scale_mult;
    const int      j    = t;
    for (uint32_t it = 0; it != d.max_it && L != 0; ++it) {
      for (uint32_t l = 0; l != L; ++l) {
        uint32_t e0  = c_row_ptr[bg - 1][l];
        int      deg = (int)c_row_ptr[bg - 1][l + 1] - (int)e0;
        if (j < (int)Z) {
          int8_t*         c2v_row = c2v + (size_t)e0 * Z;
          const uint32_t* tab_row = tab + e0;
          switch (deg) {
            case 3:
              process_check<3>(soft, c2v_row, tab_row, j, Z, mult);
              break;
            case 4:
              process_check<4>(soft, c2v_row, tab_row, j, Z, mult);
              break;
            case 5:
              process_check<5>(soft, c2v_row, tab_row, j, Z, mult);
              break;
            case 6:
              process_check<6>(soft, c2v_row, tab_row, j, Z, mult);
              break;
            case 7:
              process_check<7>(soft, c2v_row, tab_row, j, Z, mult);
              break;
            case 8:
              process_check<8>(soft, c2v_row, tab_row, j, Z, mult);
              break;
            case 9:
              process_check<9>(soft, c2v_row, tab_row, j, Z, mult);
              break;
            case 10:
              process_check<10>(soft, c2v_row, tab_row, j, Z, mult);
              break;
            case 19:
              process_check<19>(soft, c2v_row, tab_row, j, Z, mult);
              break;
            default:
              __trap(); // the 3GPP base graphs have no other row degree
              break;
          }
        }
        group_sync<TPC>(group);
      }
      if (mode == MODE_EARLY_STOP) {
        if (hard_decision_crc<TPC>(soft, bits_slot, K, nb, poly, misc, t, group)) {
          iters  = (int)it + 1;
          crc_ok = 1;
          break;
        }
      }
    }
    if (status == 0 && mode != MODE_EARLY_STOP) {
      crc_ok = hard_decision_crc<TPC>(soft, bits_slot, K, nb, (mode == MODE_CRC_AT_END) ? poly : 0, misc, t, group, false);
      if (mode == MODE_CRC_AT_END && crc_ok) {
        iters = d.max_it;
      }
    }
  }

  if (t == 0) {
    results[cb] = {iters, crc_ok, L, status};
    if (d.flags & FLAG_TRACK_CRC) {
      crc_flags[d.slot] = crc_ok;
    }
  }
  if (d.bits_out != nullptr) {
    // hard_decision_crc ended with a group barrier, so the slot bytes written by lane 0 of each warp are visible
    // to the group (global memory, same CTA) after a fence.
    __threadfence_block();
    group_sync<TPC>(group);
    for (uint32_t i = t; i < (K + 31) / 32; i += TPC) {
      reinterpret_cast<uint32_t*>(d.bits_out)[i] = reinterpret_cast<const uint32_t*>(bits_slot)[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Kernel 3: CRC by carry-less folding. One CTA per message. Thread t folds the 16-byte chunks c' = t, t + T, ... counted
// from the END of the message (so leading zero padding is free), Horner-style across passes; lanes are then aligned with
// x^(128 t) and XOR-reduced (warp shuffles, then shared memory).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int CRC_THREADS = 1024;

__device__ __forceinline__ uint32_t block_crc_bytes(const uint8_t* __restrict__ msg, uint32_t nbytes, int poly, uint32_t* lut /*256*/,
                                                    uint32_t* red /*33*/)
{
  const uint32_t gen = crc_gen(poly), order = crc_order(poly);
  const uint32_t mask = (1U << order) - 1;
  const int      t    = threadIdx.x;
  // Byte table: lut[b] = (b(x) * x^order) mod g.
  if (t < 256) {
    lut[t] = crc_push_bits(0, (uint32_t)t << 24, 8, gen, order);
  }
  if (t < 33) {
    red[t] = 0;
  }
  __syncthreads();
  const uint32_t nchunks = (nbytes + 15) / 16;
  const uint32_t npass   = (nchunks + CRC_THREADS - 1) / CRC_THREADS;
  const uint32_t xpass   = gf2_mulmod(c_xpow128[poly - 1][CRC_THREADS / 2], c_xpow128[poly - 1][CRC_THREADS / 2], gen, order);
  uint32_t       acc     = 0;
  bool           any     = false;
  for (int p = (int)npass - 1; p >= 0; --p) {
    uint32_t cp = (uint32_t)t + (uint32_t)p * CRC_THREADS; // chunk index from the end
    if (cp >= nchunks) {
      continue;
    }
    if (any) {
      acc = gf2_mulmod(acc, xpass, gen, order);
    }
    any          = true;
    long long b0 = (long long)nbytes - 16LL * (cp + 1);
    uint32_t  r  = 0;
#pragma unroll
    for (int i = 0; i != 16; ++i) {
      long long b    = b0 + i;
      uint32_t  byte = (b >= 0) ? msg[b] : 0U;
      r              = ((r << 8) & mask) ^ lut[((r >> (order - 8)) ^ byte) & 0xff];
    }
    acc ^= r;
  }
  if (any && t != 0) {
    acc = gf2_mulmod(acc, c_xpow128[poly - 1][t], gen, order);
  }
  acc = __reduce_xor_sync(0xffffffffU, acc);
  if ((t & 31) == 0) {
    red[t >> 5] = acc;
  }
  __syncthreads();
  if (t < 32) {
    uint32_t v = red[t];
    v          = __reduce_xor_sync(0xffffffffU, v);
    if (t == 0) {
      red[32] = v;
    }
  }
  __syncthreads();
  return red[32];
}

struct crc_job {
  const uint8_t* msg;
  uint32_t       nbits;
  uint32_t       poly;
};

__global__ void __launch_bounds__(CRC_THREADS) crc_kernel(const crc_job* __restrict__ jobs, uint32_t* __restrict__ out)
{
  __shared__ uint32_t lut[256];
  __shared__ uint32_t red[33];
  const crc_job       job    = jobs[blockIdx.x];
  uint32_t            nbytes = job.nbits / 8, rem = job.nbits % 8;
  uint32_t            crc    = block_crc_bytes(job.msg, nbytes, (int)job.poly, lut, red);
  if (threadIdx.x == 0) {
    if (rem != 0) {
      crc = crc_push_bits(crc, (uint32_t)job.msg[nbytes] << 24, rem, crc_gen((int)job.poly), crc_order((int)job.poly));
    }
    out[blockIdx.x] = crc;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// TB assembly (pusch_decoder_impl.cpp:384-497), so that a transport block of 152 code blocks is handled by 152 warps instead of
// one CTA:
//   tb_gather_kernel    one WARP per code block: copies its payload bits (dropping CB CRC, filler and zero padding)
//                       to their bit position in the transport block, and computes its share of the TB CRC24A:
//                       (payload(x) x^24 mod g) * x^(bits behind it) mod g  (CRCs are linear over GF(2)).
//                       The warp that finishes LAST for its transport block (a counter per TB) also finalises it:
//   tb_finalize_warp    all code blocks ok? XOR of the shares == 0? result + CRC-flag reset (one warp per transport block)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TBG_WARPS = 16;

/// Byte tables tabs[k][b] = (b(x) x^(8k) x^order) mod g, k = 0..3: copied from the precomputed global tables (4 KB).
__device__ __forceinline__ void build_crc_tables(uint32_t* tabs, int poly, int t, int nthreads)
{
  const uint32_t* src = g_crc_tabs[poly - 1];
  for (int i = t; i < 1024; i += nthreads) {
    tabs[i] = src[i];
  }
}

/// CRC state after the first nb bits of a message held as 32-bit words, by ONE warp. BSWAP: the words are in memory
/// byte order (MSB-first bytes, as the decoders store them), else MSB-first numeric words.
template <bool BSWAP>
__device__ __forceinline__ uint32_t warp_crc_words(const uint32_t* words, uint32_t nb, int poly, const uint32_t* tabs, int lane)
{
  const uint32_t gen = crc_gen(poly), order = crc_order(poly);
  const uint32_t nfull = nb / 32, rem = nb % 32;
  const uint32_t per   = (nfull + 31) / 32;
  uint32_t       w0 = min(nfull, (uint32_t)lane * per), w1 = min(nfull, w0 + per);
  uint32_t       acc = 0;
  const uint32_t sh  = (32 - order) / 8; // tables that multiply a register byte by x^32
  for (uint32_t w = w0; w < w1; ++w) {
    uint32_t word = BSWAP ? __byte_perm(words[w], 0, 0x0123) : words[w];
    uint32_t r    = tabs[3 * 256 + (word >> 24)] ^ tabs[2 * 256 + ((word >> 16) & 0xff)] ^ tabs[256 + ((word >> 8) & 0xff)] ^
                 tabs[word & 0xff];
    uint32_t m = tabs[sh * 256 + (acc & 0xff)] ^ tabs[(sh + 1) * 256 + ((acc >> 8) & 0xff)];
    if (order == 24) {
      m ^= tabs[(sh + 2) * 256 + (acc >> 16)];
    }
    acc = m ^ r;
  }
  if (w1 > w0 && w1 != nfull) {
    acc = gf2_mulmod(acc, c_xpow32[poly - 1][nfull - w1], gen, order);
  }
  acc = __reduce_xor_sync(0xffffffffU, acc);
  if (rem != 0) {
    uint32_t last = BSWAP ? __byte_perm(words[nfull], 0, 0x0123) : words[nfull];
    acc           = crc_push_bits(acc, last, rem, gen, order);
  }
  return acc;
}

/// Share of thread `tid` (of `nthreads`) in the CRC state after the first nb bits of a message held as MSB-first numeric
/// 32-bit words: the XOR of all shares is the CRC. The bit string is taken right-aligned in ceil(nb / 32) words - leading
/// zero bits do not change a zero-initialised CRC - so that every share covers whole words and no serial tail is left.
__device__ __forceinline__ uint32_t crc_share_words(const uint32_t* words, uint32_t nb, int poly, const uint32_t* tabs,
                                                    uint32_t tid, uint32_t nthreads)
{
  const uint32_t gen = crc_gen(poly), order = crc_order(poly);
  const uint32_t nw  = (nb + 31) / 32;
  const uint32_t rs  = (32 - nb % 32) % 32; // right shift of the whole string
  const uint32_t per = (nw + nthreads - 1) / nthreads;
  const uint32_t w0 = min(nw, tid * per), w1 = min(nw, w0 + per);
  const uint32_t sh  = (32 - order) / 8; // tables that multiply a register byte by x^32
  uint32_t       acc = 0;
  uint32_t       prev = (w0 != 0 && w0 < nw) ? words[w0 - 1] : 0U;
  for (uint32_t w = w0; w < w1; ++w) {
    const uint32_t cur  = words[w];
    const uint32_t word = __funnelshift_r(cur, prev, rs); // (prev:cur) >> rs; rs == 0 gives cur
    prev                = cur;
    uint32_t r = tabs[3 * 256 + (word >> 24)] ^ tabs[2 * 256 + ((word >> 16) & 0xff)] ^ tabs[256 + ((word >> 8) & 0xff)] ^
                 tabs[word & 0xff];
    uint32_t m = tabs[sh * 256 + (acc & 0xff)] ^ tabs[(sh + 1) * 256 + ((acc >> 8) & 0xff)];
    if (order == 24) {
      m ^= tabs[(sh + 2) * 256 + (acc >> 16)];
    }
    acc = m ^ r;
  }
  if (w1 > w0 && w1 != nw) {
    const uint32_t m = __ldg(&g_xpow32[poly - 1][nw - w1]);
    acc              = (order == 24) ? gf2_mulmod_fixed<24>(acc, m, gen) : gf2_mulmod_fixed<16>(acc, m, gen);
  }
  return acc;
}

/// x^e mod g, by one warp: lane i contributes x^(2^i) if bit i of e is set, the factors are multiplied in a butterfly.
__device__ __forceinline__ uint32_t warp_xpow(uint32_t e, int poly, int lane)
{
  const uint32_t gen = crc_gen(poly), order = crc_order(poly);
  uint32_t       f   = ((e >> lane) & 1U) ? c_xpow2[poly - 1][lane] : 1U;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    f = gf2_mulmod(f, __shfl_xor_sync(0xffffffffU, f, d), gen, order);
  }
  return f;
}

/// TB verdict by one warp (all code blocks ok? XOR of their CRC shares zero?), and the CRC-flag reset of a false positive
/// (pusch_decoder_impl.cpp:417-429).
__device__ __forceinline__ void tb_finalize_warp(const tb_desc& tb, uint32_t tbi, const cb_desc* __restrict__ descs,
                                                 tb_result_dev* __restrict__ tb_results, const uint32_t* crc_share,
                                                 uint32_t* crc_flags, int lane)
{
  bool     ok  = true;
  uint32_t crc = 0;
  for (uint32_t i = lane; i < tb.nof_cbs; i += 32) {
    ok = ok && (crc_flags[descs[tb.first_cb + i].slot] != 0);
    crc ^= crc_share[tb.first_cb + i];
  }
  const uint32_t all_ok = __all_sync(0xffffffffU, ok) ? 1U : 0U;
  crc                   = __reduce_xor_sync(0xffffffffU, crc);
  if (tb.nof_cbs == 1 || !all_ok) {
    // Single code block: its CRC is the TB CRC. Otherwise the TB CRC is only checked when every code block is ok.
    if (lane == 0) {
      tb_results[tbi] = {all_ok, all_ok};
    }
    return;
  }
  if (lane == 0) {
    tb_results[tbi] = {crc == 0 ? 1U : 0U, 1U};
  }
  if (crc != 0) {
    // At least one code block is a false positive: reset them all (pusch_decoder_impl.cpp:425-428).
    for (uint32_t i = lane; i < tb.nof_cbs; i += 32) {
      crc_flags[descs[tb.first_cb + i].slot] = 0;
    }
  }
}

__global__ void __launch_bounds__(TBG_WARPS * 32) tb_gather_kernel(const tb_desc* __restrict__ tbs,
                                                                   const cb_desc* __restrict__ descs,
                                                                   const uint32_t* __restrict__ tb_of_cb,
                                                                   uint32_t nof_cbs,
                                                                   const uint8_t* __restrict__ bits_base,
                                                                   uint8_t* __restrict__ tb_out,
                                                                   uint32_t* crc_share,
                                                                   tb_result_dev* __restrict__ tb_results,
                                                                   uint32_t* crc_flags,
                                                                   uint32_t* tb_done,
                                                                   const cb_result* __restrict__ cb_results,
                                                                   cb_result* __restrict__ cb_results_host)
{
  __shared__ uint32_t tabs[1024];
  const int           t = threadIdx.x, lane = t & 31;
  const int           warp = __shfl_sync(0xffffffffU, t >> 5, 0);
  build_crc_tables(tabs, 1, t, TBG_WARPS * 32);
  __syncthreads();
  const uint32_t cb = blockIdx.x * TBG_WARPS + warp;
  if (cb >= nof_cbs) {
    return;
  }
  const uint32_t tbi = tb_of_cb[cb];
  if (tbi == 0xffffffffU) {
    return;
  }
  const tb_desc   tb   = tbs[tbi];
  if (cb_results_host != nullptr && lane == 0) {
    // Small batch: the per-code-block results reach the page-locked result buffer from here (one kernel touches host
    // memory instead of every decoder CTA, whose completion would wait for the link).
    cb_results_host[cb] = cb_results[cb];
  }
  // Called by every warp once its code block is in place: the last one of a transport block finalises it. The counter
  // returns to zero, so it needs no clearing between batches.
  auto code_block_done = [&]() {
    __threadfence(); // this warp's share and TB words before the count
    uint32_t prev = 0;
    if (lane == 0) {
      prev = atomicAdd(&tb_done[tbi], 1U);
    }
    prev = __shfl_sync(0xffffffffU, prev, 0);
    if (prev + 1 == tb.nof_cbs) {
      __threadfence(); // the other warps' shares after the count
      if (lane == 0) {
        tb_done[tbi] = 0;
      }
      tb_finalize_warp(tb, tbi, descs, tb_results, crc_share, crc_flags, lane);
    }
  };
  const uint32_t  r    = cb - tb.first_cb;
  const uint32_t* src  = reinterpret_cast<const uint32_t*>(bits_base + (size_t)descs[cb].slot * BITS_STRIDE);
  uint32_t*       out  = reinterpret_cast<uint32_t*>(tb_out + tb.out_offset);
  if (tb.nof_cbs == 1) {
    // The code-block CRC is the TB CRC: the payload is the first TBS bits.
    const uint32_t nw = (tb.tbs_bits / 8 + 3) / 4;
    for (uint32_t i = lane; i < nw; i += 32) {
      out[i] = src[i];
    }
    if (lane == 0) {
      crc_share[cb] = 0;
    }
    code_block_done();
    return;
  }
  const uint32_t Lp = tb.cb_data_bits;
  // (1) share of the TB CRC
  uint32_t crc = warp_crc_words<true>(src, Lp, 1, tabs, lane);
  uint32_t pw  = warp_xpow(Lp * (tb.nof_cbs - 1 - r), 1, lane);
  if (lane == 0) {
    crc_share[cb] = gf2_mulmod(crc, pw, crc_gen(1), crc_order(1));
  }
  // (2) payload bits [0, Lp) -> TB bits [o, o + Lp). A 32-bit output word belongs to the code block holding its first bit.
  const uint32_t  o     = r * Lp;
  const uint32_t  m0    = (o + 31) / 32;
  const uint32_t  m1    = (o + Lp + 31) / 32;
  const bool      last  = (r + 1 == tb.nof_cbs);
  // The HARQ slots of a TB's code blocks need not be consecutive (rx_buffer absolute code-block ids): cb_desc::slot.
  const uint32_t* nxt   = reinterpret_cast<const uint32_t*>(bits_base + (size_t)descs[last ? cb : cb + 1].slot * BITS_STRIDE);
  for (uint32_t m = m0 + lane; m < m1; m += 32) {
    uint32_t w  = 32 * m - o;
    uint32_t sw = w >> 5, sh = w & 31;
    uint32_t A  = __byte_perm(src[sw], 0, 0x0123);
    uint32_t v  = A;
    if (sh != 0) {
      uint32_t B = __byte_perm(src[sw + 1], 0, 0x0123);
      v          = (A << sh) | (B >> (32 - sh));
    }
    uint32_t avail = Lp - w;
    if (avail < 32) {
      uint32_t tail = last ? 0U : (__byte_perm(nxt[0], 0, 0x0123) >> avail);
      v             = (v & (0xffffffffU << (32 - avail))) | tail;
    }
    out[m] = __byte_perm(v, 0, 0x0123);
  }
  code_block_done();
}

// ---------------------------------------------------------------------------------------------------------------------
// Host -> device gather of a batch's soft bits when they come in many separate page-locked pieces (one buffer per
// decoder instance behind the plugin interface, pusch_decoder_buffer::get_next_block_view): the SMs read the mapped host
// memory themselves. 64 cudaMemcpyAsync calls of 1.36 MB reach 82 % of what ONE copy of 87 MB reaches on this link (the
// copy engine ramps up per copy, on one stream or on several); a kernel keeps the link full across the pieces.
// One CTA moves 8 KB per turn: 128 threads x four 16-byte loads in flight (<= 32 registers, so that a CTA fits beside two
// decoder CTAs on an SM).
// ---------------------------------------------------------------------------------------------------------------------
struct h2d_piece {
  const uint8_t* src; // device-visible address of page-locked host memory, 16-byte aligned
  uint8_t*       dst; // device memory, 16-byte aligned
  uint64_t       bytes;
};
constexpr uint32_t H2D_MAX_PIECES = 256;
constexpr uint32_t H2D_CHUNK      = 8192;
constexpr uint32_t H2D_THREADS    = 128;

__global__ void __launch_bounds__(H2D_THREADS, 16) h2d_gather_kernel(const h2d_piece* __restrict__ pieces, uint32_t nof_pieces)
{
  __shared__ h2d_piece sp[H2D_MAX_PIECES];
  __shared__ uint32_t  first[H2D_MAX_PIECES + 1]; // first chunk of every piece
  for (uint32_t i = threadIdx.x; i < nof_pieces; i += H2D_THREADS) {
    sp[i] = pieces[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    for (uint32_t i = 0; i != nof_pieces; ++i) {
      first[i] = acc;
      acc += static_cast<uint32_t>((sp[i].bytes + H2D_CHUNK - 1) / H2D_CHUNK);
    }
    first[nof_pieces] = acc;
  }
  __syncthreads();
  const uint32_t total = first[nof_pieces];
  uint32_t       p     = 0;
  for (uint32_t chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    while (chunk >= first[p + 1]) {
      ++p;
    }
    const uint64_t off = static_cast<uint64_t>(chunk - first[p]) * H2D_CHUNK;
    const uint64_t rem = sp[p].bytes - off;
    const uint32_t len = rem < H2D_CHUNK ? static_cast<uint32_t>(rem) : H2D_CHUNK;
    const uint4*   src = reinterpret_cast<const uint4*>(sp[p].src + off);
    uint4*         dst = reinterpret_cast<uint4*>(sp[p].dst + off);
    const uint32_t nv  = len / 16;
    uint4          v[4];
#pragma unroll
    for (uint32_t k = 0; k != 4; ++k) {
      uint32_t i = threadIdx.x + k * H2D_THREADS;
      if (i < nv) {
        v[k] = __ldcs(src + i);
      }
    }
#pragma unroll
    for (uint32_t k = 0; k != 4; ++k) {
      uint32_t i = threadIdx.x + k * H2D_THREADS;
      if (i < nv) {
        dst[i] = v[k];
      }
    }
    const uint32_t tail = len - nv * 16;
    if (threadIdx.x < tail) {
      sp[p].dst[off + nv * 16 + threadIdx.x] = sp[p].src[off + nv * 16 + threadIdx.x];
    }
  }
}

} // namespace pusch_dec
