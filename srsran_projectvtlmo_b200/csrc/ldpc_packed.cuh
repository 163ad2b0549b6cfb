// Kernel 2b: packed layered LDPC decoder - FOUR (or two) code blocks per CTA, one thread per lifted check, the code blocks
// in the binary16 lanes of two (one) registers (arithmetic: ldpc_packed_math.h, verified on the CPU against the oracle by
// tests/test_packed_math_cpu.py). Used for groups of code blocks that share base graph, lifting size (Z >= 144,
// Z % 16 == 0), CRC, mode, iteration limit and scaling, whose state fits in shared memory - i.e. the high-rate PUSCH
// transport blocks of BASELINE config 2 (BG1, Z = 384, 4 layers). Single code blocks run on ldpc_decode_q4_kernel (four
// lifted checks of ONE code block per thread, same arithmetic), the remaining shapes on ldpc_decode_kernel.
//
// The number of layers is NOT derived from the data here (ldpc_decoder_impl.cpp:86-114 trims trailing zero LLRs): the
// host's upper bound (layer_cap) is processed, because a layer whose extension node holds only zero LLRs is a no-op
// for every other node (min1 = 0 => all its messages are 0, soft' = clamp(soft - 0) + 0), see DESIGN.md.
//
// Shared memory per CTA (NR = registers per thread, 2 or 1):
//   tab    per edge: (byte offset of the variable node in `soft`, circulant shift)
//   soft   (K_b + L) * Z x 4 NR bytes: the halves S + 1152 of the 2 NR code blocks per variable-node lift
//   c2v    edges * Z x 2 NR bytes: the message bytes c + 128 of the 2 NR code blocks per lifted edge, check order
//          (thread-private between barriers)
//   hb     4 x K/32 words of hard decisions, crc byte tables 4 x 256 words, flags / CRC shares, four lane_state records
//
// Forms of ldpc_decode4_kernel<TPC, ZT, NR, TM>:
//   NR = 2, TM = 1   four code blocks per CTA, messages in TENSOR MEMORY (one 32-bit column per lifted edge), v2c parked in the
//                    soft array, 80 registers: two CTAs per SM with 256 columns each - the headline form; with all 512 columns
//                    one CTA per SM for groups of up to ~13 layers (HARQ retransmissions)
//   NR = 1, TM = 1   two code blocks with MANY layers (up to all 46) per CTA, all 512 columns (two edges per column), the last
//                    layers' messages in shared memory, v2c in registers, one CTA per SM; optional cp.async.bulk input staging
//   NR = 2, TM = 0   the round-1 form: messages in shared memory, one CTA per SM (also for lifting sizes with Z % 32 != 0)
//   NR = 1, TM = 0   two code blocks per CTA, messages in shared memory: small batches (latency mode)
#pragma once
#include "ldpc_packed_math.h"
#include <type_traits>

namespace pusch_dec {

struct grp_desc {
  uint32_t cb[4];     ///< indices into the batch's cb_desc array
  uint32_t n;         ///< number of valid entries (1..4)
  uint32_t layer_cap; ///< layers processed (max over the members' bounds)
  uint32_t pad[2];
};

struct dec4_layout {
  uint32_t tab_off, soft_off, c2v_off, hb_off, crc_off, misc_off, total;
};

/// Columns of tensor memory one thread needs for its messages of layers [0, layer_cap): one 32-bit column per lifted
/// edge (the message bytes of the four code blocks), every layer padded to a multiple of four columns (the granule of the
/// tcgen05.ld / tcgen05.st shapes used).
__host__ __device__ inline uint32_t dec4_tmem_cols(uint32_t bg, uint32_t layer_cap)
{
  uint32_t n = 0;
  for (uint32_t l = 0; l != layer_cap; ++l) {
#ifdef __CUDA_ARCH__
    uint32_t deg = c_row_ptr[bg - 1][l + 1] - c_row_ptr[bg - 1][l];
#else
    const uint16_t* rp  = (bg == 1) ? NR_BG1_ROW_PTR : NR_BG2_ROW_PTR;
    uint32_t        deg = rp[l + 1] - rp[l];
#endif
    n += (deg + 3U) & ~3U;
  }
  return n;
}

/// Tensor-memory columns per warp when `tm_cols` columns are shared by the TPC / 128 warps of a lane quadrant.
__host__ __device__ inline uint32_t dec4_tmem_cols_per_warp(uint32_t tm_cols, uint32_t tpc)
{
  return (tm_cols / (tpc / 128U)) & ~3U;
}

/// Two code blocks per thread (one register): columns of tensor memory for the messages of layers [0, layers) - two
/// bytes per lifted edge, two edges per 32-bit column, every layer starting on a column of its own.
__host__ __device__ inline uint32_t dec2_tmem_cols(uint32_t bg, uint32_t layers)
{
  uint32_t n = 0;
  for (uint32_t l = 0; l != layers; ++l) {
#ifdef __CUDA_ARCH__
    uint32_t deg = c_row_ptr[bg - 1][l + 1] - c_row_ptr[bg - 1][l];
#else
    const uint16_t* rp  = (bg == 1) ? NR_BG1_ROW_PTR : NR_BG2_ROW_PTR;
    uint32_t        deg = rp[l + 1] - rp[l];
#endif
    n += (deg + 1U) / 2U;
  }
  return n;
}

/// Layers (of `layer_cap`) whose messages fit in `cols_per_warp` columns of tensor memory with two code blocks per thread;
/// the messages of the remaining layers stay in shared memory.
__host__ __device__ inline uint32_t dec2_tm_layers(uint32_t bg, uint32_t layer_cap, uint32_t cols_per_warp)
{
  uint32_t lt = layer_cap;
  while (lt != 0 && dec2_tmem_cols(bg, lt) > cols_per_warp) {
    --lt;
  }
  return lt;
}

/// `lanes` = code blocks per CTA (4 or 2). `tmem`: the messages of layers [0, tm_layers) live in tensor memory, not in
/// shared memory (tm_layers defaults to all of them).
__host__ __device__ inline dec4_layout dec4_smem_layout(uint32_t bg, uint32_t Z, uint32_t layer_cap, uint32_t lanes = 4,
                                                         bool tmem = false, uint32_t tm_layers = 0xffffffffU)
{
#ifdef __CUDA_ARCH__
  uint32_t nedges = c_row_ptr[bg - 1][layer_cap];
#else
  uint32_t nedges = ((bg == 1) ? NR_BG1_ROW_PTR : NR_BG2_ROW_PTR)[layer_cap];
#endif
  // Two-code-block form only: 32-bit words (two edges each) of the layers whose messages stay in shared memory.
  uint32_t spill_words = (tmem && tm_layers < layer_cap) ? dec2_tmem_cols(bg, layer_cap) - dec2_tmem_cols(bg, tm_layers) : 0U;
  uint32_t    Kb = (bg == 1) ? 22 : 10;
  dec4_layout l;
  l.tab_off  = 0;
  l.soft_off = (nedges * 8 + 15) & ~15U;
  l.c2v_off  = l.soft_off + (Kb + layer_cap) * Z * 2 * lanes;
  l.hb_off   = l.c2v_off + (tmem ? spill_words * Z * 4 : nedges * Z * lanes);
  l.crc_off  = l.hb_off + 4 * (Kb * Z / 32) * 4;
  l.misc_off = l.crc_off + 4 * 256 * 4;
  l.total    = l.misc_off + 128 + 4 * 64; // 32 flag / scratch words + four lane_state records
  if (tmem) {
    l.total += 112; // first message column of every layer (up to 46 + 1 entries of 16 bits) + the TMEM base address
  }
  return l;
}

/// Bulk-copy staging of the decoder inputs (forms that run one CTA per SM and have the room): the `lanes` inputs of a CTA
/// are copied from their HARQ slots into shared memory behind the layout by one cp.async.bulk each (the TMA unit's 1-D
/// path: no tensor map, 16-byte aligned source and size), completion on an mbarrier; the threads build tables, allocate
/// and clear tensor memory meanwhile and then convert from shared memory. Stride per code block, and total bytes
/// (staging + the mbarrier). Flagged to the kernel by DEC4_BULK_FLAG in its `tm_cols` argument.
constexpr uint32_t DEC4_BULK_FLAG = 0x10000U;
__host__ __device__ inline uint32_t dec4_bulk_stride(uint32_t bg, uint32_t Z, uint32_t layer_cap)
{
  return ((((bg == 1) ? 22U : 10U) + layer_cap - 2U) * Z + 15U) & ~15U;
}
__host__ __device__ inline uint32_t dec4_bulk_bytes(uint32_t bg, uint32_t Z, uint32_t layer_cap, uint32_t lanes)
{
  return lanes * dec4_bulk_stride(bg, Z, layer_cap) + 16U;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); // visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
/// `bytes` (multiple of 16) from global `src` (16-byte aligned) to shared `dst`, completion counted on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst), b = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src),
               "r"(bytes), "r"(b)
               : "memory");
}

/// One lifted check (thread j) of a layer of degree DEG for the 2 * NR code blocks of the group (NR registers of two lanes
/// per thread). `tab_row[e]` = (byte offset of the edge's variable node in the soft array, circulant shift); the soft
/// addresses are computed once and kept for the write-back (the modulo by a multiply-high: the address arithmetic issues on
/// the FMA pipe). Soft values: 4 * NR bytes per variable lift, messages: 2 * NR bytes per lifted edge.
template <int DEG, int NR>
__device__ __forceinline__ void process_check_n(uint8_t* __restrict__     soft_bytes,
                                                uint8_t* __restrict__     c2v_j,
                                                const uint2* __restrict__ tab_row,
                                                uint32_t                  j,
                                                uint32_t                  Z,
                                                uint32_t                  zmagic,
                                                uint32_t                  mult)
{
  uint32_t addr[DEG];
  if constexpr (NR == 2) {
    pk::check4<DEG> ck;
    uint32_t*       cj = reinterpret_cast<uint32_t*>(c2v_j);
    ck.begin();
#pragma unroll
    for (int e = 0; e != DEG; ++e) {
      const uint2 te = tab_row[e];
      uint32_t    k  = j + te.y;
      k = __viaddmin_u32(k, 0U - Z, k); // (j + shift) mod Z: min(k - Z, k) on unsigned values, one VIADDMNMX
      addr[e]       = te.x + k * 8;
      const uint2 s = *reinterpret_cast<const uint2*>(soft_bytes + addr[e]);
      ck.gather(e, s.x, s.y, cj[e * Z]);
    }
    ck.reduce(mult);
#pragma unroll
    for (int e = 0; e != DEG; ++e) {
      uint32_t s0, s1;
      cj[e * Z]                                        = ck.scatter(e, s0, s1);
      *reinterpret_cast<uint2*>(soft_bytes + addr[e]) = make_uint2(s0, s1);
    }
  } else {
    pk::check2<DEG> ck;
    uint16_t*       cj = reinterpret_cast<uint16_t*>(c2v_j);
    ck.begin();
#pragma unroll
    for (int e = 0; e != DEG; ++e) {
      const uint2 te = tab_row[e];
      uint32_t    k  = j + te.y;
      k = __viaddmin_u32(k, 0U - Z, k); // (j + shift) mod Z: min(k - Z, k) on unsigned values, one VIADDMNMX
      addr[e] = te.x + k * 4;
      ck.gather(e, *reinterpret_cast<const uint32_t*>(soft_bytes + addr[e]), cj[e * Z]);
    }
    ck.reduce(mult);
#pragma unroll
    for (int e = 0; e != DEG; ++e) {
      uint32_t s0;
      cj[e * Z]                                           = (uint16_t)ck.scatter(e, s0);
      *reinterpret_cast<uint32_t*>(soft_bytes + addr[e]) = s0;
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Tensor memory as the message store. The check-to-variable messages are thread-private (thread j owns lifted check j of
// every layer), 117 KB per group of four BG1 / Z = 384 / 4-layer code blocks - more than half of the shared memory the
// kernel needs, and the reason only one CTA fits on an SM. Blackwell's tensor memory (256 KB per SM, 128 lanes x 512
// 32-bit columns, otherwise idle in a kernel without tcgen05.mma) has exactly that access shape: a warp reads and writes
// the 32 lanes of its quadrant (warp % 4), one lane per thread, any run of columns per instruction (tcgen05.ld / st
// 32x32b.xN -> LDTM / STTM). Thread (warp w, lane i) keeps its messages in TMEM lane 32 (w % 4) + i, columns
// [(w / 4) * CPW, ...): one column per lifted edge (the four message bytes of the four code blocks), a whole layer is one
// or two instructions instead of DEG shared-memory loads and stores. With the messages gone, shared memory holds the soft
// values only (80 KB + 9 KB of hard bits / CRC tables) and two CTAs share an SM.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols)
{
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync()
{
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_fence_after_sync()
{
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld()
{
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st()
{
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
/// N consecutive columns of the calling thread's TMEM lane -> r[0..N) (N = 4, 8, 16). Warp-collective.
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* r)
{
  static_assert(N == 4 || N == 8 || N == 16, "unsupported shape");
  if constexpr (N == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
  } else if constexpr (N == 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
  } else {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
  }
}
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t* r)
{
  static_assert(N == 4 || N == 8 || N == 16, "unsupported shape");
  if constexpr (N == 4) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3])
                 : "memory");
  } else if constexpr (N == 8) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
  } else {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
                     taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
  }
}
/// A run of DP (multiple of 4, <= 20) columns as at most two instructions.
template <int DP>
__device__ __forceinline__ void tmem_ld_row(uint32_t taddr, uint32_t* r)
{
  if constexpr (DP == 4) {
    tmem_ld<4>(taddr, r);
  } else if constexpr (DP == 8) {
    tmem_ld<8>(taddr, r);
  } else if constexpr (DP == 12) {
    tmem_ld<8>(taddr, r);
    tmem_ld<4>(taddr + 8, r + 8);
  } else {
    static_assert(DP == 20, "unsupported row length");
    tmem_ld<16>(taddr, r);
    tmem_ld<4>(taddr + 16, r + 16);
  }
}
template <int DP>
__device__ __forceinline__ void tmem_st_row(uint32_t taddr, const uint32_t* r)
{
  if constexpr (DP == 4) {
    tmem_st<4>(taddr, r);
  } else if constexpr (DP == 8) {
    tmem_st<8>(taddr, r);
  } else if constexpr (DP == 12) {
    tmem_st<8>(taddr, r);
    tmem_st<4>(taddr + 8, r + 8);
  } else {
    static_assert(DP == 20, "unsupported row length");
    tmem_st<16>(taddr, r);
    tmem_st<4>(taddr + 16, r + 16);
  }
}

/// N (1 .. 10) consecutive columns of the calling thread's TMEM lane as a sum of power-of-two shapes. Warp-collective.
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* r)
{
  static_assert(N >= 1 && N <= 10, "unsupported run of columns");
  if constexpr (N >= 8) {
    tmem_ld<8>(taddr, r);
    if constexpr (N > 8) {
      tmem_ld_cols<N - 8>(taddr + 8, r + 8);
    }
  } else if constexpr (N >= 4) {
    tmem_ld<4>(taddr, r);
    if constexpr (N > 4) {
      tmem_ld_cols<N - 4>(taddr + 4, r + 4);
    }
  } else if constexpr (N >= 2) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
    if constexpr (N > 2) {
      tmem_ld_cols<N - 2>(taddr + 2, r + 2);
    }
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr));
  }
}
template <int N>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t* r)
{
  static_assert(N >= 1 && N <= 10, "unsupported run of columns");
  if constexpr (N >= 8) {
    tmem_st<8>(taddr, r);
    if constexpr (N > 8) {
      tmem_st_cols<N - 8>(taddr + 8, r + 8);
    }
  } else if constexpr (N >= 4) {
    tmem_st<4>(taddr, r);
    if constexpr (N > 4) {
      tmem_st_cols<N - 4>(taddr + 4, r + 4);
    }
  } else if constexpr (N >= 2) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
    if constexpr (N > 2) {
      tmem_st_cols<N - 2>(taddr + 2, r + 2);
    }
  } else {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r[0]) : "memory");
  }
}

/// One lifted check (thread j) of a layer of degree DEG for TWO code blocks (one register of two lanes), the form for code
/// blocks with many layers (low code rates, HARQ retransmissions: up to all 46 layers of base graph 1 - the state of four
/// such code blocks fits neither shared nor tensor memory). The messages (one byte per code block and lifted edge) are
/// packed two edges per 32-bit word: in tensor memory (`MSG_TMEM`: `taddr` = the calling warp's lane quadrant and the
/// first column of the layer, ceil(DEG / 2) columns) or, for the few layers that do not fit there, in shared memory
/// (`msg_smem`: the thread's word 0, the words of a check Z words apart). One CTA of Z threads per SM: the v2c values
/// and the soft addresses stay in registers between the two passes, and all DEG soft loads are in flight together.
/// `tab_row[e]` = (byte offset of the edge's variable node from `smem`, 4 x circulant shift), `j4` = 4 j, `Z4` = 4 Z.
template <int DEG, bool MSG_TMEM>
__device__ __forceinline__ void process_check_t1(uint8_t* __restrict__     smem,
                                                 uint32_t                  taddr,
                                                 uint32_t* __restrict__    msg_smem,
                                                 const uint2* __restrict__ tab_row,
                                                 uint32_t                  j4,
                                                 uint32_t                  Z4,
                                                 uint32_t                  Z,
                                                 uint32_t                  mult)
{
  constexpr int          DC = (DEG + 1) / 2;
  uint32_t               addr[DEG];
  uint32_t               sv[DEG];
  uint32_t               cw[DC];
  pk::check_lanes<DEG, 1> ck;
  if constexpr (MSG_TMEM) {
    tmem_ld_cols<DC>(taddr, cw);
  } else {
#pragma unroll
    for (int i = 0; i != DC; ++i) {
      cw[i] = msg_smem[i * Z];
    }
  }
#pragma unroll
  for (int e = 0; e != DEG; ++e) {
    const uint2 te = tab_row[e];
    uint32_t    k  = j4 + te.y;
    k = __viaddmin_u32(k, 0U - Z4, k); // 4 ((j + shift) mod Z): min(k - 4 Z, k) on unsigned values, one VIADDMNMX
    addr[e] = te.x + k;
    sv[e]   = *reinterpret_cast<const uint32_t*>(smem + addr[e]);
  }
  ck.begin();
  if constexpr (MSG_TMEM) {
    tmem_wait_ld();
  }
#pragma unroll
  for (int e = 0; e != DEG; ++e) {
    // the half c + 1152 of both lanes: 0x64 above the two message bytes of edge e (low or high half of its word)
    uint32_t c = pk::prmt2(cw[e / 2], 0x64646464U, (e & 1) ? 0x4342U : 0x4140U);
    ck.gather(e, &sv[e], &c);
  }
  ck.reduce(mult);
  uint32_t cn[DEG];
#pragma unroll
  for (int e = 0; e != DEG; ++e) {
    uint32_t sn;
    ck.scatter(e, &sn, &cn[e]);
    *reinterpret_cast<uint32_t*>(smem + addr[e]) = sn;
  }
#pragma unroll
  for (int i = 0; i != DC; ++i) {
    // message bytes = the low bytes of the lanes: (edge 2i lane 0, lane 1, edge 2i + 1 lane 0, lane 1)
    cw[i] = (2 * i + 1 < DEG) ? pk::prmt2(cn[2 * i], cn[2 * i + 1], 0x6420U) : pk::prmt2(cn[2 * i], pk::C2V_ZERO4, 0x4420U);
  }
  if constexpr (MSG_TMEM) {
    tmem_st_cols<DC>(taddr, cw);
    tmem_wait_st(); // complete before the layer barrier (see process_check_t)
  } else {
#pragma unroll
    for (int i = 0; i != DC; ++i) {
      msg_smem[i * Z] = cw[i];
    }
  }
}

/// Calls f(std::integral_constant<int, DEG>) for the degree of a base-graph row (3 .. 10 and 19; BG2's 3 .. 10).
template <typename F>
__device__ __forceinline__ void for_row_degree(int deg, F&& f)
{
  switch (deg) {
    case 3:
      f(std::integral_constant<int, 3>{});
      break;
    case 4:
      f(std::integral_constant<int, 4>{});
      break;
    case 5:
      f(std::integral_constant<int, 5>{});
      break;
    case 6:
      f(std::integral_constant<int, 6>{});
      break;
    case 7:
      f(std::integral_constant<int, 7>{});
      break;
    case 8:
      f(std::integral_constant<int, 8>{});
      break;
    case 9:
      f(std::integral_constant<int, 9>{});
      break;
    case 10:
      f(std::integral_constant<int, 10>{});
      break;
    default:
      f(std::integral_constant<int, 19>{});
      break;
  }
}

#ifndef DEC4T_QR
#define DEC4T_QR 0 ///< trailing edges of a check whose v2c values stay in registers
#endif
#ifndef DEC4T_PF
#define DEC4T_PF 3 ///< shared-memory loads issued ahead of the edge being processed
#endif
#ifndef DEC4T_LATE_WAIT
#define DEC4T_LATE_WAIT 0 ///< 1: wait for a layer's message stores at the start of the NEXT layer instead of behind them
#endif

/// One lifted check (thread j) of a layer of degree DEG for four code blocks, messages in tensor memory (`taddr` = the
/// calling warp's lane quadrant and the first column of this layer). The v2c values are not kept in registers between the
/// two passes: the first pass leaves them in the soft array IN PLACE of the soft values they were computed from (the
/// lifted variable (column, (j + shift) mod Z) is touched by this thread only during a layer), the second pass reads them
/// back and overwrites them with the new soft values. That costs one 64-bit store and load per edge and saves 38
/// registers, which is what lets two 384-thread CTAs (80 registers per thread) share an SM. The last QR edges keep their
/// v2c in registers.
/// `tab_row[e]` = (byte offset of the edge's variable node from `smem`, 8 x circulant shift), `j8` = 8 j, `Z8` = 8 Z.
template <int DEG, int QR>
__device__ __forceinline__ void process_check_t(uint8_t* __restrict__     soft_bytes,
                                                uint32_t                  taddr,
                                                const uint2* __restrict__ tab_row,
                                                uint32_t                  j8,
                                                uint32_t                  Z8,
                                                uint32_t                  mult)
{
  constexpr int    DP = (DEG + 3) & ~3;
  constexpr int    QS = (QR < DEG) ? DEG - QR : 0; // edges [0, QS) park their v2c in shared memory
  constexpr int    PF = (DEC4T_PF < DEG) ? DEC4T_PF : DEG; // loads in flight ahead of the edge being processed
  uint32_t         addr[DEG];
  uint32_t         cw[DP];
  uint32_t         qk[(DEG - QS) > 0 ? (DEG - QS) : 1][2];
  uint2            ring[PF];
  pk::check_acc<2> ck;
#if DEC4T_LATE_WAIT
  tmem_wait_st(); // the previous layer's stores (the messages read below were stored an iteration ago)
#endif
  tmem_ld_row<DP>(taddr, cw);
  // The parked v2c values share the soft array with the soft values, so the compiler must keep every load behind the
  // stores that precede it in program order: the loads of the next PF edges are therefore issued BEFORE the store of the
  // current edge (software pipeline), else every edge would wait for a full shared-memory round trip.
  auto edge_addr = [&](int e) {
    const uint2 te = tab_row[e];
    uint32_t    k  = j8 + te.y;
    k = __viaddmin_u32(k, 0U - Z8, k); // 8 ((j + shift) mod Z): min(k - 8 Z, k) on unsigned values, one VIADDMNMX
    addr[e]        = te.x + k;
  };
#pragma unroll
  for (int e = 0; e != PF; ++e) {
    edge_addr(e);
    ring[e] = *reinterpret_cast<const uint2*>(soft_bytes + addr[e]);
  }
  ck.begin();
  tmem_wait_ld();
#pragma unroll
  for (int e = 0; e != DEG; ++e) {
    const uint2 sv = ring[e % PF];
    if (e + PF < DEG) {
      edge_addr(e + PF);
      ring[e % PF] = *reinterpret_cast<const uint2*>(soft_bytes + addr[e + PF]);
    }
    uint32_t s[2] = {sv.x, sv.y};
    uint32_t c[2] = {pk::prmt2(cw[e], 0x64646464U, 0x4240U), pk::prmt2(cw[e], 0x64646464U, 0x4341U)};
    uint32_t q[2];
    ck.gather_q(s, c, q);
    if (e < QS) {
      *reinterpret_cast<uint2*>(soft_bytes + addr[e]) = make_uint2(q[0], q[1]);
    } else {
      qk[e - QS][0] = q[0];
      qk[e - QS][1] = q[1];
    }
  }
  constexpr int PQ = (PF < QS) ? PF : QS;
#pragma unroll
  for (int e = 0; e != PQ; ++e) {
    ring[e] = *reinterpret_cast<const uint2*>(soft_bytes + addr[e]);
  }
  ck.reduce(mult);
#pragma unroll
  for (int e = 0; e != DEG; ++e) {
    uint32_t q[2], sn[2], cn[2];
    if (e < QS) {
      q[0] = ring[e % PF].x;
      q[1] = ring[e % PF].y;
      if (e + PF < QS) {
        ring[e % PF] = *reinterpret_cast<const uint2*>(soft_bytes + addr[e + PF]);
      }
    } else {
      q[0] = qk[e - QS][0];
      q[1] = qk[e - QS][1];
    }
    ck.scatter_q(q, sn, cn);
    cw[e]                                           = pk::prmt2(cn[0], cn[1], 0x6240U);
    *reinterpret_cast<uint2*>(soft_bytes + addr[e]) = make_uint2(sn[0], sn[1]);
  }
#pragma unroll
  for (int e = DEG; e < DP; ++e) {
    cw[e] = pk::C2V_ZERO4;
  }
  tmem_st_row<DP>(taddr, cw);
  // The messages are read again one iteration (several barriers) later; the wait makes the stores complete before the
  // layer barrier, which is the ordering the tcgen05 memory model asks for between a store and a later load.
#if !DEC4T_LATE_WAIT
  tmem_wait_st();
#endif
}


/// Per code block state of a packed group, kept in shared memory so that it does not occupy registers in the layer loop.
struct lane_state {
  const int8_t* src;       ///< decoder input (HARQ slot)
  uint8_t*      bits_out;  ///< batch output copy (may be null)
  uint32_t*     slot_bits; ///< data bits of the HARQ slot
  uint32_t      n_load, nbits, cbi, slot, flags;
  int32_t       iters;
  uint32_t      crc_ok, live, done;
};
static_assert(sizeof(lane_state) == 64, "lane_state is 64 bytes");

/// ZT: lifting size known at compile time (0 = taken from the descriptor). The specialisation for Z = 384 - the lifting size
/// of every full-size transport block - turns the per-edge message offsets into immediates and the shift wrap into a compare.
/// Launch bound of TPC + 32 threads: the kernel is launched with TPC threads, the slack only caps the registers at
/// 152 per thread (65536 / 416 rounded down to the allocation unit) so that, next to one decoder CTA (384 x 152 registers,
/// 207 KB of shared memory), an SM still has room for one 128-thread rate-dematcher CTA of the NEXT batch: the dematcher
/// then runs in the issue slots the decoder leaves idle instead of after it.
#ifndef DEC4_LB_EXTRA
#define DEC4_LB_EXTRA 32
#endif
/// NR: registers (of two code blocks each) per thread: 2 = four code blocks per CTA, one CTA per SM; 1 = two code blocks per
/// CTA with half the shared memory, two CTAs per SM - for groups whose state does not fit four code blocks (HARQ
/// retransmissions whose combined buffer spans many layers) and for small batches (a single transport block spreads over
/// twice as many SMs).
/// TM: the messages live in tensor memory (`tm_cols` columns allocated by the CTA: 256 -> two CTAs per SM, 512 -> one),
/// the v2c values pass through the soft array, 80 registers per thread: two four-code-block CTAs (24 warps) per SM.
/// TM with NR = 1: code blocks with MANY layers (up to all 46 of base graph 1: low code rates, BASELINE config 1, HARQ
/// retransmissions), two per CTA, one CTA per SM: all 512 columns hold messages (two edges per column), the messages of the
/// last layers that do not fit stay in shared memory, the v2c values stay in registers (168 per thread).
template <int TPC, int ZT = 0, int NR = 2, int TM = 0>
__global__ void __launch_bounds__((NR == 2 && !TM) ? TPC + DEC4_LB_EXTRA : TPC, ((NR == 2 && !TM) || (NR == 1 && TM)) ? 1 : 2) ldpc_decode4_kernel(const cb_desc* __restrict__ descs,
                                                               const grp_desc* __restrict__ groups,
                                                               cb_result* __restrict__ results,
                                                               const int8_t* __restrict__ soft_base,
                                                               uint8_t* __restrict__ bits_base,
                                                               uint32_t* __restrict__ crc_flags,
                                                               uint32_t tm_cols)
{
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int       t    = threadIdx.x;
  const int       lane = t & 31;
  // Broadcast from lane 0 so that the compiler knows the warp index is warp-uniform (uniform loop bounds around ballots).
  const int       warp = __shfl_sync(0xffffffffU, t >> 5, 0);
  constexpr int   NW   = TPC / 32;
  constexpr int   NC   = 2 * NR;  // code blocks per CTA
  constexpr int   SE   = 4 * NR;  // bytes of soft values per variable lift
  constexpr int   CE   = 2 * NR;  // bytes of messages per lifted edge
  const grp_desc& g    = groups[blockIdx.x];
  const cb_desc&  d0   = descs[g.cb[0]];
  const uint32_t  Z = ZT ? (uint32_t)ZT : (uint32_t)d0.Z, bg = d0.bg, Kb = (bg == 1) ? 22 : 10, K = Kb * Z, L = g.layer_cap;
  const uint32_t  mode = d0.mode, max_it = d0.max_it, mult = d0.scale_mult;
  const int       poly = (mode == MODE_NO_CRC) ? 0 : (int)d0.crc_poly;
  const uint32_t  HBW  = K / 32;
  // NR == 1 with TM: layers whose messages live in tensor memory (the host sized the shared memory with the same rule).
  const bool      bulk    = (TM != 0) && (tm_cols & DEC4_BULK_FLAG) != 0; // inputs staged by bulk copies (see dec4_bulk_bytes)
  tm_cols &= 0xffffU;
  const uint32_t  cpw  = TM ? dec4_tmem_cols_per_warp(tm_cols, TPC) : 0U;
  const uint32_t  Lt   = (TM != 0 && NR == 1) ? dec2_tm_layers(bg, L, cpw) : L;

  const dec4_layout lay  = dec4_smem_layout(bg, Z, L, NC, TM != 0, Lt);
  uint2*            tab  = reinterpret_cast<uint2*>(smem_raw + lay.tab_off);
  uint8_t*          soft = smem_raw + lay.soft_off; // SE bytes per variable lift
  uint8_t*          c2v  = smem_raw + lay.c2v_off;  // CE bytes per lifted edge, check order (TM: unused)
  uint32_t*         hb   = reinterpret_cast<uint32_t*>(smem_raw + lay.hb_off);
  uint32_t*         tabs = reinterpret_cast<uint32_t*>(smem_raw + lay.crc_off);
  uint32_t*         misc = reinterpret_cast<uint32_t*>(smem_raw + lay.misc_off);
  lane_state*       st   = reinterpret_cast<lane_state*>(smem_raw + lay.misc_off + 128);
  // TM: first message column of every layer, then the TMEM base address the allocation returned.
  uint16_t*         lcol   = reinterpret_cast<uint16_t*>(smem_raw + lay.misc_off + 128 + 4 * 64);
  uint32_t*         tm_ptr = reinterpret_cast<uint32_t*>(smem_raw + lay.misc_off + 128 + 4 * 64 + 96);
  (void)c2v;
  (void)lcol;
  (void)tm_ptr;
  // misc: [0..3] any non-zero input, [4..7] / [8..11] any zero soft bit (even / odd check rounds), [16..16+NW) CRC shares

  // ---- per code block setup (one thread each; the state lives in shared memory, not in registers) ---------------------------
  if (t < 32) {
    misc[t] = 0;
  }
  if (t < 4) {
    const int      c      = t;
    const bool     valid  = c < (int)g.n && c < NC;
    const cb_desc& d      = descs[g.cb[valid ? c : 0]];
    const uint32_t cap_in = (Kb + L) * Z - 2 * Z;
    lane_state     ls;
    ls.src       = (d.flags & FLAG_USE_HARQ) ? soft_base + (size_t)d.slot * SOFT_STRIDE : d.llr;
    ls.bits_out  = d.bits_out;
    ls.slot_bits = reinterpret_cast<uint32_t*>(bits_base + (size_t)d.slot * BITS_STRIDE);
    ls.n_load    = valid ? min(min(d.n_in, d.scan_len), cap_in) : 0U;
    ls.nbits     = K - d.nof_filler;
    ls.cbi       = g.cb[valid ? c : 0];
    ls.slot      = d.slot;
    ls.flags     = d.flags;
    ls.iters     = -1;
    ls.crc_ok    = 0;
    ls.live      = valid ? 1U : 0U;
    ls.done      = 0; // 2 = skipped: decoded in an earlier transmission
    if (valid && (d.flags & FLAG_TRACK_CRC) && !d.new_data && crc_flags[d.slot] != 0) {
      // Already decoded in an earlier transmission: only dematched (pusch_decoder_impl.cpp:335-345).
      ls.live         = 0;
      ls.n_load       = 0;
      ls.done         = 2;
      results[ls.cbi] = {0, 1U, 0U, 2U};
    }
    st[c] = ls;
  }
  __syncthreads();
  uint8_t*       stage        = smem_raw + lay.total;
  const uint32_t stage_stride = dec4_bulk_stride(bg, Z, L);
  uint64_t*      stage_bar    = reinterpret_cast<uint64_t*>(stage + NC * stage_stride);
  if (TM != 0 && bulk && t == 0) {
    mbar_init(stage_bar, 1);
    uint32_t total = 0;
#pragma unroll
    for (int c = 0; c != NC; ++c) {
      total += (st[c].n_load + 15U) & ~15U;
    }
    mbar_expect_tx(stage_bar, total);
#pragma unroll
    for (int c = 0; c != NC; ++c) {
      const uint32_t len = (st[c].n_load + 15U) & ~15U; // (the bytes up to the 16-byte boundary belong to the same slot)
      if (len != 0) {
        bulk_g2s(stage + c * stage_stride, st[c].src, len, stage_bar);
      }
    }
  }
  {
    uint32_t any_live = 0;
#pragma unroll
    for (int c = 0; c != 4; ++c) {
      any_live |= st[c].live;
      if (st[c].done == 2 && st[c].bits_out != nullptr) {
        for (uint32_t i = t; i < HBW; i += TPC) {
          reinterpret_cast<uint32_t*>(st[c].bits_out)[i] = st[c].slot_bits[i];
        }
      }
    }
    if (any_live == 0) {
      return; // every code block of the group was already decoded (retransmission of an acknowledged TB)
    }
  }

  // ---- prologue ----------------------------------------------------------------------------------------------------------
  const uint32_t nedges = c_row_ptr[bg - 1][L];
  for (uint32_t e = t; e < nedges; e += TPC) {
    if constexpr (TM != 0) {
      // absolute offset in the CTA's shared memory, shift in bytes: three instructions from here to the address
      tab[e] = make_uint2(lay.soft_off + (uint32_t)c_col[bg - 1][e] * Z * SE, (c_shift[bg - 1][d0.ils][e] % Z) * SE);
    } else {
      tab[e] = make_uint2((uint32_t)c_col[bg - 1][e] * Z * SE, c_shift[bg - 1][d0.ils][e] % Z);
    }
  }
  if constexpr (TM != 0) {
    // Tensor memory for the messages (every thread of the CTA returns together above, so the allocation is always paired
    // with the release at the end). Two co-resident CTAs take 256 columns each; nothing else on the SM allocates.
    if (warp == 0) {
      tmem_alloc(tm_ptr, tm_cols);
    }
    if (t <= (int)L) {
      lcol[t] = (uint16_t)((NR == 2) ? dec4_tmem_cols(bg, (uint32_t)t) : dec2_tmem_cols(bg, (uint32_t)t));
    }
    if constexpr (NR == 1) {
      // Messages of the layers kept in shared memory: zero (byte 0x80 per code block and edge).
      uint4*         c4 = reinterpret_cast<uint4*>(c2v);
      const uint32_t n4 = (lay.hb_off - lay.c2v_off) / 16;
      const uint4    zz = make_uint4(pk::C2V_ZERO4, pk::C2V_ZERO4, pk::C2V_ZERO4, pk::C2V_ZERO4);
      for (uint32_t i = t; i < n4; i += TPC) {
        c4[i] = zz;
      }
    }
  }
  {
    // Input of the code blocks, 16 variable nodes per code block and step (Z % 16 == 0): the 128-bit loads of a
    // step are all in flight together, the messages are cleared while they travel.
    const uint32_t n16   = (Kb + L) * Z / 16;
    const uint32_t punct = 2 * Z / 16;
    uint32_t       nz[NC];
    const int8_t*  src[NC];
    uint32_t       n_load[NC];
#pragma unroll
    for (int c = 0; c != NC; ++c) {
      nz[c]     = 0;
      src[c]    = st[c].src;
      n_load[c] = st[c].n_load;
    }
    auto fetch = [&](uint32_t v, uint4* w) {
#pragma unroll
      for (int c = 0; c != NC; ++c) {
        w[c] = make_uint4(0, 0, 0, 0);
      }
      if (v >= punct && v < n16) {
        const uint32_t p = (v - punct) * 16;
        if (TM != 0 && bulk) {
          // staged by the bulk copies: the same bytes from shared memory (whole 16-byte units were copied)
#pragma unroll
          for (int c = 0; c != NC; ++c) {
            if (p < n_load[c]) {
              w[c] = *reinterpret_cast<const uint4*>(stage + c * stage_stride + p);
              if (p + 16 > n_load[c]) {
                w[c].y = (p + 4 < n_load[c]) ? w[c].y : 0U;
                w[c].z = (p + 8 < n_load[c]) ? w[c].z : 0U;
                w[c].w = (p + 12 < n_load[c]) ? w[c].w : 0U;
              }
            }
          }
          return;
        }
#pragma unroll
        for (int c = 0; c != NC; ++c) {
          if (p + 16 <= n_load[c]) {
            w[c] = __ldg(reinterpret_cast<const uint4*>(src[c] + p));
          } else if (p < n_load[c]) {
            // last, partial step: 4-byte granules (the bytes of the slot beyond the input are zero, see cb_desc::scan_len)
            const uint32_t* q = reinterpret_cast<const uint32_t*>(src[c] + p);
            w[c].x = __ldg(q);
            w[c].y = (p + 4 < n_load[c]) ? __ldg(q + 1) : 0U;
            w[c].z = (p + 8 < n_load[c]) ? __ldg(q + 2) : 0U;
            w[c].w = (p + 12 < n_load[c]) ? __ldg(q + 3) : 0U;
          }
        }
      }
    };
    uint4 w[NC];
    if (!(TM != 0 && bulk)) {
      fetch(t, w);
    }
    if constexpr (TM == 0) {
      uint4*         c4 = reinterpret_cast<uint4*>(c2v);
      const uint32_t n4 = nedges * Z * CE / 16;
      const uint4    zz = make_uint4(pk::C2V_ZERO4, pk::C2V_ZERO4, pk::C2V_ZERO4, pk::C2V_ZERO4);
      for (uint32_t i = t; i < n4; i += TPC) {
        c4[i] = zz;
      }
    }
    if (poly != 0) {
      build_crc_tables(tabs, poly, t, TPC);
    }
    if (TM != 0 && bulk) {
      __syncthreads(); // the mbarrier was initialised by thread 0 after the previous barrier
      mbar_wait(stage_bar, 0);
      fetch(t, w);
    }
    for (uint32_t v = t; v < n16; v += TPC) {
      uint32_t ww[NC][4];
#pragma unroll
      for (int c = 0; c != NC; ++c) {
        nz[c] |= w[c].x | w[c].y | w[c].z | w[c].w;
        ww[c][0] = w[c].x ^ 0x80808080U;
        ww[c][1] = w[c].y ^ 0x80808080U;
        ww[c][2] = w[c].z ^ 0x80808080U;
        ww[c][3] = w[c].w ^ 0x80808080U;
      }
      fetch(v + TPC, w); // next step's loads travel while this one is converted
#pragma unroll
      for (int i = 0; i != 4; ++i) {
        // selb: byte b of the first operand -> bytes 0,1; of the second -> bytes 2,3 (then masked to the lane's low byte)
        if constexpr (NR == 2) {
#pragma unroll
          for (int b = 0; b != 4; b += 2) {
            uint32_t sel0 = (uint32_t)b * 0x11U + 0x4400U + (uint32_t)b * 0x1100U;
            uint32_t sel1 = sel0 + 0x1111U;
            uint4    o;
            o.x = pk::soft_from_biased_bytes(__byte_perm(ww[0][i], ww[2][i], sel0) & 0x00ff00ffU);
            o.y = pk::soft_from_biased_bytes(__byte_perm(ww[1][i], ww[3][i], sel0) & 0x00ff00ffU);
            o.z = pk::soft_from_biased_bytes(__byte_perm(ww[0][i], ww[2][i], sel1) & 0x00ff00ffU);
            o.w = pk::soft_from_biased_bytes(__byte_perm(ww[1][i], ww[3][i], sel1) & 0x00ff00ffU);
            *reinterpret_cast<uint4*>(soft + (size_t)(v * 16 + i * 4 + b) * SE) = o;
          }
        } else {
          uint4 o;
          o.x = pk::soft_from_biased_bytes(__byte_perm(ww[0][i], ww[1][i], 0x4400U) & 0x00ff00ffU);
          o.y = pk::soft_from_biased_bytes(__byte_perm(ww[0][i], ww[1][i], 0x5511U) & 0x00ff00ffU);
          o.z = pk::soft_from_biased_bytes(__byte_perm(ww[0][i], ww[1][i], 0x6622U) & 0x00ff00ffU);
          o.w = pk::soft_from_biased_bytes(__byte_perm(ww[0][i], ww[1][i], 0x7733U) & 0x00ff00ffU);
          *reinterpret_cast<uint4*>(soft + (size_t)(v * 16 + i * 4) * SE) = o;
        }
      }
    }
#pragma unroll
    for (int c = 0; c != NC; ++c) {
      uint32_t any = __reduce_or_sync(0xffffffffU, nz[c]);
      if (lane == 0 && any != 0) {
        atomicOr(&misc[c], 1U);
      }
    }
  }
  if constexpr (TM != 0) {
    tmem_fence_before_sync();
  }
  __syncthreads();
  uint32_t tm_warp = 0; // TMEM address of this warp's lane quadrant and column range
  if constexpr (TM != 0) {
    tmem_fence_after_sync();
    tm_warp = *tm_ptr + ((uint32_t)(warp & 3) << 21) + (uint32_t)(warp >> 2) * cpw;
    // Messages of the first iteration: zero (byte 0x80 per code block).
    uint32_t zz[4] = {pk::C2V_ZERO4, pk::C2V_ZERO4, pk::C2V_ZERO4, pk::C2V_ZERO4};
    const uint32_t used = lcol[Lt]; // (rounded up to four columns below: cpw is a multiple of four)
    for (uint32_t col = 0; col < used; col += 4) {
      tmem_st<4>(tm_warp + col, zz);
    }
    tmem_wait_st();
  }
  // Code blocks that are not decoded: not live, or all-zero input with early stop (the reference returns before touching
  // the output, ldpc_decoder_impl.cpp:88-94). Every thread keeps the same mask of finished code blocks.
  uint32_t done_mask = 0;
#pragma unroll
  for (int c = 0; c != 4; ++c) {
    const bool zero_in = st[c].live && misc[c] == 0 && mode == MODE_EARLY_STOP;
    if (!st[c].live || zero_in) {
      done_mask |= 1U << c;
    }
    if (zero_in && st[c].bits_out != nullptr) {
      for (uint32_t i = t; i < HBW; i += TPC) {
        reinterpret_cast<uint32_t*>(st[c].bits_out)[i] = st[c].slot_bits[i]; // untouched output
      }
    }
  }

  // ---- iterations --------------------------------------------------------------------------------------------------------
  constexpr int  WPC    = NW / NC; // warps per code block in the CRC step
  const uint32_t j      = t;
  const uint32_t zmagic = 0xffffffffU / Z + 1; // ceil(2^32 / Z): floor(k / Z) = umulhi(k, zmagic) for k < 2 Z
  uint32_t       checks = 0;                   // hard-decision rounds so far (parity selects the "any zero" flag set)
  for (uint32_t it = 0; it != max_it && done_mask != 0xfU; ++it) {
    for (uint32_t l = 0; l != L; ++l) {
      uint32_t e0  = c_row_ptr[bg - 1][l];
      int      deg = (int)c_row_ptr[bg - 1][l + 1] - (int)e0;
      if constexpr (TM != 0 && NR == 1) {
        if (j < Z) {
          const uint2* tab_row = tab + e0;
          if (l < Lt) {
            const uint32_t ta = tm_warp + lcol[l];
            for_row_degree(deg, [&](auto dg) {
              process_check_t1<decltype(dg)::value, true>(smem_raw, ta, nullptr, tab_row, j * SE, Z * SE, Z, mult);
            });
          } else {
            uint32_t* mw = reinterpret_cast<uint32_t*>(c2v) + (size_t)(lcol[l] - lcol[Lt]) * Z + j;
            for_row_degree(deg, [&](auto dg) {
              process_check_t1<decltype(dg)::value, false>(smem_raw, 0U, mw, tab_row, j * SE, Z * SE, Z, mult);
            });
          }
        }
      } else if constexpr (TM != 0) {
        // Z is a multiple of 32 here: whole warps are in or out (tcgen05.ld / st are warp-collective).
        if (j < Z) {
          const uint2*   tab_row = tab + e0;
          const uint32_t ta      = tm_warp + lcol[l];
          switch (deg) {
            case 3:
              process_check_t<3, DEC4T_QR>(smem_raw, ta, tab_row, j * SE, Z * SE, mult);
              break;
            case 4:
              process_check_t<4, DEC4T_QR>(smem_raw, ta, tab_row, j * SE, Z * SE, mult);
              break;
            case 5:
              process_check_t<5, DEC4T_QR>(smem_raw, ta, tab_row, j * SE, Z * SE, mult);
              break;
            case 6:
              process_check_t<6, DEC4T_QR>(smem_raw, ta, tab_row, j * SE, Z * SE, mult);
              break;
            case 7:
              process_check_t<7, DEC4T_QR>(smem_raw, ta, tab_row, j * SE, Z * SE, mult);
              break;
            case 8:
              process_check_t<8, DEC4T_QR>(smem_raw, ta, tab_row, j * SE, Z * SE, mult);
              break;
            case 9:
              process_check_t<9, DEC4T_QR>(smem_raw, ta, tab_row, j * SE, Z * SE, mult);
              break;
            case 10:
              process_check_t<10, DEC4T_QR>(smem_raw, ta, tab_row, j * SE, Z * SE, mult);
              break;
            default:
              process_check_t<19, DEC4T_QR>(smem_raw, ta, tab_row, j * SE, Z * SE, mult);
              break;
          }
        }
      } else if (j < Z) {
        uint8_t*     c2v_row = c2v + ((size_t)e0 * Z + j) * CE;
        const uint2* tab_row = tab + e0;
        uint8_t*     sb      = soft;
        switch (deg) {
          case 3:
            process_check_n<3, NR>(sb, c2v_row, tab_row, j, Z, zmagic, mult);
            break;
          case 4:
            process_check_n<4, NR>(sb, c2v_row, tab_row, j, Z, zmagic, mult);
            break;
          case 5:
            process_check_n<5, NR>(sb, c2v_row, tab_row, j, Z, zmagic, mult);
            break;
          case 6:
            process_check_n<6, NR>(sb, c2v_row, tab_row, j, Z, zmagic, mult);
            break;
          case 7:
            process_check_n<7, NR>(sb, c2v_row, tab_row, j, Z, zmagic, mult);
            break;
          case 8:
            process_check_n<8, NR>(sb, c2v_row, tab_row, j, Z, zmagic, mult);
            break;
          case 9:
            process_check_n<9, NR>(sb, c2v_row, tab_row, j, Z, zmagic, mult);
            break;
          case 10:
            process_check_n<10, NR>(sb, c2v_row, tab_row, j, Z, zmagic, mult);
            break;
          default:
            process_check_n<19, NR>(sb, c2v_row, tab_row, j, Z, zmagic, mult);
            break;
        }
      }
      __syncthreads();
    }
    const bool last_it = (it + 1 == max_it);
    if (mode != MODE_EARLY_STOP && !last_it) {
      continue;
    }
    // Hard decision of the first K soft bits of every code block: bit = (llr <= 0), MSB first; "any zero" flags.
    uint32_t* zflag = misc + 4 + 4 * (checks & 1U);
    {
      // Lane l takes variable 32 w + 31 - l: ballot bit l is then already in MSB-first order. The half (soft + 1152) - 1153
      // is negative exactly for soft <= 0; a running unsigned minimum of pattern ^ pattern(0) finds zero soft values.
      uint32_t z0 = 0xffffffffU, z1 = 0xffffffffU;
      for (uint32_t w = warp; w < HBW; w += NW) {
        if constexpr (NR == 2) {
          const uint2    s  = *reinterpret_cast<const uint2*>(soft + (size_t)(w * 32 + 31 - lane) * SE);
          const uint32_t d0 = pk::hadd2(s.x, PK_REP2(0xe481U)), d1 = pk::hadd2(s.y, PK_REP2(0xe481U));
          z0 = pk::minu2(z0, s.x ^ pk::SOFT_ZERO2);
          z1 = pk::minu2(z1, s.y ^ pk::SOFT_ZERO2);
          // (the AND is opaque to the optimiser, which would turn the test into shift + and + compare: 3 instructions)
          uint32_t l0, l1;
          asm("and.b32 %0, %1, 0x8000;" : "=r"(l0) : "r"(d0));
          asm("and.b32 %0, %1, 0x8000;" : "=r"(l1) : "r"(d1));
          const uint32_t b0 = __ballot_sync(0xffffffffU, l0 != 0);
          const uint32_t b1 = __ballot_sync(0xffffffffU, l1 != 0);
          const uint32_t b2 = __ballot_sync(0xffffffffU, (int32_t)d0 < 0);
          const uint32_t b3 = __ballot_sync(0xffffffffU, (int32_t)d1 < 0);
          if (lane < 4) {
            const uint32_t blo = (lane & 1) ? b1 : b0, bhi = (lane & 1) ? b3 : b2;
            hb[lane * HBW + w] = (lane & 2) ? bhi : blo;
          }
        } else {
          const uint32_t s  = *reinterpret_cast<const uint32_t*>(soft + (size_t)(w * 32 + 31 - lane) * SE);
          const uint32_t d0 = pk::hadd2(s, PK_REP2(0xe481U));
          z0                = pk::minu2(z0, s ^ pk::SOFT_ZERO2);
          uint32_t l0;
          asm("and.b32 %0, %1, 0x8000;" : "=r"(l0) : "r"(d0));
          const uint32_t b0 = __ballot_sync(0xffffffffU, l0 != 0);
          const uint32_t b1 = __ballot_sync(0xffffffffU, (int32_t)d0 < 0);
          if (lane < 2) {
            hb[lane * HBW + w] = lane ? b1 : b0;
          }
        }
      }
      z0 = pk::minu2(z0, __shfl_xor_sync(0xffffffffU, z0, 16));
      z0 = pk::minu2(z0, __shfl_xor_sync(0xffffffffU, z0, 8));
      z0 = pk::minu2(z0, __shfl_xor_sync(0xffffffffU, z0, 4));
      z0 = pk::minu2(z0, __shfl_xor_sync(0xffffffffU, z0, 2));
      z0 = pk::minu2(z0, __shfl_xor_sync(0xffffffffU, z0, 1));
      if constexpr (NR == 2) {
        z1 = pk::minu2(z1, __shfl_xor_sync(0xffffffffU, z1, 16));
        z1 = pk::minu2(z1, __shfl_xor_sync(0xffffffffU, z1, 8));
        z1 = pk::minu2(z1, __shfl_xor_sync(0xffffffffU, z1, 4));
        z1 = pk::minu2(z1, __shfl_xor_sync(0xffffffffU, z1, 2));
        z1 = pk::minu2(z1, __shfl_xor_sync(0xffffffffU, z1, 1));
      }
      if (lane == 0) {
        // code block of a lane: NR == 2: (s.x low, s.y low, s.x high, s.y high) = 0..3; NR == 1: (low, high) = 0, 1
        if ((z0 & 0xffffU) == 0) {
          atomicOr(&zflag[0], 1U);
        }
        if ((z0 >> 16) == 0) {
          atomicOr(&zflag[NR == 2 ? 2 : 1], 1U);
        }
        if (NR == 2 && (z1 & 0xffffU) == 0) {
          atomicOr(&zflag[1], 1U);
        }
        if (NR == 2 && (z1 >> 16) == 0) {
          atomicOr(&zflag[3], 1U);
        }
      }
    }
    __syncthreads();
    {
      // CRC of every unfinished code block by WPC warps; the shares are combined by every thread after the barrier.
      const int c     = warp / WPC;
      uint32_t  share = 0;
      if (poly != 0 && !((done_mask >> c) & 1U)) {
        share = crc_share_words(hb + c * HBW, st[c].nbits, poly, tabs, (uint32_t)(warp % WPC) * 32 + lane, WPC * 32);
      }
      share = __reduce_xor_sync(0xffffffffU, share);
      if (lane == 0) {
        misc[16 + warp] = share;
      }
    }
    __syncthreads();
    uint32_t newly = 0;
#pragma unroll
    for (int c = 0; c != 4; ++c) {
      if ((done_mask >> c) & 1U) {
        continue;
      }
      uint32_t crc = 0;
#pragma unroll
      for (int k = 0; k != WPC; ++k) {
        crc ^= misc[16 + c * WPC + k];
      }
      const bool ok = poly != 0 && crc == 0 && (mode != MODE_EARLY_STOP || zflag[c] == 0);
      if (ok || last_it) {
        // The output holds the hard decision of the last iteration run (ldpc_decoder_impl.cpp:126-146).
        uint32_t* slot = st[c].slot_bits;
        uint32_t* out  = reinterpret_cast<uint32_t*>(st[c].bits_out);
        for (uint32_t i = t; i < HBW; i += TPC) {
          uint32_t wd = __byte_perm(hb[c * HBW + i], 0, 0x0123);
          slot[i]     = wd;
          if (out != nullptr) {
            out[i] = wd;
          }
        }
      }
      if (ok) {
        newly |= 1U << c;
      }
    }
    if (t < 4) {
      if ((newly >> t) & 1U) {
        st[t].crc_ok = 1;
        st[t].iters  = (mode == MODE_EARLY_STOP) ? (int)it + 1 : (int)max_it;
      }
      misc[4 + 4 * ((checks + 1) & 1U) + t] = 0; // "any zero" flags of the next round (last read two barriers ago)
    }
    done_mask |= newly;
    ++checks;
  }

  if (t < 4 && st[t].live) {
    results[st[t].cbi] = {st[t].iters, st[t].crc_ok, L, 0U};
    if (st[t].flags & FLAG_TRACK_CRC) {
      crc_flags[st[t].slot] = st[t].crc_ok;
    }
  }
  if constexpr (TM != 0) {
    tmem_wait_st();
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
      tmem_dealloc(*tm_ptr, tm_cols);
    }
  }
}

} // namespace pusch_dec

namespace pusch_dec {

// ---------------------------------------------------------------------------------------------------------------------
// Kernel 2d: ONE code block per CTA on the packed arithmetic - the four 16-bit lanes hold four lifted checks of the same
// code block (j, j + Z/4, j + Z/2, j + 3Z/4; see "intra-code-block packing" in ldpc_packed_math.h), Z/4 threads per
// code block. No grouping constraint: ragged batches (any mix of base graph, lifting size with Z % 4 == 0, layers,
// modes), every code block early-stops on its own, and up to four CTAs share an SM so the serial phases of one (input
// load, CRC) overlap the layer sweeps of the others. Shared memory per code block: 2 bytes per variable lift + 1 byte per
// lifted edge (BASELINE config 2: 51 KB; config 1, 46 layers: 173 KB).
// ---------------------------------------------------------------------------------------------------------------------
struct decq_layout {
  uint32_t tab_off, soft_off, c2v_off, hb_off, crc_off, misc_off, total;
};

__host__ __device__ inline decq_layout decq_smem_layout(uint32_t bg, uint32_t Z, uint32_t layer_cap)
{
#ifdef __CUDA_ARCH__
  uint32_t nedges = c_row_ptr[bg - 1][layer_cap];
#else
  uint32_t nedges = ((bg == 1) ? NR_BG1_ROW_PTR : NR_BG2_ROW_PTR)[layer_cap];
#endif
  uint32_t    Kb = (bg == 1) ? 22 : 10;
  decq_layout l;
  l.tab_off  = 0;
  l.soft_off = (nedges * 8 + 15) & ~15U;
  l.c2v_off  = l.soft_off + (((Kb + layer_cap) * Z * 2 + 15) & ~15U);
  l.hb_off   = l.c2v_off + ((nedges * Z + 15) & ~15U);
  l.crc_off  = l.hb_off + (((Kb * Z + 31) / 32 * 4 + 4 + 15) & ~15U);
  l.misc_off = l.crc_off + 4 * 256 * 4;
  l.total    = l.misc_off + 128;
  return l;
}

/// Selector pairs for pk::rot4 by q quarter turns (entries 0..3) and back (entries 4..7 = rotation by (4 - q) mod 4).
__constant__ uint2 c_rot4[8] = {{0x3210U, 0x7654U}, {0x7654U, 0x1032U}, {0x1032U, 0x5476U}, {0x5476U, 0x3210U},
                                {0x3210U, 0x7654U}, {0x5476U, 0x3210U}, {0x1032U, 0x5476U}, {0x7654U, 0x1032U}};

template <int DEG>
__device__ __forceinline__ void process_check_q4(uint2* __restrict__       soft,
                                                 uint32_t* __restrict__    c2v_row,
                                                 const uint2* __restrict__ tab_row,
                                                 const uint2* __restrict__ rot,
                                                 uint32_t                  j,
                                                 uint32_t                  Z,
                                                 uint32_t                  Z4,
                                                 uint32_t                  zmagic,
                                                 uint32_t                  z4magic,
                                                 uint32_t                  mult)
{
  pk::check4<DEG> ck;
  uint32_t        aq[DEG]; // soft-buffer index << 2 | quarter turns, kept for the write-back
  ck.begin();
#pragma unroll
  for (int e = 0; e != DEG; ++e) {
    const uint2 te = tab_row[e];            // (column * Z/4, shift)
    uint32_t    k  = j + te.y;
    k -= __umulhi(k, zmagic) * Z;
    const uint32_t q  = __umulhi(k, z4magic); // quarter of the column the first check lands in
    const uint32_t ad = te.x + k - q * Z4;
    aq[e]             = (ad << 2) | q;
    const uint2 s     = soft[ad];
    const uint2 r     = rot[q];
    ck.gather(e, __byte_perm(s.x, s.y, r.x), __byte_perm(s.x, s.y, r.y), c2v_row[e * Z4 + j]);
  }
  ck.reduce(mult);
#pragma unroll
  for (int e = 0; e != DEG; ++e) {
    uint32_t s0, s1;
    c2v_row[e * Z4 + j] = ck.scatter(e, s0, s1);
    const uint2 r       = rot[4 + (aq[e] & 3U)];
    soft[aq[e] >> 2]    = make_uint2(__byte_perm(s0, s1, r.x), __byte_perm(s0, s1, r.y));
  }
}

template <int TPB>
__global__ void __launch_bounds__(TPB, (TPB == 96) ? 4 : 6) ldpc_decode_q4_kernel(const cb_desc* __restrict__ descs,
                                                                                   const uint32_t* __restrict__ order,
                                                                                   cb_result* __restrict__ results,
                                                                                   const int8_t* __restrict__ soft_base,
                                                                                   uint8_t* __restrict__ bits_base,
                                                                                   uint32_t* __restrict__ crc_flags)
{
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int      t    = threadIdx.x;
  const int      lane = t & 31;
  const int      warp = __shfl_sync(0xffffffffU, t >> 5, 0);
  const uint32_t cb   = order[blockIdx.x];
  const cb_desc& d    = descs[cb];
  const uint32_t Z = d.Z, bg = d.bg, Kb = (bg == 1) ? 22 : 10, K = Kb * Z, Z4 = Z / 4;
  const uint32_t HBW  = (K + 31) / 32;
  uint32_t*      slot_bits = reinterpret_cast<uint32_t*>(bits_base + (size_t)d.slot * BITS_STRIDE);

  // Code blocks whose CRC is already ok are only dematched (pusch_decoder_impl.cpp:335-345).
  if ((d.flags & FLAG_TRACK_CRC) && !d.new_data && crc_flags[d.slot] != 0) {
    if (t == 0) {
      results[cb] = {0, 1U, 0U, 2U};
    }
    if (d.bits_out != nullptr) {
      for (uint32_t i = t; i < HBW; i += TPB) {
        reinterpret_cast<uint32_t*>(d.bits_out)[i] = slot_bits[i];
      }
    }
    return;
  }

  const uint32_t    Lcap = d.layer_cap;
  const decq_layout lay  = decq_smem_layout(bg, Z, Lcap);
  uint2*            tab  = reinterpret_cast<uint2*>(smem_raw + lay.tab_off);
  uint2*            soft = reinterpret_cast<uint2*>(smem_raw + lay.soft_off);
  uint32_t*         c2v  = reinterpret_cast<uint32_t*>(smem_raw + lay.c2v_off);
  uint32_t*         hb   = reinterpret_cast<uint32_t*>(smem_raw + lay.hb_off);
  uint32_t*         tabs = reinterpret_cast<uint32_t*>(smem_raw + lay.crc_off);
  uint32_t*         misc = reinterpret_cast<uint32_t*>(smem_raw + lay.misc_off);
  uint2*            rot  = reinterpret_cast<uint2*>(smem_raw + lay.misc_off + 64);
  // misc: [0] index after the last non-zero input LLR, [1] any zero soft bit, [2] crc ok

  const uint32_t mode = d.mode, max_it = d.max_it, mult = d.scale_mult;
  const int      poly = (mode == MODE_NO_CRC) ? 0 : (int)d.crc_poly;
  const uint32_t nb   = K - d.nof_filler;
  const uint32_t zmagic = 0xffffffffU / Z + 1, z4magic = 0xffffffffU / Z4 + 1;

  // ---- prologue ----------------------------------------------------------------------------------------------------------
  if (t < 8) {
    rot[t] = c_rot4[t];
  }
  if (t < 16) {
    misc[t] = 0;
  }
  const uint32_t nedges_cap = c_row_ptr[bg - 1][Lcap];
  for (uint32_t e = t; e < nedges_cap; e += TPB) {
    tab[e] = make_uint2((uint32_t)c_col[bg - 1][e] * Z4, c_shift[bg - 1][d.ils][e] % Z);
  }
  for (uint32_t i = t; i < nedges_cap * Z4; i += TPB) {
    c2v[i] = pk::C2V_ZERO4;
  }
  if (poly != 0) {
    build_crc_tables(tabs, poly, t, TPB);
  }
  __syncthreads();
  {
    const int8_t*  src    = (d.flags & FLAG_USE_HARQ) ? soft_base + (size_t)d.slot * SOFT_STRIDE : d.llr;
    const uint32_t cap_in = (Kb + Lcap) * Z - 2 * Z;
    const uint32_t n_load = min(min(d.n_in, d.scan_len), cap_in);
    const uint32_t ncols  = Kb + Lcap;
    uint32_t       last   = 0;
    if ((Z4 & 3U) == 0 && (reinterpret_cast<uintptr_t>(src) & 3U) == 0) {
      // Four consecutive bases per thread: one 32-bit load per quarter of the column, transposed in registers.
      const uint32_t qpc   = Z4 / 4; // quads per column
      const uint32_t nquad = ncols * qpc;
      const uint32_t qmag  = 0xffffffffU / qpc + 1;
      for (uint32_t qi = t; qi < nquad; qi += TPB) {
        const uint32_t c  = (qpc == 1) ? qi : __umulhi(qi, qmag); // (the magic constant overflows for a divisor of 1)
        const uint32_t b  = (qi - c * qpc) * 4;
        uint32_t       w[4];
#pragma unroll
        for (uint32_t m = 0; m != 4; ++m) {
          const int idx = (int)(c * Z + m * Z4 + b) - (int)(2 * Z); // index into the decoder input
          w[m]          = 0;
          if (idx >= 0 && (uint32_t)idx < n_load) {
            w[m] = __ldg(reinterpret_cast<const uint32_t*>(src + idx));
            if ((uint32_t)idx + 4 > n_load) {
              w[m] &= 0xffffffffU >> (8 * ((uint32_t)idx + 4 - n_load)); // input ends inside this word
            }
            if (w[m] != 0) {
              last = max(last, (uint32_t)idx + 4 - (__clz(w[m]) >> 3));
            }
          }
          w[m] ^= 0x80808080U;
        }
        uint4* dst = reinterpret_cast<uint4*>(soft + c * Z4 + b);
        uint2  o[4];
#pragma unroll
        for (int bb = 0; bb != 4; ++bb) {
          uint32_t selb = (uint32_t)bb * 0x11U + 0x4400U + (uint32_t)bb * 0x1100U; // byte bb of x -> byte 0, of y -> byte 2
          o[bb] = make_uint2(pk::soft_from_biased_bytes(__byte_perm(w[0], w[2], selb) & 0x00ff00ffU),
                             pk::soft_from_biased_bytes(__byte_perm(w[1], w[3], selb) & 0x00ff00ffU));
        }
        dst[0] = make_uint4(o[0].x, o[0].y, o[1].x, o[1].y);
        dst[1] = make_uint4(o[2].x, o[2].y, o[3].x, o[3].y);
      }
    } else {
      uint32_t c = 0, b = t; // word (c, b), b < Z4
      while (b >= Z4) {
        b -= Z4;
        ++c;
      }
      while (c < ncols) {
        uint32_t ub[4];
#pragma unroll
        for (uint32_t m = 0; m != 4; ++m) {
          uint32_t i = c * Z + b + m * Z4; // variable index; decoder input index is i - 2 Z
          int      x = 0;
          if (i >= 2 * Z && i - 2 * Z < n_load) {
            x = __ldg(src + (i - 2 * Z));
            if (x != 0) {
              last = max(last, i - 2 * Z + 1);
            }
          }
          ub[m] = (uint32_t)(uint8_t)(x ^ 0x80);
        }
        soft[c * Z4 + b] =
            make_uint2(pk::soft_from_biased_bytes(ub[0] | (ub[2] << 16)), pk::soft_from_biased_bytes(ub[1] | (ub[3] << 16)));
        b += TPB;
        while (b >= Z4) {
          b -= Z4;
          ++c;
        }
      }
    }
    last = __reduce_max_sync(0xffffffffU, last);
    if (lane == 0 && last != 0) {
      atomicMax(&misc[0], last);
    }
  }
  __syncthreads();
  const uint32_t last = misc[0];
  // Layers actually processed (ldpc_decoder_impl.cpp:86-114); never more than the host's bound.
  uint32_t L = 0;
  if (last != 0) {
    uint32_t cbl = max(last + 2 * Z, K + 4 * Z);
    L            = min((cbl + Z - 1) / Z - Kb, Lcap);
  }

  int      iters  = -1;
  uint32_t crc_ok = 0;
  // All-zero input (ldpc_decoder_impl.cpp:88-94): with a CRC calculator the reference returns before touching the output;
  // without one the output becomes all ones (the hard decision of the untouched zero soft bits).
  const bool skip_all = (last == 0 && mode == MODE_EARLY_STOP);
  if (skip_all && d.bits_out != nullptr) {
    for (uint32_t i = t; i < HBW; i += TPB) {
      reinterpret_cast<uint32_t*>(d.bits_out)[i] = slot_bits[i]; // untouched output
    }
  }
  const uint32_t j = t;
  for (uint32_t it = 0; it != max_it && !skip_all; ++it) {
    for (uint32_t l = 0; l != L; ++l) {
      uint32_t e0  = c_row_ptr[bg - 1][l];
      int      deg = (int)c_row_ptr[bg - 1][l + 1] - (int)e0;
      if (j < Z4) {
        uint32_t*    c2v_row = c2v + (size_t)e0 * Z4;
        const uint2* tab_row = tab + e0;
        switch (deg) {
          case 3:
            process_check_q4<3>(soft, c2v_row, tab_row, rot, j, Z, Z4, zmagic, z4magic, mult);
            break;
          case 4:
            process_check_q4<4>(soft, c2v_row, tab_row, rot, j, Z, Z4, zmagic, z4magic, mult);
            break;
          case 5:
            process_check_q4<5>(soft, c2v_row, tab_row, rot, j, Z, Z4, zmagic, z4magic, mult);
            break;
          case 6:
            process_check_q4<6>(soft, c2v_row, tab_row, rot, j, Z, Z4, zmagic, z4magic, mult);
            break;
          case 7:
            process_check_q4<7>(soft, c2v_row, tab_row, rot, j, Z, Z4, zmagic, z4magic, mult);
            break;
          case 8:
            process_check_q4<8>(soft, c2v_row, tab_row, rot, j, Z, Z4, zmagic, z4magic, mult);
            break;
          case 9:
            process_check_q4<9>(soft, c2v_row, tab_row, rot, j, Z, Z4, zmagic, z4magic, mult);
            break;
          case 10:
            process_check_q4<10>(soft, c2v_row, tab_row, rot, j, Z, Z4, zmagic, z4magic, mult);
            break;
          default:
            process_check_q4<19>(soft, c2v_row, tab_row, rot, j, Z, Z4, zmagic, z4magic, mult);
            break;
        }
      }
      __syncthreads();
    }
    const bool last_it = (it + 1 == max_it);
    if (mode != MODE_EARLY_STOP && !last_it) {
      continue;
    }
    // Hard decision of the first K soft bits: bit = (llr <= 0), MSB first. A warp takes 32 consecutive bases of one
    // column: four ballots give four runs of 32 bits (one per quarter of the column), OR-ed into the packed words.
    for (uint32_t i = t; i < HBW + 1; i += TPB) {
      hb[i] = 0;
    }
    __syncthreads();
    {
      // The CTA has ceil(Z4 / 32) warps, one per run of 32 bases: warp w owns bases [32 w, 32 w + 32) of every column.
      const uint32_t b0 = 32 * warp;
      const uint32_t b  = b0 + 31 - lane; // ballot bit l <-> base b0 + 31 - l: the run is MSB-first
      const bool     in = b < Z4;
      uint32_t       nz0 = 0x00010001U, nz1 = 0x00010001U;
      for (uint32_t c = 0; c != Kb; ++c) {
        uint2 s = make_uint2(pk::SOFT_ZERO2 + 0x00010001U, pk::SOFT_ZERO2 + 0x00010001U); // positive, non-zero
        if (in) {
          s = soft[c * Z4 + b];
        }
        uint32_t p0 = pk::positive_lanes(s.x); // 1 where soft > 0
        uint32_t p1 = pk::positive_lanes(s.y);
        nz0 &= pk::minu2(s.x ^ pk::SOFT_ZERO2, 0x00010001U);
        nz1 &= pk::minu2(s.y ^ pk::SOFT_ZERO2, 0x00010001U);
        uint32_t r0 = __ballot_sync(0xffffffffU, (p0 & 0xffffU) == 0); // quarter 0: k = b
        uint32_t r1 = __ballot_sync(0xffffffffU, (p1 & 0xffffU) == 0); // quarter 1: k = b + Z/4
        uint32_t r2 = __ballot_sync(0xffffffffU, (p0 >> 16) == 0);     // quarter 2
        uint32_t r3 = __ballot_sync(0xffffffffU, (p1 >> 16) == 0);     // quarter 3
        uint32_t rlo = (lane & 1) ? r1 : r0, rhi = (lane & 1) ? r3 : r2;
        if (lane < 4) {
          const uint32_t mine = (lane & 2) ? rhi : rlo;
          const uint32_t o    = c * Z + lane * Z4 + b0; // bit offset of the first variable of this run
          const uint32_t sh   = o & 31;
          if (sh == 0 && (Z4 & 31U) == 0) {
            hb[o >> 5] = mine; // aligned runs (Z = 128, 256, 384): every word has exactly one writer
          } else if (mine != 0) {
            atomicOr(&hb[o >> 5], mine >> sh);
            if (sh != 0) {
              atomicOr(&hb[(o >> 5) + 1], mine << (32 - sh));
            }
          }
        }
      }
      nz0 = __reduce_and_sync(0xffffffffU, nz0);
      nz1 = __reduce_and_sync(0xffffffffU, nz1);
      if (lane == 0 && (nz0 & nz1) != 0x00010001U) {
        atomicOr(&misc[1], 1U);
      }
    }
    __syncthreads();
    if (warp == 0) {
      uint32_t ok = 0;
      if (poly != 0) {
        uint32_t crc = warp_crc_words<false>(hb, nb, poly, tabs, lane);
        ok           = (crc == 0 && (mode != MODE_EARLY_STOP || misc[1] == 0)) ? 1U : 0U;
      }
      if (lane == 0) {
        misc[2] = ok;
        misc[1] = 0;
      }
    }
    __syncthreads();
    const bool ok = misc[2] != 0;
    if (ok || last_it) {
      // The output holds the hard decision of the last iteration run (ldpc_decoder_impl.cpp:126-146).
      for (uint32_t i = t; i < HBW; i += TPB) {
        uint32_t wd  = __byte_perm(hb[i], 0, 0x0123); // bits beyond K are zero; the host merges a partial last byte
        slot_bits[i] = wd;
        if (d.bits_out != nullptr) {
          reinterpret_cast<uint32_t*>(d.bits_out)[i] = wd;
        }
      }
    }
    if (ok) {
      crc_ok = 1;
      iters  = (mode == MODE_EARLY_STOP) ? (int)it + 1 : (int)max_it;
      break;
    }
    __syncthreads();
  }

  if (t == 0) {
    results[cb] = {iters, crc_ok, L, 0U};
    if (d.flags & FLAG_TRACK_CRC) {
      crc_flags[d.slot] = crc_ok;
    }
  }
}

} // namespace pusch_dec
