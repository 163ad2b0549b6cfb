// C ABI (include/srsran_cuda_pusch_dec.h) of the B200 PUSCH channel-decoding accelerator: batch contexts, HARQ state in
// HBM, launch sequencing. No CPU fallback: every entry point needs a CUDA device.
#include <nvtx3/nvToolsExt.h>
#include "../../include/srsran_cuda_pusch_dec.h"
#include "pusch_dec_kernels.cuh"
#include "ldpc_packed.cuh"
#include "pusch_demod.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <tuple>
#include <vector>

using namespace pusch_dec;

namespace {
// Threads per rate-dematcher CTA (one code block): 128 so that a CTA fits beside a packed-decoder CTA on the same SM.
// From this many separate page-locked pieces on, a batch's soft bits are read by the gather kernel.
constexpr size_t H2D_GATHER_MIN_PIECES = 4;
// Largest batch (bytes of host soft bits) whose soft bits the dematcher reads from host memory itself.
constexpr size_t DIRECT_IN_MAX_BYTES = size_t(4) << 20;
// Largest batch (code blocks) whose dematcher reads the descriptors from host memory instead of waiting for their copy.
constexpr uint32_t EARLY_DM_MAX_CBS = 512;
#ifndef DIRECT_IN_DEFAULT
#define DIRECT_IN_DEFAULT 1
#endif
#ifndef DIRECT_OUT_DEFAULT
#define DIRECT_OUT_DEFAULT 1
#endif
#ifndef H2D_GATHER_DEFAULT
#define H2D_GATHER_DEFAULT 32 // CTAs of the gather kernel
#endif
#ifndef DM_THREADS
#define DM_THREADS 128
#endif
#ifdef PUSCH_DEC_HOST_PROF
static double g_prof[8];
static long   g_prof_n;
static inline double prof_now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define PROF_T(var) double var = prof_now()
#define PROF_ADD(i, a, b) g_prof[i] += (b) - (a)
#else
#define PROF_T(var)
#define PROF_ADD(i, a, b)
#endif

thread_local std::string g_create_error;

constexpr int      NOF_CONTEXTS   = 8;
constexpr uint32_t MAX_TBS_PER_CTX = 1024;
constexpr size_t   PACKED_OUT_MAX  = 512 * 1024; // results + TB bytes of a small batch that return in one copy
constexpr size_t   PACKED_META_MAX = 16 * 1024;  // group / order / TB descriptors of a small batch that travel in one copy

template <typename T>
struct pinned_buf {
  T*     p   = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n)
  {
    if (n <= cap) {
      return cudaSuccess;
    }
    if (p != nullptr) {
      cudaFreeHost(p);
      p = nullptr;
    }
    cap            = 0;
    cudaError_t e  = cudaMallocHost(reinterpret_cast<void**>(&p), n * sizeof(T));
    if (e == cudaSuccess) {
      cap = n;
    }
    return e;
  }
  void release()
  {
    if (p != nullptr) {
      cudaFreeHost(p);
    }
    p   = nullptr;
    cap = 0;
  }
};

template <typename T>
struct device_buf {
  T*     p   = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n)
  {
    if (n <= cap) {
      return cudaSuccess;
    }
    if (p != nullptr) {
      cudaFree(p);
      p = nullptr;
    }
    cap           = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), n * sizeof(T));
    if (e == cudaSuccess) {
      cap = n;
    }
    return e;
  }
  void release()
  {
    if (p != nullptr) {
      cudaFree(p);
    }
    p   = nullptr;
    cap = 0;
  }
};

struct cb_host_meta {
  uint32_t K;       // message length in bits
  uint32_t max_it;
  uint32_t slot;
};

struct tb_host_meta {
  uint32_t first_cb, nof_cbs, tbs_bits, out_offset, max_it;
  bool     polled;
};

/// One batch in flight: staging buffers, descriptors and results.
struct batch_context {
  cudaStream_t stream     = nullptr;
  cudaEvent_t  done       = nullptr; // everything (incl. D2H) complete
  cudaEvent_t  kernels    = nullptr; // kernels complete (HARQ ordering between contexts)
  cudaEvent_t  copied     = nullptr; // host -> device copies complete (copies of consecutive batches go one after the other)
  cudaStream_t tail       = nullptr; // high-priority stream of the small end-of-batch kernels and the result copies
  cudaEvent_t  decoded    = nullptr; // decode kernels complete (fork point of `tail`)
  cudaEvent_t  stage[4]   = {nullptr, nullptr, nullptr, nullptr}; // begin, copies in, dematch done, decode done
  // Decode launch classes (different lifting-size / shared-memory shapes) run concurrently on side streams.
  static constexpr int NOF_SIDE = 7;
  cudaStream_t side[NOF_SIDE]  = {};
  cudaEvent_t  fork            = nullptr;
  cudaEvent_t  join[NOF_SIDE]  = {};
  uint32_t     slot_lo    = 0xffffffffU; // range of HARQ slots this batch touches (ordering between batches)
  uint32_t     slot_hi    = 0;
  // The HARQ slots themselves as runs [first, last]: the rx buffers of a slot's transport blocks own scattered runs of
  // code-block ids, so two batches whose RANGES overlap usually share no slot at all and need no ordering.
  std::vector<std::pair<uint32_t, uint32_t>> slot_runs;
  bool                                       runs_sorted = false;
  cudaEvent_t  wait_for   = nullptr; // set by a streamed submit: the batch's kernels wait for the ingest copies
  bool         open       = false;   // accepting operations, not launched
  bool         in_flight  = false;   // launched, results not yet consumed
  bool         hal_style  = false;   // descriptors hold staging OFFSETS until fixup_and_launch (enqueue path)
  std::vector<cudaEvent_t> wait_events; // ingest copies the batch's kernels wait for (streamed transport blocks)
  // Soft-buffer extents this (not yet launched) batch has advanced: (slot, previous value). Committed when the batch is
  // launched, restored if it is abandoned - the host-side extents must describe what the device actually wrote.
  std::vector<std::pair<uint32_t, uint32_t>> extent_undo;
  uint32_t     generation = 0;

  pinned_buf<int8_t>        h_llr;
  device_buf<int8_t>        d_llr;
  size_t                    llr_used = 0;
  pinned_buf<cb_desc>       h_desc;
  device_buf<cb_desc>       d_desc;
  pinned_buf<uint32_t>      h_order;
  device_buf<uint32_t>      d_order;
  pinned_buf<grp_desc>      h_grp; // packed-decoder groups (four code blocks per CTA)
  device_buf<grp_desc>      d_grp;
  pinned_buf<uint32_t>      h_tbmap; // per code block: index of its transport block in h_tb, or 0xffffffff
  device_buf<uint32_t>      d_tbmap;
  device_buf<uint32_t>      d_tbshare; // per code block: its share of the TB CRC24A
  device_buf<uint32_t>      d_tbdone;  // per transport block: code blocks assembled so far (self-resetting counter)
  pinned_buf<cb_result>     h_res;
  device_buf<cb_result>     d_res;
  pinned_buf<uint8_t>       h_bits;
  device_buf<uint8_t>       d_bits;
  pinned_buf<tb_desc>       h_tb;
  device_buf<tb_desc>       d_tb;
  // Small batches (a single transport block: the latency case) send their group / order / TB descriptors packed in ONE copy.
  pinned_buf<uint8_t>       h_meta;
  device_buf<uint8_t>       d_meta;
  // ... and bring their results (code-block results, TB verdicts, TB bytes) back packed in ONE copy. The r_* pointers are where
  // the results of the batch in flight are read: the context's own result buffers or the packed region.
  pinned_buf<uint8_t>       h_outp;
  device_buf<uint8_t>       d_outp;
  const cb_result*          r_res       = nullptr;
  const tb_result_dev*      r_tbres     = nullptr;
  const uint8_t*            r_tbout     = nullptr;
  const uint8_t*            r_tbout_dev = nullptr;
  pinned_buf<tb_result_dev> h_tbres;
  device_buf<tb_result_dev> d_tbres;
  pinned_buf<uint8_t>       h_tbout;
  bool                      tb_on_host = true; // this batch's transport blocks were copied to h_tbout
  device_buf<uint8_t>       d_tbout;
  size_t                    tbout_used = 0;
  bool                      want_bits  = false; // per-CB decoded bits are copied back (HAL path, unit-level decode)
  device_buf<uint8_t>       d_unit_bits;        // scratch data-bit slots of unit-level decodes (not HARQ state)
  uint8_t*                  bits_base  = nullptr; // data-bit slots the decoder writes: HARQ slots or d_unit_bits

  std::vector<cb_host_meta> cb_meta;
  std::vector<tb_host_meta> tb_meta;
  struct copy_job {
    const int8_t* src;     // page-locked caller memory, nullptr: the context's own staging buffer
    size_t        dst_off;
    size_t        bytes;
    const int8_t* dev_src; // the same memory as the device sees it (mapped page-locked memory), nullptr if unknown
  };
  std::vector<copy_job>  copies;
  pinned_buf<h2d_piece> h_pieces; // descriptors of the gather kernel that replaces many separate copies

  // Device-side demodulation (pusch_demod.cuh): equalized symbols + noise variances staged in d_dm_in (host callers),
  // scrambling sequences in d_scr, one descriptor per transport block.
  struct raw_copy {
    const void* src;
    size_t      dst_off; // in d_dm_in
    size_t      bytes;
  };
  device_buf<uint8_t>      d_dm_in;
  size_t                   dm_in_used = 0;
  device_buf<uint32_t>     d_scr;
  pinned_buf<demod_desc>   h_dm;
  device_buf<demod_desc>   d_dm;
  uint32_t                 ndm          = 0;
  uint32_t                 dm_max_sym   = 0;
  uint32_t                 dm_max_words = 0;
  uint32_t                 dm_qm_mask   = 0; // modulation orders present (bit QM)
  std::vector<raw_copy>    raw_copies;
  cudaEvent_t              dm_ev[2] = {nullptr, nullptr};
};

/// Streaming ingestion of the LLRs of one transport block (pusch_decoder_buffer::on_new_softbits): every pushed block is
/// copied to the device at once on the ingest stream; the batch that decodes the transport block waits for `pushed`.
struct ingest_slot {
  static constexpr int NOF = 32;
  cudaStream_t       stream = nullptr;
  cudaEvent_t        pushed = nullptr;
  device_buf<int8_t> d_llr;
  uint32_t           used       = 0;
  uint32_t           capacity   = 0;     // LLRs announced at stream_begin
  bool               open       = false; // between stream_begin and stream_submit
  int                ctx        = -1;    // batch context that reads d_llr (until it completes)
  uint32_t           generation = 0;
};

struct hal_op {
  srsran_cuda_pusch_dec_cb_config cfg;
  bool                            configured = false;
  int                             ctx        = -1;
  uint32_t                        idx        = 0;
  uint32_t                        generation = 0;
  bool                            enqueued   = false;
};

} // namespace

struct srsran_cuda_pusch_dec {
  int      device            = 0;
  uint32_t max_cbs           = 0;
  uint32_t nof_slots         = 0;
  uint32_t combine_block     = 64; // AVX-512 flavour of the reference's combine (32 = AVX2, 0 = generic)
  uint64_t launches          = 0;
  bool     use_packed        = true; // route eligible code blocks to the packed (4 per CTA) decoder
  bool     use_tmem          = true;  // packed decoder with the messages in tensor memory (two CTAs per SM) where eligible
  bool     use_long          = true;  // many-layer code blocks two per CTA with the messages in tensor memory
  bool     use_bulk          = false; // one-CTA-per-SM forms stage their inputs with cp.async.bulk + mbarrier (variant 7):
                                      // measured equal to plain 128-bit loads, kept selectable
  bool     tb_host_copy      = true;  // decoded transport blocks are copied to the page-locked result buffer
  uint32_t h2d_gather_ctas   = H2D_GATHER_DEFAULT;    // CTAs of the gather kernel that reads many separate page-locked pieces (0: copies only)
  bool     direct_out        = DIRECT_OUT_DEFAULT != 0; // small batches: kernels write results and TB bytes into page-locked host memory
  bool     direct_in         = DIRECT_IN_DEFAULT != 0;  // small batches: the dematcher reads the soft bits from page-locked host memory
  bool     prefer_long       = false; // A/B: the many-layer pair form also where the shared-memory pair form fits
  int      last_unit_ctx     = -1;    // context of the last unit-level batch (srsran_cuda_pusch_dec_last_unit_timing)
  bool     force_pairs       = false; // groups of two code blocks per CTA (two CTAs per SM) also for large batches
  bool     prefer_q4         = false; // one code block per CTA on the packed arithmetic also where groups of four would fit
  cudaEvent_t timer_begin    = nullptr;
  cudaEvent_t timer_end      = nullptr;
  bool        timer_armed    = false;
  int      max_smem_optin    = 0;
  int      nof_sms           = 148;
  std::string last_error;

  device_buf<int8_t>   d_soft;
  device_buf<uint8_t>  d_bits;
  device_buf<uint32_t> d_crc_flags;
  std::vector<uint32_t> extent; // per slot: upper bound of the non-zero extent of the soft buffer

  batch_context ctx[NOF_CONTEXTS];
  int           open_ctx      = -1;
  int           last_launched = -1;

  hal_op hal[SRSRAN_CUDA_MAX_NOF_SEGMENTS];

  ingest_slot ingest[ingest_slot::NOF];

  // Scrambling-sequence tables of the device-side demodulator (built on first use).
  bool                 demod_ready = false;
  device_buf<uint32_t> d_scr_x1; // packed x1(n + 1600)
  device_buf<uint32_t> d_scr_x2; // [chunk][bit of c_init][32]: x2 windows at every chunk start

  // unit-level scratch
  device_buf<crc_job>  d_crc_jobs;
  device_buf<uint32_t> d_crc_out;
  device_buf<uint8_t>  d_crc_msg;
};

namespace {

#define CUDA_TRY(h, expr)                                                                                              \
  do {                                                                                                                 \
    cudaError_t e__ = (expr);                                                                                          \
    if (e__ != cudaSuccess) {                                                                                          \
      (h)->last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);                                           \
      return SRSRAN_CUDA_ERR_CUDA;                                                                                     \
    }                                                                                                                  \
  } while (0)

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

const double K0_FACTOR[2][4] = {{0, 17, 33, 56}, {0, 13, 25, 43}};

uint32_t compute_k0(uint32_t bg, uint32_t rv, uint32_t Ncb, uint32_t N, uint32_t Z)
{
  // ldpc_rate_dematcher_impl.cpp:104-105: floor((shift_factor * buffer_length) / block_length) * lifting_size.
  double tmp = (K0_FACTOR[bg - 1][rv] * Ncb) / N;
  return static_cast<uint32_t>(static_cast<uint16_t>(std::floor(tmp))) * Z;
}

/// Upper bound of the non-zero extent of a soft buffer after one dematching operation (see DESIGN.md "extent").
uint32_t extent_after_dematch(uint32_t prev, uint32_t N, uint32_t Ncb, uint32_t k0, uint32_t E, uint32_t info,
                              uint32_t sys, bool new_data)
{
  if (prev > N) {
    // The slot still holds LLRs of an earlier, longer code block beyond this one's N bytes: nothing here clears them.
    return prev;
  }
  uint32_t D         = info + (Ncb - sys);
  uint32_t s0        = (k0 < info) ? k0 : ((k0 < sys) ? info : info + (k0 - sys));
  uint32_t first_len = D - s0;
  if (E > first_len) {
    return std::max(prev, Ncb);
  }
  uint32_t end_slot = s0 + E;
  uint32_t idx_end  = (end_slot <= info) ? sys : sys + (end_slot - info);
  if (!new_data) {
    return std::max(prev, idx_end);
  }
  if (idx_end >= Ncb) {
    return std::max(prev, Ncb);
  }
  uint32_t tail_start = N - (Ncb - idx_end);
  return std::max(idx_end, std::min(prev, tail_start));
}

uint32_t layers_for_extent(uint32_t ext, uint32_t bg, uint32_t Z)
{
  uint32_t Kb  = (bg == 1) ? 22 : 10;
  uint32_t cbl = std::max(ext + 2 * Z, (Kb + 4) * Z);
  cbl          = ((cbl + Z - 1) / Z) * Z;
  uint32_t L   = cbl / Z - Kb;
  return std::min(L, (bg == 1) ? 46U : 42U);
}

uint16_t scale_mult(float sf)
{
  // mm512::scale_epi8 (avx512_support.h:69-83): identity for sf >= .9999, else (uint16)(sf * 65536).
  if (sf >= .9999) {
    return 0;
  }
  constexpr unsigned FLOAT2INT = 1U << 16U;
  return static_cast<uint16_t>(sf * FLOAT2INT);
}

/// NVTX range over a host-side phase (submit / launch / poll), visible in Nsight Systems timelines next to the kernels -
/// the counterpart of the reference's l1_tracer events around the PUSCH decoder. Without a tool
/// attached the calls are a null-pointer test.
struct nvtx_scope {
  explicit nvtx_scope(const char* name) { nvtxRangePushA(name); }
  ~nvtx_scope() { nvtxRangePop(); }
  nvtx_scope(const nvtx_scope&)            = delete;
  nvtx_scope& operator=(const nvtx_scope&) = delete;
};

int tpc_class(uint32_t Z)
{
  return Z <= 32 ? 0 : (Z <= 64 ? 1 : (Z <= 128 ? 2 : (Z <= 256 ? 3 : 4)));
}

template <int TPC, int CBS>
cudaError_t launch_decode(srsran_cuda_pusch_dec* h, cudaStream_t s, const cb_desc* descs, const uint32_t* order,
                          cb_result* res, uint8_t* bits_base, uint32_t n, uint32_t smem_per_cb)
{
  uint32_t grid = (n + CBS - 1) / CBS;
  ldpc_decode_kernel<TPC, CBS><<<grid, TPC * CBS, smem_per_cb * CBS, s>>>(
      descs, order, res, h->d_soft.p, bits_base, h->d_crc_flags.p, n, smem_per_cb);
  ++h->launches;
  return cudaGetLastError();
}

template <int TPC>
cudaError_t launch_decode4(srsran_cuda_pusch_dec* h, cudaStream_t s, const cb_desc* descs, const grp_desc* groups,
                           cb_result* res, uint8_t* bits_base, uint32_t n, uint32_t smem, bool z384 = false)
{
  if (TPC == 384 && z384) {
    ldpc_decode4_kernel<384, 384><<<n, 384, smem, s>>>(descs, groups, res, h->d_soft.p, bits_base, h->d_crc_flags.p, 0U);
  } else {
    ldpc_decode4_kernel<TPC><<<n, TPC, smem, s>>>(descs, groups, res, h->d_soft.p, bits_base, h->d_crc_flags.p, 0U);
  }
  ++h->launches;
  return cudaGetLastError();
}

/// Four code blocks per CTA with the messages in tensor memory (`tm_cols` columns per CTA: 256 -> two CTAs per SM).
template <int TPC>
cudaError_t launch_decode4t(srsran_cuda_pusch_dec* h, cudaStream_t s, const cb_desc* descs, const grp_desc* groups,
                            cb_result* res, uint8_t* bits_base, uint32_t n, uint32_t smem, uint32_t tm_cols, bool z384 = false)
{
  if (TPC == 384 && z384) {
    ldpc_decode4_kernel<384, 384, 2, 1><<<n, 384, smem, s>>>(descs, groups, res, h->d_soft.p, bits_base, h->d_crc_flags.p, tm_cols);
  } else {
    ldpc_decode4_kernel<TPC, 0, 2, 1><<<n, TPC, smem, s>>>(descs, groups, res, h->d_soft.p, bits_base, h->d_crc_flags.p, tm_cols);
  }
  ++h->launches;
  return cudaGetLastError();
}

/// Two code blocks per CTA (two CTAs per SM): the same kernel with one register of two lanes per thread.
template <int TPC>
cudaError_t launch_decode2(srsran_cuda_pusch_dec* h, cudaStream_t s, const cb_desc* descs, const grp_desc* groups,
                           cb_result* res, uint8_t* bits_base, uint32_t n, uint32_t smem, bool z384 = false)
{
  if (TPC == 384 && z384) {
    ldpc_decode4_kernel<384, 384, 1><<<n, 384, smem, s>>>(descs, groups, res, h->d_soft.p, bits_base, h->d_crc_flags.p, 0U);
  } else {
    ldpc_decode4_kernel<TPC, 0, 1><<<n, TPC, smem, s>>>(descs, groups, res, h->d_soft.p, bits_base, h->d_crc_flags.p, 0U);
  }
  ++h->launches;
  return cudaGetLastError();
}

/// Two code blocks with many layers per CTA, messages in tensor memory (all 512 columns: one CTA per SM).
template <int TPC>
cudaError_t launch_decode2t(srsran_cuda_pusch_dec* h, cudaStream_t s, const cb_desc* descs, const grp_desc* groups,
                            cb_result* res, uint8_t* bits_base, uint32_t n, uint32_t smem, uint32_t tm_cols, bool z384 = false)
{
  if (TPC == 384 && z384) {
    ldpc_decode4_kernel<384, 384, 1, 1><<<n, 384, smem, s>>>(descs, groups, res, h->d_soft.p, bits_base, h->d_crc_flags.p, tm_cols);
  } else {
    ldpc_decode4_kernel<TPC, 0, 1, 1><<<n, TPC, smem, s>>>(descs, groups, res, h->d_soft.p, bits_base, h->d_crc_flags.p, tm_cols);
  }
  ++h->launches;
  return cudaGetLastError();
}

template <int TPB>
cudaError_t launch_decode_q4(srsran_cuda_pusch_dec* h, cudaStream_t s, const cb_desc* descs, const uint32_t* order,
                             cb_result* res, uint8_t* bits_base, uint32_t n, uint32_t smem)
{
  ldpc_decode_q4_kernel<TPB><<<n, TPB, smem, s>>>(descs, order, res, h->d_soft.p, bits_base, h->d_crc_flags.p);
  ++h->launches;
  return cudaGetLastError();
}

/// Code blocks the one-code-block packed kernel (ldpc_decode_q4_kernel) accepts.
bool q4_eligible(const srsran_cuda_pusch_dec* h, const cb_desc& d)
{
  if (!h->use_packed || !(d.flags & FLAG_DECODE) || d.Z < 16 || (d.Z % 4) != 0) {
    return false;
  }
  return decq_smem_layout(d.bg, d.Z, d.layer_cap).total <= static_cast<uint32_t>(h->max_smem_optin);
}

/// Code blocks the packed decoder (ldpc_packed.cuh) accepts; `cap` = layers it would process.
bool packed_eligible(const srsran_cuda_pusch_dec* h, const cb_desc& d, uint32_t cap, uint32_t lanes)
{
  if (!h->use_packed || !(d.flags & FLAG_DECODE) || !(d.flags & FLAG_USE_HARQ) || d.mode == MODE_NO_CRC ||
      d.crc_poly == 0 || d.Z < 144 || (d.Z % 16) != 0) {
    return false;
  }
  return dec4_smem_layout(d.bg, d.Z, cap, lanes).total <= static_cast<uint32_t>(h->max_smem_optin);
}

/// Tensor-memory columns a group of four such code blocks needs per CTA (256 or 512), 0 if the tensor-memory variant of the
/// packed decoder does not apply: whole warps only (Z % 32 == 0), and the messages of `cap` layers must fit in the columns
/// of a warp (the TPC / 128 warps of a lane quadrant share the CTA's columns).
uint32_t packed_tmem_cols(const srsran_cuda_pusch_dec* h, const cb_desc& d, uint32_t cap)
{
  if (!h->use_tmem || !packed_eligible(h, d, 4, 2) || (d.Z % 32) != 0) {
    return 0;
  }
  if (dec4_smem_layout(d.bg, d.Z, cap, 4, true).total > static_cast<uint32_t>(h->max_smem_optin)) {
    return 0;
  }
  const uint32_t tpc  = d.Z <= 256 ? 256 : 384;
  const uint32_t need = dec4_tmem_cols(d.bg, cap);
  if (need <= dec4_tmem_cols_per_warp(256, tpc)) {
    return 256;
  }
  if (need <= dec4_tmem_cols_per_warp(512, tpc)) {
    return 512;
  }
  return 0;
}

/// Shared memory of a pair of such code blocks on the many-layer form of the packed decoder (two code blocks per CTA,
/// messages in all 512 columns of tensor memory, the layers that do not fit there in shared memory); 0 if it does not
/// apply: whole warps only (Z % 32 == 0), 16-byte aligned decoder input, and only where the one-code-block kernel would run
/// fewer than three CTAs per SM (shorter code blocks are better off there). Any mode (with or without CRC) and any input
/// (HARQ slot or the unit-level interface's staged soft bits).
uint32_t packed_long_smem(const srsran_cuda_pusch_dec* h, const cb_desc& d, uint32_t cap)
{
  if (!h->use_packed || !h->use_tmem || !h->use_long || !(d.flags & FLAG_DECODE) || d.Z < 144 || (d.Z % 32) != 0) {
    return 0;
  }
  if (!(d.flags & FLAG_USE_HARQ) && (reinterpret_cast<uintptr_t>(d.llr) & 15U) != 0) {
    return 0;
  }
  if (3 * (decq_smem_layout(d.bg, d.Z, cap).total + 1024) <= static_cast<uint32_t>(h->max_smem_optin)) {
    return 0;
  }
  const uint32_t tpc   = d.Z <= 256 ? 256 : 384;
  const uint32_t lt    = dec2_tm_layers(d.bg, cap, dec4_tmem_cols_per_warp(512, tpc));
  const uint32_t total = dec4_smem_layout(d.bg, d.Z, cap, 2, true, lt).total;
  return total <= static_cast<uint32_t>(h->max_smem_optin) ? total : 0U;
}

template <int TPC, int CBS>
cudaError_t set_smem_attr(int bytes)
{
  return cudaFuncSetAttribute(ldpc_decode_kernel<TPC, CBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

template <typename Handle>
int upload_tables(Handle* h)
{
  uint16_t row_ptr[2][48] = {};
  uint8_t  col[2][MAX_EDGES] = {};
  static uint16_t shift[2][8][MAX_EDGES];
  std::memset(shift, 0, sizeof(shift));
  std::memcpy(row_ptr[0], NR_BG1_ROW_PTR, sizeof(NR_BG1_ROW_PTR));
  std::memcpy(row_ptr[1], NR_BG2_ROW_PTR, sizeof(NR_BG2_ROW_PTR));
  std::memcpy(col[0], NR_BG1_COL, sizeof(NR_BG1_COL));
  std::memcpy(col[1], NR_BG2_COL, sizeof(NR_BG2_COL));
  for (int i = 0; i != 8; ++i) {
    std::memcpy(shift[0][i], NR_BG1_SHIFT[i], sizeof(NR_BG1_SHIFT[i]));
    std::memcpy(shift[1][i], NR_BG2_SHIFT[i], sizeof(NR_BG2_SHIFT[i]));
  }
  static uint32_t xp32[3][272], xp128[3][1024];
  for (int poly = 1; poly <= 3; ++poly) {
    uint32_t gen = crc_gen(poly), order = crc_order(poly);
    uint32_t x32  = crc_push_bits(1, 0, 32, gen, order);  // x^32 mod g
    uint32_t x128 = crc_push_bits(1, 0, 32, gen, order);
    x128          = gf2_mulmod(x128, x128, gen, order);    // x^64
    x128          = gf2_mulmod(x128, x128, gen, order);    // x^128
    uint32_t v    = 1;
    for (int i = 0; i != 272; ++i) {
      xp32[poly - 1][i] = v;
      v                 = gf2_mulmod(v, x32, gen, order);
    }
    v = 1;
    for (int i = 0; i != 1024; ++i) {
      xp128[poly - 1][i] = v;
      v                  = gf2_mulmod(v, x128, gen, order);
    }
  }
  CUDA_TRY(h, cudaMemcpyToSymbol(c_row_ptr, row_ptr, sizeof(row_ptr)));
  CUDA_TRY(h, cudaMemcpyToSymbol(c_col, col, sizeof(col)));
  CUDA_TRY(h, cudaMemcpyToSymbol(c_shift, shift, sizeof(shift)));
  CUDA_TRY(h, cudaMemcpyToSymbol(c_xpow32, xp32, sizeof(xp32)));
  CUDA_TRY(h, cudaMemcpyToSymbol(g_xpow32, xp32, sizeof(xp32)));
  CUDA_TRY(h, cudaMemcpyToSymbol(c_xpow128, xp128, sizeof(xp128)));
  uint32_t xp2[3][32];
  for (int poly = 1; poly <= 3; ++poly) {
    uint32_t gen = crc_gen(poly), order = crc_order(poly);
    uint32_t v   = 2; // x
    for (int i = 0; i != 32; ++i) {
      xp2[poly - 1][i] = v;
      v                = gf2_mulmod(v, v, gen, order);
    }
  }
  CUDA_TRY(h, cudaMemcpyToSymbol(c_xpow2, xp2, sizeof(xp2)));
  static uint32_t tabs[3][1024];
  for (int poly = 1; poly <= 3; ++poly) {
    uint32_t gen = crc_gen(poly), order = crc_order(poly);
    for (uint32_t i = 0; i != 1024; ++i) {
      uint32_t r        = crc_push_bits(0, (i & 0xffU) << 24, 8, gen, order);
      tabs[poly - 1][i] = crc_push_bits(r, 0, 8 * (i >> 8), gen, order);
    }
  }
  CUDA_TRY(h, cudaMemcpyToSymbol(g_crc_tabs, tabs, sizeof(tabs)));
  return SRSRAN_CUDA_OK;
}

bool is_pinned(const void* p, const void** dev_ptr = nullptr)
{
  cudaPointerAttributes attr;
  cudaError_t           e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (dev_ptr != nullptr) {
    *dev_ptr = (attr.type == cudaMemoryTypeHost) ? attr.devicePointer : nullptr;
  }
  return attr.type == cudaMemoryTypeHost;
}

/// Waits for a context's results and marks it reusable.
int finish_context(srsran_cuda_pusch_dec* h, batch_context& c)
{
  if (c.in_flight) {
    CUDA_TRY(h, cudaEventSynchronize(c.done));
  }
  return SRSRAN_CUDA_OK;
}

bool context_has_unpolled(const batch_context& c)
{
  for (const tb_host_meta& t : c.tb_meta) {
    if (!t.polled) {
      return true;
    }
  }
  return false;
}

/// Opens a fresh context (waiting for the oldest one if necessary). Returns its index or < 0.
int open_context(srsran_cuda_pusch_dec* h, uint32_t min_cbs)
{
  if (h->open_ctx >= 0) {
    return h->open_ctx;
  }
  int pick = -1;
  for (int k = 1; k <= NOF_CONTEXTS; ++k) {
    int i = (h->last_launched + k) % NOF_CONTEXTS;
    if (i < 0) {
      i = 0;
    }
    batch_context& c = h->ctx[i];
    if (c.in_flight && context_has_unpolled(c)) {
      continue;
    }
    pick = i;
    break;
  }
  if (pick < 0) {
    // Not an error of the caller: the accelerator is busy. Poll (consume) older tickets and submit again.
    h->last_error = "all batch contexts hold unpolled transport blocks";
    return SRSRAN_CUDA_ERR_BUSY;
  }
  batch_context& c = h->ctx[pick];
  if (c.in_flight) {
    int r = finish_context(h, c);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
    c.in_flight = false;
  }
  uint32_t ncb = std::max(min_cbs, h->max_cbs);
  CUDA_TRY(h, c.h_desc.reserve(ncb));
  CUDA_TRY(h, c.d_desc.reserve(ncb));
  CUDA_TRY(h, c.h_order.reserve(ncb));
  CUDA_TRY(h, c.d_order.reserve(ncb));
  CUDA_TRY(h, c.h_grp.reserve(ncb));
  CUDA_TRY(h, c.d_grp.reserve(ncb));
  CUDA_TRY(h, c.h_tbmap.reserve(ncb));
  CUDA_TRY(h, c.d_tbmap.reserve(ncb));
  CUDA_TRY(h, c.d_tbshare.reserve(ncb));
  CUDA_TRY(h, c.h_res.reserve(ncb));
  CUDA_TRY(h, c.d_res.reserve(ncb));
  // Packed descriptor / result regions of small batches (allocated here, not at the first small batch: no hiccup there).
  CUDA_TRY(h, c.h_meta.reserve(PACKED_META_MAX));
  CUDA_TRY(h, c.d_meta.reserve(PACKED_META_MAX));
  CUDA_TRY(h, c.h_outp.reserve(PACKED_OUT_MAX));
  CUDA_TRY(h, c.d_outp.reserve(PACKED_OUT_MAX));
  c.open       = true;
  c.slot_lo    = 0xffffffffU;
  c.slot_hi    = 0;
  c.slot_runs.clear();
  c.runs_sorted = false;
  c.llr_used   = 0;
  c.tbout_used = 0;
  c.want_bits  = false;
  c.bits_base  = h->d_bits.p;
  c.cb_meta.clear();
  c.tb_meta.clear();
  c.copies.clear();
  c.raw_copies.clear();
  c.wait_events.clear();
  c.extent_undo.clear();
  c.hal_style    = false;
  c.ndm          = 0;
  c.dm_in_used   = 0;
  c.dm_max_sym   = 0;
  c.dm_max_words = 0;
  c.dm_qm_mask   = 0;
  ++c.generation;
  h->open_ctx = pick;
  return pick;
}

/// Notes that the open batch touches HARQ slot `slot` (soft bits, data bits or CRC flag).
inline void note_slot(batch_context& c, uint32_t slot)
{
  c.slot_lo = std::min(c.slot_lo, slot);
  c.slot_hi = std::max(c.slot_hi, slot);
  if (!c.slot_runs.empty() && c.slot_runs.back().second + 1 == slot) {
    c.slot_runs.back().second = slot; // the code blocks of a transport block mostly take consecutive slots
  } else if (c.slot_runs.empty() || slot < c.slot_runs.back().first || slot > c.slot_runs.back().second) {
    c.slot_runs.emplace_back(slot, slot);
  }
}

/// True if the two batches share a HARQ slot.
bool slots_overlap(batch_context& a, batch_context& b)
{
  if (a.slot_lo > a.slot_hi || b.slot_lo > b.slot_hi || a.slot_lo > b.slot_hi || b.slot_lo > a.slot_hi) {
    return false;
  }
  for (batch_context* c : {&a, &b}) {
    if (!c->runs_sorted) {
      std::sort(c->slot_runs.begin(), c->slot_runs.end());
      c->runs_sorted = true;
    }
  }
  size_t i = 0, j = 0;
  while (i != a.slot_runs.size() && j != b.slot_runs.size()) {
    if (a.slot_runs[i].second < b.slot_runs[j].first) {
      ++i;
    } else if (b.slot_runs[j].second < a.slot_runs[i].first) {
      ++j;
    } else {
      return true;
    }
  }
  return false;
}

/// Drops a batch that was never launched: the soft-buffer extents it advanced are restored and the context is closed, so
/// that no later call finds it half built.
void abandon_context(srsran_cuda_pusch_dec* h, batch_context& c)
{
  for (auto it = c.extent_undo.rbegin(); it != c.extent_undo.rend(); ++it) {
    h->extent[it->first] = it->second;
  }
  c.extent_undo.clear();
  c.wait_events.clear();
  c.wait_for = nullptr;
  c.open     = false;
  if (h->open_ctx >= 0 && &h->ctx[h->open_ctx] == &c) {
    h->open_ctx = -1;
  }
}

/// Reserves `bytes` (rounded to 16) of LLR staging in the open context; returns the offset.
int stage_llrs(srsran_cuda_pusch_dec* h, batch_context& c, const int8_t* src, size_t bytes, size_t* off_out)
{
  size_t off    = c.llr_used;
  size_t padded = (bytes + 15) & ~size_t(15);
  size_t need   = off + padded;
  if (need > c.d_llr.cap) {
    // Grow (keeps what has been staged so far).
    size_t             ncap = std::max(need, c.d_llr.cap * 2 + (size_t(1) << 20));
    device_buf<int8_t> nd;
    CUDA_TRY(h, nd.reserve(ncap));
    if (c.d_llr.p != nullptr) {
      // Nothing has been copied to the device yet for this context (copies happen at launch), so no device copy.
      c.d_llr.release();
    }
    c.d_llr = nd;
  }
  const void* dev_src = nullptr;
  if (is_pinned(src, &dev_src)) {
    // Caller memory is page-locked: copy straight from it at launch. Pieces that are adjacent on both sides (e.g. the
    // transport blocks of one slot laid out back to back) are merged into one copy.
    if (!c.copies.empty() && c.copies.back().src != nullptr && c.copies.back().src + c.copies.back().bytes == src &&
        c.copies.back().dst_off + c.copies.back().bytes == off) {
      c.copies.back().bytes += bytes;
    } else {
      c.copies.push_back({src, off, bytes, static_cast<const int8_t*>(dev_src)});
    }
  } else {
    // Pageable caller memory: stage through the context's pinned buffer (src == nullptr), merging adjacent pieces. The
    // pinned buffer exists only on this path and follows the device buffer's capacity.
    if (c.h_llr.cap < c.d_llr.cap) {
      pinned_buf<int8_t> nh;
      CUDA_TRY(h, nh.reserve(c.d_llr.cap));
      if (c.h_llr.p != nullptr) {
        std::memcpy(nh.p, c.h_llr.p, std::min(off, c.h_llr.cap));
        c.h_llr.release();
      }
      c.h_llr = nh;
    }
    std::memcpy(c.h_llr.p + off, src, bytes);
    if (!c.copies.empty() && c.copies.back().src == nullptr &&
        c.copies.back().dst_off + ((c.copies.back().bytes + 15) & ~size_t(15)) == off) {
      c.copies.back().bytes = off + bytes - c.copies.back().dst_off;
    } else {
      c.copies.push_back({nullptr, off, bytes, nullptr});
    }
  }
  c.llr_used = need;
  *off_out   = off;
  return SRSRAN_CUDA_OK;
}

struct cb_params {
  uint32_t bg, Z, E, rv, Qm, Nref, F, crc_poly, max_it, mode, new_data, slot, flags, n_in;
  float    scaling;
};

/// Appends one code-block operation to the open context. `llr_dev` is the device address of its LLRs.
int add_cb(srsran_cuda_pusch_dec* h, batch_context& c, const cb_params& p, const int8_t* llr_dev, bool want_bits,
           uint32_t* idx_out)
{
  int ils = ls_index(p.Z);
  if (ils < 0 || (p.bg != 1 && p.bg != 2) || p.rv > 3 || p.max_it == 0 || p.max_it > 255) {
    h->last_error = "invalid code-block parameters";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  uint32_t Kb = (p.bg == 1) ? 22 : 10, Ns = (p.bg == 1) ? 66 : 50;
  uint32_t N = Ns * p.Z, K = Kb * p.Z, sys = (Kb - 2) * p.Z;
  uint32_t idx = static_cast<uint32_t>(c.cb_meta.size());
  if (idx >= c.h_desc.cap) {
    h->last_error = "batch context full";
    return SRSRAN_CUDA_ERR_STATE;
  }
  cb_desc d     = {};
  d.llr         = llr_dev;
  d.E           = p.E;
  d.slot        = p.slot;
  d.N           = N;
  d.Z           = static_cast<uint16_t>(p.Z);
  d.bg          = static_cast<uint8_t>(p.bg);
  d.ils         = static_cast<uint8_t>(ils);
  d.nof_filler  = p.F;
  d.crc_poly    = static_cast<uint8_t>(p.crc_poly);
  d.max_it      = static_cast<uint8_t>(p.max_it);
  d.mode        = static_cast<uint8_t>(p.mode);
  d.new_data    = static_cast<uint8_t>(p.new_data ? 1 : 0);
  d.flags       = static_cast<uint8_t>(p.flags);
  d.scale_mult  = scale_mult(p.scaling);
  d.Qm          = static_cast<uint8_t>(p.Qm == 0 ? 1 : p.Qm);
  if (p.flags & FLAG_DEMATCH) {
    if (p.slot > h->nof_slots || (p.slot == h->nof_slots && (p.flags & FLAG_DECODE))) {
      h->last_error = "HARQ slot out of range";
      return SRSRAN_CUDA_ERR_INVALID;
    }
    uint32_t Ncb = p.Nref ? std::min(p.Nref, N) : N;
    if (p.F >= sys || Ncb <= sys || p.E == 0 || (p.E % d.Qm) != 0) {
      h->last_error = "invalid rate-dematching parameters";
      return SRSRAN_CUDA_ERR_INVALID;
    }
    d.Ncb = Ncb;
    d.k0  = compute_k0(p.bg, p.rv, Ncb, N, p.Z);
    c.extent_undo.emplace_back(p.slot, h->extent[p.slot]);
    h->extent[p.slot] =
        extent_after_dematch(h->extent[p.slot], N, Ncb, d.k0, p.E, sys - p.F, sys, p.new_data != 0);
  }
  if (!(p.flags & FLAG_DECODE)) {
    d.n_in     = 0;
    d.scan_len = 0;
  } else if (p.flags & FLAG_USE_HARQ) {
    d.n_in     = N;
    d.scan_len = std::min(N, (h->extent[p.slot] + 15) & ~15U);
  } else {
    if (p.n_in > N || p.n_in < K + 2 * p.Z) {
      h->last_error = "invalid decoder input length";
      return SRSRAN_CUDA_ERR_INVALID;
    }
    d.n_in     = p.n_in;
    d.scan_len = p.n_in;
  }
  d.layer_cap = layers_for_extent(d.scan_len, p.bg, p.Z);
  if (want_bits) {
    d.bits_out  = c.d_bits.p + static_cast<size_t>(idx) * BITS_STRIDE;
    c.want_bits = true;
  }
  c.h_desc.p[idx]  = d;
  c.h_tbmap.p[idx] = 0xffffffffU;
  if (p.flags & (FLAG_DEMATCH | FLAG_USE_HARQ | FLAG_TRACK_CRC)) {
    note_slot(c, p.slot);
  } else {
    // unit-level decode: scratch data-bit slots of this context only, no HARQ state
  }
  c.cb_meta.push_back({K, p.max_it, p.slot});
  *idx_out = idx;
  return SRSRAN_CUDA_OK;
}

int launch_context(srsran_cuda_pusch_dec* h, int ci)
{
  nvtx_scope nvtx_range("pusch_dec.launch_batch");
  batch_context& c   = h->ctx[ci];
  uint32_t       ncb = static_cast<uint32_t>(c.cb_meta.size());
  uint32_t       ntb = static_cast<uint32_t>(c.tb_meta.size());
  cudaStream_t   s   = c.stream;
  c.open             = false;
  h->open_ctx        = -1;
  if (ncb == 0) {
    return SRSRAN_CUDA_OK;
  }
  PROF_T(l0);
  // 1. Host -> device: LLRs (direct from pinned caller memory when possible), descriptors.
  if (h->timer_armed) {
    h->timer_armed = false;
    CUDA_TRY(h, cudaEventRecord(h->timer_begin, s));
  }
  CUDA_TRY(h, cudaEventRecord(c.stage[0], s));
  // The copies of this batch start when those of the previous batch are through: they share one PCIe link, and taking
  // turns lets the previous batch's kernels start (and overlap this batch's copies) as early as possible.
  if (h->last_launched >= 0 && h->last_launched != ci && !c.copies.empty()) {
    CUDA_TRY(h, cudaStreamWaitEvent(s, h->ctx[h->last_launched].copied, 0));
  }
  // Soft bits in many separate page-locked pieces (one buffer per decoder instance behind the plugin interface): one
  // gather kernel reads them over the link instead of one copy-engine job per piece (h2d_gather_kernel). It runs on the
  // batch's high-priority stream so that its few CTAs are placed ahead of the previous batch's pending decoder CTAs.
  bool gathered = h->h2d_gather_ctas != 0 && c.copies.size() >= H2D_GATHER_MIN_PIECES && c.copies.size() <= H2D_MAX_PIECES;
  for (size_t i = 0; gathered && i != c.copies.size(); ++i) {
    const batch_context::copy_job& j = c.copies[i];
    gathered = j.dev_src != nullptr && (reinterpret_cast<uintptr_t>(j.dev_src) % 16) == 0 && (j.dst_off % 16) == 0;
  }
  if (gathered) {
    CUDA_TRY(h, c.h_pieces.reserve(H2D_MAX_PIECES));
    for (size_t i = 0; i != c.copies.size(); ++i) {
      const batch_context::copy_job& j = c.copies[i];
      c.h_pieces.p[i] = {reinterpret_cast<const uint8_t*>(j.dev_src), reinterpret_cast<uint8_t*>(c.d_llr.p + j.dst_off), j.bytes};
    }
    CUDA_TRY(h, cudaEventRecord(c.fork, s));
    CUDA_TRY(h, cudaStreamWaitEvent(c.tail, c.fork, 0));
    h2d_gather_kernel<<<h->h2d_gather_ctas, H2D_THREADS, 0, c.tail>>>(c.h_pieces.p, static_cast<uint32_t>(c.copies.size()));
    ++h->launches;
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaEventRecord(c.join[0], c.tail));
    CUDA_TRY(h, cudaStreamWaitEvent(s, c.join[0], 0));
  } else {
    for (const batch_context::copy_job& j : c.copies) {
      const int8_t* src = (j.src != nullptr) ? j.src : c.h_llr.p + j.dst_off;
      CUDA_TRY(h, cudaMemcpyAsync(c.d_llr.p + j.dst_off, src, j.bytes, cudaMemcpyHostToDevice, s));
    }
  }
  for (const batch_context::raw_copy& j : c.raw_copies) {
    CUDA_TRY(h, cudaMemcpyAsync(c.d_dm_in.p + j.dst_off, j.src, j.bytes, cudaMemcpyHostToDevice, s));
  }
  if (c.ndm != 0) {
    // 1b. Soft bits born on the device: scrambling sequences (one warp per 32-Kbit chunk), then one thread per symbol.
    CUDA_TRY(h, cudaMemcpyAsync(c.d_dm.p, c.h_dm.p, c.ndm * sizeof(demod_desc), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaEventRecord(c.dm_ev[0], s));
    scr_seq_kernel<<<dim3((c.dm_max_words + SCR_CHUNK_WORDS - 1) / SCR_CHUNK_WORDS, c.ndm), 32, 0, s>>>(
        c.d_dm.p, h->d_scr_x1.p, h->d_scr_x2.p);
    ++h->launches;
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, launch_pusch_demod(c.d_dm.p, c.ndm, c.dm_max_sym, c.dm_qm_mask, s, &h->launches));
    CUDA_TRY(h, cudaEventRecord(c.dm_ev[1], s));
  }
  // A small batch without soft-bit copies (direct_in, or device-resident soft bits): the rate dematcher reads its
  // descriptors from the page-locked host array itself and starts at once, while the descriptors' device copies (two jobs of
  // ~6 us each, needed by the decoders and the TB assembly) travel on a side stream beside it.
  const bool     early_dm = h->direct_in && c.copies.empty() && c.raw_copies.empty() && c.ndm == 0 && ncb <= EARLY_DM_MAX_CBS;
  const cb_desc* dm_descs = early_dm ? c.h_desc.p : c.d_desc.p;
  // Everything the kernels of this batch wait for, then the rate dematcher (steps 3 and 4a).
  auto dematch_stage = [&](uint32_t stage_bytes, bool unstaged) -> int {
    CUDA_TRY(h, cudaEventRecord(c.copied, s));
    if (c.wait_for != nullptr) {
      CUDA_TRY(h, cudaStreamWaitEvent(s, c.wait_for, 0));
      c.wait_for = nullptr;
    }
    for (cudaEvent_t e : c.wait_events) {
      CUDA_TRY(h, cudaStreamWaitEvent(s, e, 0));
    }
    c.wait_events.clear();
    // 3. HARQ ordering: kernels of this context run after the kernels of the previously launched context.
    //    Batches whose HARQ slot ranges are disjoint share no state (soft bits, data bits, CRC flags are per slot; every other
    //    buffer is per context), so they may overlap: the next batch's dematching fills the tail of this batch's decoding.
    for (int k = 0; k != NOF_CONTEXTS; ++k) {
      batch_context& o = h->ctx[k];
      if (k == ci || !o.in_flight || o.slot_lo > o.slot_hi) {
        continue;
      }
      if (slots_overlap(c, o)) {
        CUDA_TRY(h, cudaStreamWaitEvent(s, o.kernels, 0));
      }
    }
    // 4. Kernels.
    CUDA_TRY(h, cudaEventRecord(c.stage[1], s));
    if (stage_bytes != 0) {
      // + 32: the de-interleaved image is read with aligned 32-bit loads that may run a few bytes past E.
      uint32_t smem = ((stage_bytes + 15) & ~15U) + 32;
      rate_dematch_kernel<true><<<ncb, DM_THREADS, smem, s>>>(dm_descs, h->d_soft.p, h->combine_block);
      ++h->launches;
      CUDA_TRY(h, cudaGetLastError());
    }
    if (unstaged) {
      rate_dematch_kernel<false><<<ncb, 256, 0, s>>>(dm_descs, h->d_soft.p, h->combine_block);
      ++h->launches;
      CUDA_TRY(h, cudaGetLastError());
    }
    CUDA_TRY(h, cudaEventRecord(c.stage[2], s));
    return SRSRAN_CUDA_OK;
  };
  if (early_dm) {
    // ... before the host groups the code blocks for the decoders (~10 us for one transport block of 152).
    uint32_t stage_bytes = 0;
    bool     unstaged    = false;
    for (uint32_t i = 0; i != ncb; ++i) {
      const cb_desc& d = c.h_desc.p[i];
      if (d.flags & FLAG_DEMATCH) {
        if (d.E <= DM_STAGE_CAP) {
          stage_bytes = std::max(stage_bytes, d.E);
        } else {
          unstaged = true;
        }
      }
    }
    int r = dematch_stage(stage_bytes, unstaged);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
  }
  PROF_T(l1);
  // 2. Group the decode operations into launch classes (threads per code block x shared-memory bucket).
  struct klass {
    int                   tpc;
    uint32_t              smem;
    std::vector<uint32_t> idx;
  };
  std::vector<klass> classes;
  uint32_t           dm_stage_bytes = 0; // largest E among the code blocks dematched through shared memory
  bool               dm_unstaged    = false;
  // Packed decoder: runs of consecutive code blocks with the same shape are grouped four per CTA.
  struct pklass {
    int      tpc;
    uint32_t lanes; // code blocks per CTA: 4, or 2 when the state of four does not fit in shared memory
    uint32_t smem;
    uint32_t first, count; // range in h_grp
    uint32_t z;            // lifting size (one per class: the Z = 384 specialisation of the kernel)
    uint32_t tm_cols;      // tensor-memory columns per CTA (0: messages in shared memory)
  };
  uint32_t grp_lanes = 4;
  bool     grp_tm    = false; // the open group runs on the tensor-memory variant
  std::vector<pklass> pclasses;
  uint32_t            ngrp     = 0;
  bool                grp_open = false;
  auto same_shape = [](const cb_desc& a, const cb_desc& b) {
    return a.bg == b.bg && a.Z == b.Z && a.mode == b.mode && a.max_it == b.max_it && a.scale_mult == b.scale_mult &&
           a.crc_poly == b.crc_poly;
  };
  std::vector<klass> qclasses; // ldpc_decode_q4_kernel: tpc = threads per CTA (32 / 64 / 96)
  // One code block per thread group: the packed single-code-block kernel where it applies, else ldpc_decode_kernel;
  // launch classes by threads per code block x shared-memory bucket.
  auto add_single = [&](uint32_t i) {
    const cb_desc&  d    = c.h_desc.p[i];
    if (q4_eligible(h, d)) {
      uint32_t need = (decq_smem_layout(d.bg, d.Z, d.layer_cap).total + 1023) & ~1023U;
      // Buckets keep the number of launches small; a finer grid for the small sizes, where occupancy matters.
      // 2 KB steps up to 64 KB (the sizes where one more CTA per SM is at stake), coarser above.
      uint32_t bsz = (need <= 64 * 1024) ? ((need + 2047) & ~2047U) : std::min<uint32_t>((need + 16383) & ~16383U, 227 * 1024);
      int    tp    = d.Z <= 128 ? 32 : (d.Z <= 256 ? 64 : 96);
      klass* found = nullptr;
      for (klass& k : qclasses) {
        if (k.tpc == tp && k.smem == bsz) {
          found = &k;
          break;
        }
      }
      if (found == nullptr) {
        qclasses.push_back({tp, bsz, {}});
        found = &qclasses.back();
      }
      found->idx.push_back(i);
      return true;
    }
    dec_smem_layout lay  = dec_layout(d.bg, d.Z, d.layer_cap);
    uint32_t        need = (lay.total + 1023) & ~1023U;
    // Buckets: coarse enough for few launches, fine enough for occupancy.
    static const uint32_t buckets[] = {16, 24, 32, 40, 48, 64, 80, 96, 112, 128, 160, 192, 227};
    uint32_t              bsz       = 227 * 1024;
    for (uint32_t b : buckets) {
      if (need <= b * 1024) {
        bsz = b * 1024;
        break;
      }
    }
    if (need > 227 * 1024) {
      return false;
    }
    int    tc    = tpc_class(d.Z);
    klass* found = nullptr;
    for (klass& k : classes) {
      if (k.tpc == tc && k.smem == bsz) {
        found = &k;
        break;
      }
    }
    if (found == nullptr) {
      classes.push_back({tc, bsz, {}});
      found = &classes.back();
    }
    found->idx.push_back(i);
    return true;
  };
  auto close_group = [&]() {
    if (!grp_open) {
      return;
    }
    grp_open           = false;
    const grp_desc& g  = c.h_grp.p[ngrp];
    if (g.n == 1 && add_single(g.cb[0])) {
      // A lone code block gains nothing from the packed kernels (they occupy a whole SM): general kernel.
      return;
    }
    const cb_desc&  d  = c.h_desc.p[g.cb[0]];
    const bool      lg = grp_tm && grp_lanes == 2; // many-layer form
    uint32_t        tm = lg ? 512U : (grp_tm ? packed_tmem_cols(h, d, g.layer_cap) : 0U);
    uint32_t        sm = lg ? packed_long_smem(h, d, g.layer_cap)
                            : dec4_smem_layout(d.bg, d.Z, g.layer_cap, grp_lanes, tm != 0).total;
    if (tm == 512 && h->use_bulk && sm + dec4_bulk_bytes(d.bg, d.Z, g.layer_cap, grp_lanes) <= static_cast<uint32_t>(h->max_smem_optin)) {
      // One CTA per SM anyway (all 512 columns of tensor memory): room to stage the inputs by bulk copies.
      sm += dec4_bulk_bytes(d.bg, d.Z, g.layer_cap, grp_lanes);
      tm |= DEC4_BULK_FLAG;
    }
    sm = (sm + 1023) & ~1023U;
    int             tp = d.Z <= 256 ? 256 : 384;
    if (pclasses.empty() || pclasses.back().tpc != tp || pclasses.back().smem != sm || pclasses.back().lanes != grp_lanes ||
        pclasses.back().z != d.Z || pclasses.back().tm_cols != tm) {
      pclasses.push_back({tp, grp_lanes, sm, ngrp, 0, d.Z, tm});
    }
    ++pclasses.back().count;
    ++ngrp;
  };
  // Latency mode: when four-per-CTA groups would leave more than half of the SMs idle (a single transport block), two
  // code blocks per CTA spread the batch over twice as many SMs and halve the work per thread.
  // Consecutive code blocks mostly repeat the previous one's shape (the code blocks of a transport block): the shape
  // dependent decisions are kept and reused while the shape fields compare equal.
  auto same_class = [&same_shape](const cb_desc& a, const cb_desc& b) {
    return same_shape(a, b) && a.flags == b.flags && a.layer_cap == b.layer_cap;
  };
  uint32_t nof_packable = 0, nof_long = 0;
  {
    const cb_desc* prev = nullptr;
    uint32_t       inc = 0, inc_long = 0;
    for (uint32_t i = 0; i != ncb; ++i) {
      const cb_desc& d = c.h_desc.p[i];
      if (prev == nullptr || !same_class(*prev, d)) {
        inc      = ((d.flags & FLAG_DECODE) && packed_eligible(h, d, d.layer_cap, 2)) ? 1U : 0U;
        inc_long = (inc == 0 && packed_long_smem(h, d, d.layer_cap) != 0) ? 1U : 0U;
        prev     = &d;
      }
      nof_packable += inc;
      nof_long += inc_long;
    }
  }
  // The many-layer pair form is taken wherever two such code blocks of the same shape follow each other (mixed-shape batches
  // are processed in shape order, see submit_batch): one check per thread instead of four makes a pair finish earlier than
  // two single-code-block CTAs even on an otherwise idle GPU (BASELINE config 3: 218 -> 148 us per slot, tools/c3_probe.py).
  // A lone code block still goes to the one-code-block kernel (close_group).
  const bool long_batch = nof_long != 0;
  const cb_desc* memo_desc  = nullptr;
  uint32_t       memo_lanes = 0;
  bool           memo_tm    = false;
  const bool small_batch = h->force_pairs || (nof_packable + 3) / 4 <= static_cast<uint32_t>(h->nof_sms) / 2;
  for (uint32_t i = 0; i != ncb; ++i) {
    const cb_desc& d = c.h_desc.p[i];
    if (d.flags & FLAG_DEMATCH) {
      if (d.E <= DM_STAGE_CAP) {
        dm_stage_bytes = std::max(dm_stage_bytes, d.E);
      } else {
        dm_unstaged = true;
      }
    }
    if (!(d.flags & FLAG_DECODE)) {
      continue;
    }
    if (memo_desc == nullptr || !same_class(*memo_desc, d)) {
      const bool inter_cb = !(h->prefer_q4 && q4_eligible(h, d));
      // Groups of four on the tensor-memory variant where it applies (also for shapes whose messages would not fit in
      // shared memory four at a time), else groups of four / two with the messages in shared memory.
      memo_tm    = inter_cb && !small_batch && packed_tmem_cols(h, d, d.layer_cap) != 0;
      memo_lanes = !inter_cb ? 0U : (memo_tm || (!small_batch && packed_eligible(h, d, d.layer_cap, 4)))
                                  ? 4U
                                  : (packed_eligible(h, d, d.layer_cap, 2) ? 2U : 0U);
      if (inter_cb && long_batch && (memo_lanes == 0 || (h->prefer_long && memo_lanes == 2 && !memo_tm)) &&
          packed_long_smem(h, d, d.layer_cap) != 0) {
        // Many layers: pairs with the messages in tensor memory instead of one code block per CTA.
        memo_lanes = 2;
        memo_tm    = true;
      }
      memo_desc  = &d;
    }
    const uint32_t lanes_fit = memo_lanes;
    if (lanes_fit != 0) {
      if (grp_open) {
        grp_desc&      g   = c.h_grp.p[ngrp];
        const cb_desc& f   = c.h_desc.p[g.cb[0]];
        uint32_t       cap = std::max(g.layer_cap, d.layer_cap);
        if (g.n < grp_lanes && lanes_fit == grp_lanes && memo_tm == grp_tm &&
            ((same_class(f, d) && g.layer_cap == d.layer_cap) ||
             (same_shape(f, d) &&
              (grp_tm ? (grp_lanes == 2 ? packed_long_smem(h, d, cap) != 0 : packed_tmem_cols(h, d, cap) != 0)
                      : packed_eligible(h, d, cap, grp_lanes))))) {
          g.cb[g.n++] = i;
          g.layer_cap = cap;
          continue;
        }
        close_group();
      }
      grp_desc& g = c.h_grp.p[ngrp];
      g           = {};
      g.cb[0]     = i;
      g.n         = 1;
      g.layer_cap = d.layer_cap;
      grp_lanes   = lanes_fit;
      grp_tm      = memo_tm;
      grp_open    = true;
      continue;
    }
    close_group();
    if (!add_single(i)) {
      // Nothing of this batch has reached the device yet: undo it (extents included).
      abandon_context(h, c);
      h->last_error = "code block does not fit in shared memory";
      return SRSRAN_CUDA_ERR_INVALID;
    }
  }
  close_group();
  uint32_t pos = 0;
  for (klass& k : classes) {
    for (uint32_t i : k.idx) {
      c.h_order.p[pos++] = i;
    }
  }
  for (klass& k : qclasses) {
    for (uint32_t i : k.idx) {
      c.h_order.p[pos++] = i;
    }
  }
  PROF_T(l2);
  // Device addresses of the small descriptor arrays: their own buffers, or - for a small batch, where every copy is ~5 us of
  // the latency - one packed region sent with a single copy.
  // Stream of the descriptor copies: with the dematcher already launched (early_dm) they leave the batch stream.
  cudaStream_t    ds       = early_dm ? c.side[0] : s;
  const grp_desc* dv_grp   = c.d_grp.p;
  const uint32_t* dv_order = c.d_order.p;
  const tb_desc*  dv_tb    = c.d_tb.p;
  const uint32_t* dv_tbmap = c.d_tbmap.p;
  {
    const size_t b_grp = (ngrp * sizeof(grp_desc) + 15) & ~size_t(15), b_ord = (pos * sizeof(uint32_t) + 15) & ~size_t(15);
    const size_t b_tb  = (ntb * sizeof(tb_desc) + 15) & ~size_t(15);
    const size_t b_map = (ntb != 0) ? ((ncb * sizeof(uint32_t) + 15) & ~size_t(15)) : 0;
    const size_t total = b_grp + b_ord + b_tb + b_map;
    CUDA_TRY(h, cudaMemcpyAsync(c.d_desc.p, c.h_desc.p, ncb * sizeof(cb_desc), cudaMemcpyHostToDevice, ds));
    if (total != 0 && total <= PACKED_META_MAX) {
      uint8_t* m = c.h_meta.p;
      std::memcpy(m, c.h_grp.p, ngrp * sizeof(grp_desc));
      std::memcpy(m + b_grp, c.h_order.p, pos * sizeof(uint32_t));
      if (ntb != 0) {
        std::memcpy(m + b_grp + b_ord, c.h_tb.p, ntb * sizeof(tb_desc));
        std::memcpy(m + b_grp + b_ord + b_tb, c.h_tbmap.p, ncb * sizeof(uint32_t));
      }
      CUDA_TRY(h, cudaMemcpyAsync(c.d_meta.p, m, total, cudaMemcpyHostToDevice, ds));
      dv_grp   = reinterpret_cast<const grp_desc*>(c.d_meta.p);
      dv_order = reinterpret_cast<const uint32_t*>(c.d_meta.p + b_grp);
      dv_tb    = reinterpret_cast<const tb_desc*>(c.d_meta.p + b_grp + b_ord);
      dv_tbmap = reinterpret_cast<const uint32_t*>(c.d_meta.p + b_grp + b_ord + b_tb);
    } else {
      if (ngrp != 0) {
        CUDA_TRY(h, cudaMemcpyAsync(c.d_grp.p, c.h_grp.p, ngrp * sizeof(grp_desc), cudaMemcpyHostToDevice, ds));
      }
      if (pos != 0) {
        CUDA_TRY(h, cudaMemcpyAsync(c.d_order.p, c.h_order.p, pos * sizeof(uint32_t), cudaMemcpyHostToDevice, ds));
      }
      if (ntb != 0) {
        CUDA_TRY(h, cudaMemcpyAsync(c.d_tb.p, c.h_tb.p, ntb * sizeof(tb_desc), cudaMemcpyHostToDevice, ds));
        CUDA_TRY(h, cudaMemcpyAsync(c.d_tbmap.p, c.h_tbmap.p, ncb * sizeof(uint32_t), cudaMemcpyHostToDevice, ds));
      }
    }
  }
  if (early_dm) {
    CUDA_TRY(h, cudaEventRecord(c.join[0], ds));
  }
  cb_result*     dv_res   = c.d_res.p;
  tb_result_dev* dv_tbres = c.d_tbres.p;
  uint8_t*       dv_tbout = c.d_tbout.p;
  c.r_res                 = c.h_res.p;
  c.r_tbres               = c.h_tbres.p;
  c.r_tbout               = c.h_tbout.p;
  size_t packed_out       = 0; // bytes of the packed result region (0: separate copies)
  bool   direct_out       = false;
  cb_result* dv_res_host  = nullptr;
  {
    const size_t b_tbres = (ntb * sizeof(tb_result_dev) + 15) & ~size_t(15), b_res = (ncb * sizeof(cb_result) + 15) & ~size_t(15);
    const size_t total   = b_tbres + b_res + c.tbout_used;
    if (ntb != 0 && h->tb_host_copy && !c.want_bits && total <= PACKED_OUT_MAX) {
      // The kernels only ever WRITE these three arrays (every element once): for such a small batch they write them
      // straight into the page-locked result buffer (mapped into the device's address space) and no copy follows.
      direct_out          = h->direct_out;
      uint8_t* const outp = direct_out ? c.h_outp.p : c.d_outp.p;
      dv_tbres            = reinterpret_cast<tb_result_dev*>(outp);
      dv_res              = reinterpret_cast<cb_result*>(c.d_outp.p + b_tbres); // decoders write device memory ...
      dv_res_host         = direct_out ? reinterpret_cast<cb_result*>(outp + b_tbres) : nullptr; // ... the TB assembly forwards it
      dv_tbout            = outp + b_tbres + b_res;
      c.r_tbres  = reinterpret_cast<const tb_result_dev*>(c.h_outp.p);
      c.r_res    = reinterpret_cast<const cb_result*>(c.h_outp.p + b_tbres);
      c.r_tbout  = c.h_outp.p + b_tbres + b_res;
      packed_out = total;
    }
  }
  c.r_tbout_dev = dv_tbout;
  PROF_T(l3);
  if (!early_dm) {
    int r = dematch_stage(dm_stage_bytes, dm_unstaged);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
  } else {
    CUDA_TRY(h, cudaStreamWaitEvent(s, c.join[0], 0)); // the decoders and the TB assembly read the descriptors' device copy
  }
  // Launch classes are independent (disjoint code blocks): class 0 stays on the batch stream, the others fork onto side
  // streams and join before the TB assembly, so a slot of mixed small transport blocks costs its slowest class, not the sum.
  const size_t nof_classes = pclasses.size() + classes.size() + qclasses.size();
  uint32_t     side_used   = 0;
  size_t       class_no    = 0;
  if (nof_classes > 1) {
    CUDA_TRY(h, cudaEventRecord(c.fork, s));
  }
  auto class_stream = [&]() -> cudaStream_t {
    size_t k = class_no++;
    if (k == 0) {
      return s;
    }
    int          si = static_cast<int>((k - 1) % batch_context::NOF_SIDE);
    cudaStream_t st = c.side[si];
    if (!(side_used & (1U << si))) {
      side_used |= 1U << si;
      cudaStreamWaitEvent(st, c.fork, 0);
    }
    return st;
  };
  for (const pklass& k : pclasses) {
    const grp_desc* grp = dv_grp + k.first;
    cudaStream_t    st  = class_stream();
    cudaError_t     e;
    if (k.lanes == 2 && k.tm_cols != 0) {
      e = (k.tpc == 256)
              ? launch_decode2t<256>(h, st, c.d_desc.p, grp, dv_res, c.bits_base, k.count, k.smem, k.tm_cols)
              : launch_decode2t<384>(h, st, c.d_desc.p, grp, dv_res, c.bits_base, k.count, k.smem, k.tm_cols, k.z == 384);
    } else if (k.lanes == 2) {
      e = (k.tpc == 256) ? launch_decode2<256>(h, st, c.d_desc.p, grp, dv_res, c.bits_base, k.count, k.smem)
                         : launch_decode2<384>(h, st, c.d_desc.p, grp, dv_res, c.bits_base, k.count, k.smem, k.z == 384);
    } else if (k.tm_cols != 0) {
      e = (k.tpc == 256)
              ? launch_decode4t<256>(h, st, c.d_desc.p, grp, dv_res, c.bits_base, k.count, k.smem, k.tm_cols)
              : launch_decode4t<384>(h, st, c.d_desc.p, grp, dv_res, c.bits_base, k.count, k.smem, k.tm_cols, k.z == 384);
    } else {
      e = (k.tpc == 256) ? launch_decode4<256>(h, st, c.d_desc.p, grp, dv_res, c.bits_base, k.count, k.smem)
                         : launch_decode4<384>(h, st, c.d_desc.p, grp, dv_res, c.bits_base, k.count, k.smem, k.z == 384);
    }
    CUDA_TRY(h, e);
  }
  pos = 0;
  for (klass& k : classes) {
    uint32_t        n   = static_cast<uint32_t>(k.idx.size());
    const uint32_t* ord = dv_order + pos;
    cudaStream_t    st  = class_stream();
    cudaError_t     e   = cudaSuccess;
    switch (k.tpc) {
      case 0:
        e = launch_decode<32, 4>(h, st, c.d_desc.p, ord, dv_res, c.bits_base, n, k.smem);
        break;
      case 1:
        e = launch_decode<64, 2>(h, st, c.d_desc.p, ord, dv_res, c.bits_base, n, k.smem);
        break;
      case 2:
        e = launch_decode<128, 1>(h, st, c.d_desc.p, ord, dv_res, c.bits_base, n, k.smem);
        break;
      case 3:
        e = launch_decode<256, 1>(h, st, c.d_desc.p, ord, dv_res, c.bits_base, n, k.smem);
        break;
      default:
        e = launch_decode<384, 1>(h, st, c.d_desc.p, ord, dv_res, c.bits_base, n, k.smem);
        break;
    }
    CUDA_TRY(h, e);
    pos += n;
  }
  for (klass& k : qclasses) {
    uint32_t        n   = static_cast<uint32_t>(k.idx.size());
    const uint32_t* ord = dv_order + pos;
    cudaStream_t    st  = class_stream();
    cudaError_t     e   = (k.tpc == 32) ? launch_decode_q4<32>(h, st, c.d_desc.p, ord, dv_res, c.bits_base, n, k.smem)
                          : (k.tpc == 64) ? launch_decode_q4<64>(h, st, c.d_desc.p, ord, dv_res, c.bits_base, n, k.smem)
                                          : launch_decode_q4<96>(h, st, c.d_desc.p, ord, dv_res, c.bits_base, n, k.smem);
    CUDA_TRY(h, e);
    pos += n;
  }
  for (int si = 0; si != batch_context::NOF_SIDE; ++si) {
    if (side_used & (1U << si)) {
      CUDA_TRY(h, cudaEventRecord(c.join[si], c.side[si]));
      CUDA_TRY(h, cudaStreamWaitEvent(s, c.join[si], 0));
    }
  }
  CUDA_TRY(h, cudaEventRecord(c.stage[3], s));
  // The small end-of-batch kernels and the result copies go to a high-priority stream: when the next batch overlaps this
  // one, its decode CTAs must not keep these few CTAs (and with them the completion of this batch) waiting.
  // A small batch (the latency case) stays on its own stream: the hop to another stream costs more than it saves there.
  cudaStream_t ts = (packed_out != 0) ? s : c.tail;
  CUDA_TRY(h, cudaEventRecord(c.decoded, s));
  if (ts != s) {
    CUDA_TRY(h, cudaStreamWaitEvent(ts, c.decoded, 0));
  }
  if (ntb != 0) {
    // One warp per code block; the last warp of a transport block also finalises it (TB CRC verdict, CRC-flag reset).
    tb_gather_kernel<<<(ncb + TBG_WARPS - 1) / TBG_WARPS, TBG_WARPS * 32, 0, ts>>>(
        dv_tb, c.d_desc.p, dv_tbmap, ncb, h->d_bits.p, dv_tbout, c.d_tbshare.p, dv_tbres, h->d_crc_flags.p, c.d_tbdone.p,
        dv_res, dv_res_host);
    ++h->launches;
    CUDA_TRY(h, cudaGetLastError());
  }
  CUDA_TRY(h, cudaEventRecord(c.kernels, ts));
  // 5. Device -> host.
  if (packed_out != 0) {
    c.tb_on_host = true;
    if (!direct_out) {
      CUDA_TRY(h, cudaMemcpyAsync(c.h_outp.p, c.d_outp.p, packed_out, cudaMemcpyDeviceToHost, ts));
    }
  } else {
  CUDA_TRY(h, cudaMemcpyAsync(c.h_res.p, c.d_res.p, ncb * sizeof(cb_result), cudaMemcpyDeviceToHost, ts));
  if (c.want_bits) {
    CUDA_TRY(h, cudaMemcpyAsync(c.h_bits.p, c.d_bits.p, static_cast<size_t>(ncb) * BITS_STRIDE, cudaMemcpyDeviceToHost, ts));
  }
  if (ntb != 0) {
    CUDA_TRY(h, cudaMemcpyAsync(c.h_tbres.p, c.d_tbres.p, ntb * sizeof(tb_result_dev), cudaMemcpyDeviceToHost, ts));
    c.tb_on_host = h->tb_host_copy;
    if (c.tb_on_host) {
      CUDA_TRY(h, cudaMemcpyAsync(c.h_tbout.p, c.d_tbout.p, c.tbout_used, cudaMemcpyDeviceToHost, ts));
    }
  }
  }
  CUDA_TRY(h, cudaEventRecord(c.done, ts));
  // The batch stream joins the tail so that the next use of this context (and stream-ordered waits on it) see it complete.
  if (ts != s) {
    CUDA_TRY(h, cudaStreamWaitEvent(s, c.done, 0));
  }
  PROF_T(l4);
  PROF_ADD(2, l0, l1);
  PROF_ADD(3, l1, l2);
  PROF_ADD(4, l2, l3);
  PROF_ADD(5, l3, l4);
  c.extent_undo.clear(); // the dematcher of this batch is on the device: the extents stand
  c.in_flight      = true;
  h->last_launched = ci;
  return SRSRAN_CUDA_OK;
}

int ensure_bits_buffers(srsran_cuda_pusch_dec* h, batch_context& c)
{
  CUDA_TRY(h, c.h_bits.reserve(c.h_desc.cap * BITS_STRIDE));
  CUDA_TRY(h, c.d_bits.reserve(c.h_desc.cap * BITS_STRIDE + 16));
  return SRSRAN_CUDA_OK;
}

/// Segmentation (ldpc_segmenter_impl.cpp:58-68,254-331 and ldpc.h:128-228), host arithmetic.
int segment(uint32_t tbs, uint32_t bg, uint32_t Qm, uint32_t nof_layers, uint32_t nof_llrs,
            srsran_cuda_pusch_dec_cb_meta* out)
{
  static const uint16_t ALL_Z[51] = {2,  3,  4,  5,  6,  7,  8,  9,   10,  11,  12,  13,  14,  15,  16,  18,  20,
                                     22, 24, 26, 28, 30, 32, 36, 40,  44,  48,  52,  56,  60,  64,  72,  80,  88,
                                     96, 104, 112, 120, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288, 320, 352, 384};
  if ((bg != 1 && bg != 2) || Qm == 0 || nof_layers == 0 || tbs == 0) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  uint32_t tb_crc  = (tbs <= 3824) ? 16 : 24;
  uint32_t B       = tbs + tb_crc;
  uint32_t max_seg = (bg == 1) ? 8448 : 3840;
  uint32_t C       = (B <= max_seg) ? 1 : (B + (max_seg - 24) - 1) / (max_seg - 24);
  if (C > SRSRAN_CUDA_MAX_NOF_SEGMENTS) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  uint32_t Bp      = B + ((C > 1) ? 24 * C : 0);
  uint32_t ref_len = 22;
  if (bg == 2) {
    ref_len = (B > 640) ? 10 : (B > 560) ? 9 : (B > 192) ? 8 : 6;
  }
  uint32_t Z = 0;
  for (uint16_t z : ALL_Z) {
    if (z * C * ref_len >= Bp) {
      Z = z;
      break;
    }
  }
  if (Z == 0) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  uint32_t K        = ((bg == 1) ? 22 : 10) * Z;
  uint32_t cb_crc   = (C > 1) ? 24 : 0;
  uint32_t max_info = (Bp + C - 1) / C - cb_crc;
  uint32_t nsl      = (nof_llrs / Qm) / nof_layers;
  uint32_t nshort   = C - (nsl % C);
  uint32_t off      = 0;
  for (uint32_t i = 0; i != C; ++i) {
    uint32_t per = (i < nshort) ? nsl / C : (nsl + C - 1) / C;
    uint32_t E   = per * nof_layers * Qm;
    out[i]       = {bg, Z, K * ((bg == 1) ? 3U : 5U), E, K - (max_info + cb_crc), off, (C == 1) ? tb_crc : cb_crc};
    off += E;
  }
  return (off == nof_llrs) ? static_cast<int>(C) : SRSRAN_CUDA_ERR_INVALID;
}

/// Adds all code blocks of one TB to the open context. `llr_dev` = device address of the TB's LLRs.
int add_tb(srsran_cuda_pusch_dec* h, batch_context& c, const srsran_cuda_pusch_dec_tb_config& cfg, const int8_t* llr_dev,
           uint32_t nof_llrs, const uint32_t* cb_slots = nullptr, uint32_t nof_cb_slots = 0)
{
  srsran_cuda_pusch_dec_cb_meta metas[SRSRAN_CUDA_MAX_NOF_SEGMENTS];
  if (cfg.tbs_bits % 8 != 0) {
    h->last_error = "TBS must be a multiple of 8";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  uint32_t Qm = cfg.modulation == 0 ? 1 : cfg.modulation;
  int      C  = segment(cfg.tbs_bits, cfg.base_graph, Qm, cfg.nof_layers, nof_llrs, metas);
  if (C < 0) {
    h->last_error = "segmentation failed (inconsistent TBS / number of LLRs)";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  if (c.tb_meta.size() >= MAX_TBS_PER_CTX) {
    h->last_error = "too many transport blocks in one batch";
    return SRSRAN_CUDA_ERR_STATE;
  }
  if (cb_slots != nullptr && nof_cb_slots < static_cast<uint32_t>(C)) {
    h->last_error = "fewer HARQ code-block ids than code blocks";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  // pusch_decoder_impl.cpp:35-46.
  uint32_t crc_poly = (C > 1) ? SRSRAN_CUDA_CRC24B : ((cfg.tbs_bits > 3824) ? SRSRAN_CUDA_CRC24A : SRSRAN_CUDA_CRC16);
  uint32_t first_cb = static_cast<uint32_t>(c.cb_meta.size());
  // The code blocks of a transport block differ only in their LLRs, their HARQ slot and (two values) their length E: when
  // E and the slot's soft-buffer extent equal those of the previous code block, its descriptor is copied and patched.
  bool     have_prev = false;
  uint32_t prev_E = 0, prev_F = 0, prev_ext_before = 0, prev_ext_after = 0;
  for (int i = 0; i != C; ++i) {
    const uint32_t slot = (cb_slots != nullptr) ? cb_slots[i] : cfg.harq_first_slot + i;
    if (have_prev && metas[i].rm_length == prev_E && metas[i].nof_filler_bits == prev_F && slot < h->nof_slots &&
        h->extent[slot] == prev_ext_before) {
      const uint32_t idx = static_cast<uint32_t>(c.cb_meta.size());
      if (idx >= c.h_desc.cap) {
        h->last_error = "batch context full";
        return SRSRAN_CUDA_ERR_STATE;
      }
      cb_desc d         = c.h_desc.p[idx - 1];
      d.llr             = llr_dev + metas[i].cw_offset;
      d.slot            = slot;
      c.h_desc.p[idx]   = d;
      c.extent_undo.emplace_back(slot, h->extent[slot]);
      h->extent[slot]   = prev_ext_after;
      note_slot(c, slot);
      const cb_host_meta pm = c.cb_meta.back();
      c.cb_meta.push_back({pm.K, pm.max_it, slot});
      continue;
    }
    prev_ext_before = (slot < h->nof_slots) ? h->extent[slot] : 0;
    cb_params p = {};
    p.bg        = cfg.base_graph;
    p.Z         = metas[i].lifting_size;
    p.E         = metas[i].rm_length;
    p.rv        = cfg.rv;
    p.Qm        = Qm;
    p.Nref      = cfg.Nref;
    p.F         = metas[i].nof_filler_bits;
    p.crc_poly  = crc_poly;
    p.max_it    = cfg.nof_ldpc_iterations;
    p.mode      = cfg.use_early_stop ? MODE_EARLY_STOP : MODE_CRC_AT_END;
    p.new_data  = cfg.new_data;
    p.slot      = slot;
    p.flags     = FLAG_DEMATCH | FLAG_DECODE | FLAG_USE_HARQ | FLAG_TRACK_CRC;
    p.scaling   = 0.8F; // pusch_codeblock_decoder.cpp:47-50 keeps the default scaling factor
    uint32_t idx;
    int      r = add_cb(h, c, p, llr_dev + metas[i].cw_offset, false, &idx);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
    have_prev      = true;
    prev_E         = p.E;
    prev_F         = p.F;
    prev_ext_after = h->extent[slot];
  }
  uint32_t K        = metas[0].full_length / ((cfg.base_graph == 1) ? 3 : 5);
  tb_desc  t        = {};
  t.first_cb        = first_cb;
  t.nof_cbs         = C;
  t.first_slot      = (cb_slots != nullptr) ? cb_slots[0] : cfg.harq_first_slot;
  t.tbs_bits        = cfg.tbs_bits;
  t.cb_data_bits    = K - metas[0].nof_crc_bits - metas[0].nof_filler_bits;
  t.out_offset      = static_cast<uint32_t>(c.tbout_used);
  uint32_t ti       = static_cast<uint32_t>(c.tb_meta.size());
  c.h_tb.p[ti]      = t;
  for (int i = 0; i != C; ++i) {
    c.h_tbmap.p[first_cb + i] = ti;
  }
  c.tb_meta.push_back({first_cb, static_cast<uint32_t>(C), cfg.tbs_bits, t.out_offset, cfg.nof_ldpc_iterations, false});
  c.tbout_used += (cfg.tbs_bits / 8 + 3 + 15) & ~size_t(15);
  return SRSRAN_CUDA_OK;
}

int prepare_tb_buffers(srsran_cuda_pusch_dec* h, batch_context& c, uint32_t nof_tbs, size_t tb_bytes_total)
{
  CUDA_TRY(h, c.h_tb.reserve(MAX_TBS_PER_CTX));
  CUDA_TRY(h, c.d_tb.reserve(MAX_TBS_PER_CTX));
  CUDA_TRY(h, c.h_tbres.reserve(MAX_TBS_PER_CTX));
  if (c.d_tbdone.p == nullptr) {
    // Per transport block: code blocks assembled so far (tb_gather_kernel; returns to zero by itself).
    CUDA_TRY(h, c.d_tbdone.reserve(MAX_TBS_PER_CTX));
    CUDA_TRY(h, cudaMemset(c.d_tbdone.p, 0, MAX_TBS_PER_CTX * sizeof(uint32_t)));
  }
  CUDA_TRY(h, c.d_tbres.reserve(MAX_TBS_PER_CTX));
  size_t need = c.tbout_used + tb_bytes_total + 32 * static_cast<size_t>(nof_tbs);
  if (need > c.h_tbout.cap) {
    if (c.tbout_used != 0) {
      h->last_error = "TB output staging exhausted";
      return SRSRAN_CUDA_ERR_STATE;
    }
    CUDA_TRY(h, c.h_tbout.reserve(need));
    CUDA_TRY(h, c.d_tbout.reserve(need));
  }
  return SRSRAN_CUDA_OK;
}

/// Packed words of a 31-stage binary recurrence s(n + 31) = XOR of s(n + t), t in `taps`, from the initial state `init`
/// (bit i = s(i)), starting at bit 1600: word w bit j = s(1600 + 32 w + j). The first 31 words are produced bit by bit, the
/// rest with the word-level form of the same recurrence (Frobenius: P(x)^32 = P(x^32) over GF(2)).
void lfsr_words(uint32_t init, uint32_t taps, uint32_t* words, uint32_t nwords)
{
  uint32_t st = init & 0x7fffffffU;
  auto     step = [&]() {
    uint32_t f = __builtin_popcount(st & taps) & 1U;
    uint32_t o = st & 1U;
    st         = (st >> 1) | (f << 30);
    return o;
  };
  for (int i = 0; i != 1600; ++i) {
    step();
  }
  for (uint32_t w = 0; w != std::min<uint32_t>(31, nwords); ++w) {
    uint32_t v = 0;
    for (int j = 0; j != 32; ++j) {
      v |= step() << j;
    }
    words[w] = v;
  }
  for (uint32_t w = 31; w < nwords; ++w) {
    uint32_t v = 0;
    for (int t = 0; t != 4; ++t) {
      if ((taps >> t) & 1U) {
        v ^= words[w - 31 + t];
      }
    }
    words[w] = v;
  }
}

/// Builds the demodulator's constant tables (the reference's binary32 expressions) and the scrambling-sequence tables.
int init_demod(srsran_cuda_pusch_dec* h)
{
  if (h->demod_ready) {
    return SRSRAN_CUDA_OK;
  }
  demod_tables t = {};
  auto fill = [&](int ti, float unit, float width_mult, int n, const int* slope_mult, const float* icpt_num, float icpt_den) {
    t.width[ti] = width_mult * unit;
    t.inv[ti]   = 1.0F / t.width[ti];
    t.n[ti]     = n;
    for (int i = 0; i != n; ++i) {
      t.slope[ti][i] = static_cast<float>(slope_mult[i]) * unit;
      t.icpt[ti][i]  = icpt_num[i] / icpt_den;
    }
  };
  const float s42 = 1.0F / std::sqrt(42.0F), s170 = 1.0F / std::sqrt(170.0F);
  {
    // demodulation_mapper_qam64.cpp:42-84
    static const int   s01[8] = {16, 12, 8, 4, 4, 8, 12, 16};
    static const float i01[8] = {24, 12, 4, 0, 0, -4, -12, -24};
    static const int   s23[8] = {8, 4, 4, 8, -8, -4, -4, -8};
    static const float i23[8] = {20, 8, 8, 12, 12, 8, 8, 20};
    static const int   s45[4] = {4, -4, 4, -4};
    static const float i45[4] = {12, -4, -4, 12};
    fill(0, s42, 2, 8, s01, i01, 21);
    fill(1, s42, 2, 8, s23, i23, 21);
    fill(2, s42, 4, 4, s45, i45, 21);
  }
  {
    // demodulation_mapper_qam256.cpp:42-172
    static const int   s01[16] = {32, 28, 24, 20, 16, 12, 8, 4, 4, 8, 12, 16, 20, 24, 28, 32};
    static const float i01[16] = {112, 84, 60, 40, 24, 12, 4, 0, 0, -4, -12, -24, -40, -60, -84, -112};
    static const int   s23[16] = {16, 12, 8, 4, 4, 8, 12, 16, -16, -12, -8, -4, -4, -8, -12, -16};
    static const float i23[16] = {88, 60, 36, 16, 16, 28, 36, 40, 40, 36, 28, 16, 16, 36, 60, 88};
    static const int   s45[16] = {8, 4, 4, 8, -8, -4, -4, -8, 8, 4, 4, 8, -8, -4, -4, -8};
    static const float i45[16] = {52, 24, 24, 44, -20, -8, -8, -12, -12, -8, -8, -20, 44, 24, 24, 52};
    static const int   s67[8]  = {4, -4, 4, -4, 4, -4, 4, -4};
    static const float i67[8]  = {28, -20, 12, -4, -4, 12, -20, 28};
    fill(3, s170, 2, 16, s01, i01, 85);
    fill(4, s170, 2, 16, s23, i23, 85);
    fill(5, s170, 2, 16, s45, i45, 85);
    fill(6, s170, 4, 8, s67, i67, 85);
  }
  t.sq10  = 1.0F / std::sqrt(10.0F);
  t.gain2 = 2.0F * 1.41421356237309504880F;
  CUDA_TRY(h, cudaMemcpyToSymbol(c_demod, &t, sizeof(t)));
  CUDA_TRY(h, cudaMemcpyToSymbol(g_demod_slope_icpt, &t, 2 * 7 * 16 * sizeof(float))); // slope[7][16] then icpt[7][16]
  // Scrambling sequences (TS 38.211 5.2.1): x1 from state 1 (taps n, n + 3), x2 from every single bit of c_init (taps
  // n .. n + 3); the x2 sequence of any c_init is the XOR of the basis sequences of its set bits.
  std::vector<uint32_t> seq(SCR_MAX_WORDS);
  std::vector<uint32_t> win(static_cast<size_t>(SCR_NOF_CHUNKS) * 31 * 32, 0);
  CUDA_TRY(h, h->d_scr_x1.reserve(SCR_MAX_WORDS));
  CUDA_TRY(h, h->d_scr_x2.reserve(win.size()));
  lfsr_words(1U, 0x9U, seq.data(), SCR_MAX_WORDS);
  CUDA_TRY(h, cudaMemcpy(h->d_scr_x1.p, seq.data(), SCR_MAX_WORDS * sizeof(uint32_t), cudaMemcpyHostToDevice));
  for (uint32_t bit = 0; bit != 31; ++bit) {
    lfsr_words(1U << bit, 0xfU, seq.data(), SCR_MAX_WORDS);
    for (uint32_t ch = 0; ch != SCR_NOF_CHUNKS; ++ch) {
      for (uint32_t l = 0; l != 31; ++l) {
        uint32_t w = ch * SCR_CHUNK_WORDS + l;
        win[(static_cast<size_t>(ch) * 31 + bit) * 32 + l] = (w < SCR_MAX_WORDS) ? seq[w] : 0U;
      }
    }
  }
  CUDA_TRY(h, cudaMemcpy(h->d_scr_x2.p, win.data(), win.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  h->demod_ready = true;
  return SRSRAN_CUDA_OK;
}

/// Validates a demodulation configuration; returns the number of symbols (REs x layers) or 0.
uint32_t demod_nof_symbols(const srsran_cuda_pusch_demod_config& d)
{
  if ((d.modulation != 1 && d.modulation != 2 && d.modulation != 4 && d.modulation != 6 && d.modulation != 8) ||
      d.nof_layers == 0 || d.nof_layers > 4 || d.nof_ofdm_symbols == 0 || d.nof_ofdm_symbols > DM_MAX_OFDM ||
      (d.pi2_bpsk != 0 && d.modulation != 1) || d.n_id > 1023 || d.rnti > 65535) {
    return 0;
  }
  uint64_t re = 0;
  for (uint32_t i = 0; i != d.nof_ofdm_symbols; ++i) {
    re += d.re_per_symbol[i];
  }
  uint64_t nsym = re * d.nof_layers;
  if (nsym == 0 || nsym * d.modulation > static_cast<uint64_t>(SCR_MAX_WORDS) * 32) {
    return 0;
  }
  return static_cast<uint32_t>(nsym);
}

/// Appends one codeword to the open context's demodulation list. `sym_dev` / `nv_dev` / `llr_dev`: device addresses.
int add_demod(srsran_cuda_pusch_dec* h, batch_context& c, const srsran_cuda_pusch_demod_config& d, uint32_t nsym,
              const float* sym_dev, const float* nv_dev, int8_t* llr_dev, size_t scr_word_off)
{
  (void)h;
  demod_desc& o    = c.h_dm.p[c.ndm++];
  o                = {};
  o.sym            = reinterpret_cast<const float2*>(sym_dev);
  o.nv             = nv_dev;
  o.llr            = llr_dev;
  o.scr            = c.d_scr.p + scr_word_off;
  o.nsym           = nsym;
  o.qm             = d.modulation;
  o.pi2            = d.pi2_bpsk;
  o.nl             = d.nof_layers;
  o.max_block_subc = 4096 / (d.nof_layers * d.modulation); // pusch_demodulator_impl.h:71, .cpp:177
  o.mbs_magic      = static_cast<uint32_t>((0x100000000ULL + o.max_block_subc - 1) / o.max_block_subc);
  o.c_init         = (d.rnti << 15) + d.n_id;              // pusch_demodulator_impl.cpp:139
  o.nof_ofdm       = d.nof_ofdm_symbols;
  uint32_t acc     = 0;
  for (uint32_t i = 0; i != d.nof_ofdm_symbols; ++i) {
    o.re_start[i] = acc;
    acc += d.re_per_symbol[i];
  }
  for (uint32_t i = d.nof_ofdm_symbols; i <= DM_MAX_OFDM; ++i) {
    o.re_start[i] = acc;
  }
  c.dm_qm_mask |= 1U << d.modulation;
  c.dm_max_sym   = std::max(c.dm_max_sym, nsym);
  c.dm_max_words = std::max(c.dm_max_words, (nsym * d.modulation + 31) / 32);
  return SRSRAN_CUDA_OK;
}

int make_ticket(int ctx, uint32_t tb_index, uint32_t generation)
{
  return static_cast<int>(((generation & 0x3ff) << 20) | (static_cast<uint32_t>(ctx) << 16) | tb_index);
}

} // namespace

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" {

int srsran_cuda_pusch_dec_create(int device, uint32_t max_cbs_in_flight, uint32_t nof_harq_cb_slots,
                                 srsran_cuda_pusch_dec_t** handle)
{
  if (handle == nullptr || max_cbs_in_flight == 0 || nof_harq_cb_slots == 0) {
    g_create_error = "invalid arguments";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  *handle   = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
    g_create_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "bad index");
    cudaGetLastError();
    return SRSRAN_CUDA_ERR_NO_DEVICE;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    g_create_error = "cudaSetDevice failed";
    return SRSRAN_CUDA_ERR_NO_DEVICE;
  }
  srsran_cuda_pusch_dec* h = new (std::nothrow) srsran_cuda_pusch_dec();
  if (h == nullptr) {
    return SRSRAN_CUDA_ERR_NO_MEMORY;
  }
  h->device    = device;
  h->max_cbs   = max_cbs_in_flight;
  h->nof_slots = nof_harq_cb_slots;
  auto fail    = [&](int code) {
    g_create_error = h->last_error;
    srsran_cuda_pusch_dec_destroy(h);
    return code;
  };
  cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaDeviceGetAttribute(&h->nof_sms, cudaDevAttrMultiProcessorCount, device);
  if (upload_tables(h) != SRSRAN_CUDA_OK) {
    return fail(SRSRAN_CUDA_ERR_CUDA);
  }
  int smem = std::min(h->max_smem_optin, 227 * 1024);
  if (set_smem_attr<32, 4>(smem) != cudaSuccess || set_smem_attr<64, 2>(smem) != cudaSuccess ||
      set_smem_attr<128, 1>(smem) != cudaSuccess || set_smem_attr<256, 1>(smem) != cudaSuccess ||
      set_smem_attr<384, 1>(smem) != cudaSuccess) {
    h->last_error = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
    return fail(SRSRAN_CUDA_ERR_CUDA);
  }
  if (cudaFuncSetAttribute(ldpc_decode4_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode4_kernel<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode4_kernel<384, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode4_kernel<256, 0, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode4_kernel<384, 0, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode4_kernel<384, 384, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode4_kernel<256, 0, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode4_kernel<384, 0, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode4_kernel<384, 384, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode_q4_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode_q4_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode_q4_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode4_kernel<256, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode4_kernel<384, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(ldpc_decode4_kernel<384, 384, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
    h->last_error = "cudaFuncSetAttribute(ldpc_decode4_kernel) failed";
    return fail(SRSRAN_CUDA_ERR_CUDA);
  }
  h->max_smem_optin = smem;
  if (cudaFuncSetAttribute(rate_dematch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(DM_STAGE_CAP + 48)) != cudaSuccess) {
    h->last_error = "cudaFuncSetAttribute(rate_dematch_kernel) failed";
    return fail(SRSRAN_CUDA_ERR_CUDA);
  }
  size_t soft_bytes = (static_cast<size_t>(nof_harq_cb_slots) + 1) * SOFT_STRIDE; // + 1 scratch slot (unit-level)
  size_t bits_bytes = static_cast<size_t>(nof_harq_cb_slots) * BITS_STRIDE + 16;
  if (h->d_soft.reserve(soft_bytes) != cudaSuccess || h->d_bits.reserve(bits_bytes) != cudaSuccess ||
      h->d_crc_flags.reserve(nof_harq_cb_slots) != cudaSuccess) {
    h->last_error = "HARQ buffer allocation failed";
    cudaGetLastError();
    return fail(SRSRAN_CUDA_ERR_NO_MEMORY);
  }
  // Zero-initialised once, never cleared afterwards (rx_buffer_pool.h:62-63).
  if (cudaMemset(h->d_soft.p, 0, soft_bytes) != cudaSuccess || cudaMemset(h->d_bits.p, 0, bits_bytes) != cudaSuccess ||
      cudaMemset(h->d_crc_flags.p, 0, nof_harq_cb_slots * sizeof(uint32_t)) != cudaSuccess) {
    h->last_error = "HARQ buffer initialisation failed";
    return fail(SRSRAN_CUDA_ERR_CUDA);
  }
  h->extent.assign(nof_harq_cb_slots + 1, 0);
  int prio_low = 0, prio_high = 0;
  cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high);
  for (batch_context& c : h->ctx) {
    bool ok = cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&c.done) == cudaSuccess && cudaEventCreate(&c.kernels) == cudaSuccess &&
              cudaEventCreateWithFlags(&c.copied, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&c.decoded, cudaEventDisableTiming) == cudaSuccess &&
              cudaStreamCreateWithPriority(&c.tail, cudaStreamNonBlocking, prio_high) == cudaSuccess;
    for (cudaEvent_t& e : c.stage) {
      ok = ok && cudaEventCreate(&e) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreate(&c.dm_ev[0]) == cudaSuccess && cudaEventCreate(&c.dm_ev[1]) == cudaSuccess;
    for (int k = 0; k != batch_context::NOF_SIDE; ++k) {
      ok = ok && cudaStreamCreateWithFlags(&c.side[k], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&c.join[k], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) {
      h->last_error = "stream / event creation failed";
      return fail(SRSRAN_CUDA_ERR_CUDA);
    }
  }
  if (cudaEventCreate(&h->timer_begin) != cudaSuccess || cudaEventCreate(&h->timer_end) != cudaSuccess) {
    h->last_error = "event creation failed";
    return fail(SRSRAN_CUDA_ERR_CUDA);
  }
  if (cudaDeviceSynchronize() != cudaSuccess) {
    h->last_error = "device synchronisation failed";
    return fail(SRSRAN_CUDA_ERR_CUDA);
  }
  *handle = h;
  return SRSRAN_CUDA_OK;
}

void srsran_cuda_pusch_dec_destroy(srsran_cuda_pusch_dec_t* h)
{
#ifdef PUSCH_DEC_HOST_PROF
  if (g_prof_n > 8) {
    std::fprintf(stderr, "host profile per submit (us): build %.1f launch_context %.1f [llr copies %.1f grouping %.1f desc copies %.1f kernels+tail %.1f] n=%ld\n",
                 g_prof[0] / (g_prof_n - 8), g_prof[1] / (g_prof_n - 8), g_prof[2] / (g_prof_n - 8), g_prof[3] / (g_prof_n - 8), g_prof[4] / (g_prof_n - 8), g_prof[5] / (g_prof_n - 8), g_prof_n);
  }
#endif
  if (h == nullptr) {
    return;
  }
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  h->d_scr_x1.release();
  h->d_scr_x2.release();
  for (ingest_slot& g : h->ingest) {
    if (g.stream != nullptr) {
      cudaStreamDestroy(g.stream);
    }
    if (g.pushed != nullptr) {
      cudaEventDestroy(g.pushed);
    }
    g.d_llr.release();
  }
  for (batch_context& c : h->ctx) {
    c.h_llr.release();
    c.h_pieces.release();
    c.d_llr.release();
    c.h_desc.release();
    c.d_desc.release();
    c.h_order.release();
    c.d_order.release();
    c.h_grp.release();
    c.d_grp.release();
    c.h_tbmap.release();
    c.d_tbmap.release();
    c.d_tbshare.release();
    c.h_res.release();
    c.d_res.release();
    c.h_bits.release();
    c.d_bits.release();
    c.h_tb.release();
    c.d_tb.release();
    c.h_meta.release();
    c.d_meta.release();
    c.h_outp.release();
    c.d_outp.release();
    c.h_tbres.release();
    c.d_tbdone.release();
    c.d_tbres.release();
    c.h_tbout.release();
    c.d_tbout.release();
    c.d_unit_bits.release();
    c.d_dm_in.release();
    c.d_scr.release();
    c.h_dm.release();
    c.d_dm.release();
    for (cudaEvent_t e : c.dm_ev) {
      if (e != nullptr) {
        cudaEventDestroy(e);
      }
    }
    if (c.done != nullptr) {
      cudaEventDestroy(c.done);
    }
    if (c.kernels != nullptr) {
      cudaEventDestroy(c.kernels);
    }
    if (c.copied != nullptr) {
      cudaEventDestroy(c.copied);
    }
    if (c.decoded != nullptr) {
      cudaEventDestroy(c.decoded);
    }
    if (c.tail != nullptr) {
      cudaStreamDestroy(c.tail);
    }
    for (cudaEvent_t e : c.stage) {
      if (e != nullptr) {
        cudaEventDestroy(e);
      }
    }
    if (c.fork != nullptr) {
      cudaEventDestroy(c.fork);
    }
    for (int k = 0; k != batch_context::NOF_SIDE; ++k) {
      if (c.join[k] != nullptr) {
        cudaEventDestroy(c.join[k]);
      }
      if (c.side[k] != nullptr) {
        cudaStreamDestroy(c.side[k]);
      }
    }
    if (c.stream != nullptr) {
      cudaStreamDestroy(c.stream);
    }
  }
  if (h->timer_begin != nullptr) {
    cudaEventDestroy(h->timer_begin);
  }
  if (h->timer_end != nullptr) {
    cudaEventDestroy(h->timer_end);
  }
  h->d_soft.release();
  h->d_bits.release();
  h->d_crc_flags.release();
  h->d_crc_jobs.release();
  h->d_crc_out.release();
  h->d_crc_msg.release();
  delete h;
}

const char* srsran_cuda_pusch_dec_last_error(const srsran_cuda_pusch_dec_t* h)
{
  return (h == nullptr) ? g_create_error.c_str() : h->last_error.c_str();
}

uint64_t srsran_cuda_pusch_dec_launch_count(const srsran_cuda_pusch_dec_t* h)
{
  return (h == nullptr) ? 0 : h->launches;
}

int srsran_cuda_pusch_dec_set_combine_flavour(srsran_cuda_pusch_dec_t* h, uint32_t simd_block)
{
  if (h == nullptr || (simd_block != 0 && simd_block != 32 && simd_block != 64)) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  h->combine_block = simd_block;
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_set_decoder_variant(srsran_cuda_pusch_dec_t* h, uint32_t variant)
{
  if (h == nullptr || variant > 7) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  h->use_packed  = (variant != 1);
  h->use_tmem    = (variant != 2); // 2: the round-1 packed decoder (messages in shared memory, one CTA per SM) for A/B runs
  h->prefer_q4   = (variant == 3);
  h->force_pairs = (variant == 4);
  h->use_long    = (variant != 5);
  h->prefer_long = (variant == 6);
  h->use_bulk    = (variant == 7);
  return SRSRAN_CUDA_OK;
}

void* srsran_cuda_pusch_dec_host_alloc(size_t bytes)
{
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void srsran_cuda_pusch_dec_host_free(void* p)
{
  if (p != nullptr) {
    cudaFreeHost(p);
  }
}

// ---- HAL-style ------------------------------------------------------------------------------------------------------

int srsran_cuda_pusch_dec_reserve_queue(srsran_cuda_pusch_dec_t* h)
{
  return (h == nullptr) ? SRSRAN_CUDA_ERR_INVALID : SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_free_queue(srsran_cuda_pusch_dec_t* h)
{
  if (h == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  for (hal_op& op : h->hal) {
    op.configured = false;
    op.enqueued   = false;
  }
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_configure(srsran_cuda_pusch_dec_t* h, uint32_t cb_index,
                                    const srsran_cuda_pusch_dec_cb_config* config)
{
  if (h == nullptr || config == nullptr || cb_index >= SRSRAN_CUDA_MAX_NOF_SEGMENTS) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  h->hal[cb_index].cfg        = *config;
  h->hal[cb_index].configured = true;
  h->hal[cb_index].enqueued   = false;
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_enqueue(srsran_cuda_pusch_dec_t* h, uint32_t cb_index, const int8_t* llrs, uint32_t E,
                                  const int8_t* /*softbuf*/, uint32_t /*N*/)
{
  if (h == nullptr || llrs == nullptr || cb_index >= SRSRAN_CUDA_MAX_NOF_SEGMENTS) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  hal_op& op = h->hal[cb_index];
  if (!op.configured || op.cfg.cw_length != E) {
    h->last_error = "enqueue without matching configure";
    return SRSRAN_CUDA_ERR_STATE;
  }
  cudaSetDevice(h->device);
  int ci = open_context(h, 0);
  if (ci < 0) {
    return ci;
  }
  batch_context& c = h->ctx[ci];
  c.hal_style      = true;
  if (c.cb_meta.size() >= h->max_cbs) {
    return 0; // queue full: the caller dequeues first
  }
  int r = ensure_bits_buffers(h, c);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  size_t off;
  r = stage_llrs(h, c, llrs, E, &off);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  const srsran_cuda_pusch_dec_cb_config& cfg = op.cfg;
  cb_params                              p   = {};
  p.bg       = cfg.base_graph;
  p.Z        = cfg.lifting_size;
  p.E        = E;
  p.rv       = cfg.rv;
  p.Qm       = cfg.modulation == 0 ? 1 : cfg.modulation;
  p.Nref     = cfg.Nref;
  p.F        = cfg.nof_filler_bits;
  p.crc_poly = cfg.cb_crc_type == SRSRAN_CUDA_CB_CRC16 ? SRSRAN_CUDA_CRC16
                                                       : (cfg.cb_crc_type == SRSRAN_CUDA_CB_CRC24B ? SRSRAN_CUDA_CRC24B : SRSRAN_CUDA_CRC24A);
  p.max_it   = cfg.max_nof_ldpc_iterations;
  p.mode     = cfg.use_early_stop ? MODE_EARLY_STOP : MODE_CRC_AT_END;
  p.new_data = cfg.new_data;
  p.slot     = cfg.absolute_cb_id;
  p.flags    = FLAG_DEMATCH | FLAG_DECODE | FLAG_USE_HARQ;
  p.scaling  = 0.8F;
  uint32_t idx;
  // The device address is fixed up at launch (the staging buffer may still grow): store the offset for now.
  r = add_cb(h, c, p, reinterpret_cast<const int8_t*>(off), true, &idx);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  op.ctx        = ci;
  op.idx        = idx;
  op.generation = c.generation;
  op.enqueued   = true;
  return 1;
}

static int fixup_and_launch(srsran_cuda_pusch_dec_t* h, int ci)
{
  batch_context& c = h->ctx[ci];
  if (!c.hal_style) {
    // An open context with absolute device pointers can only be the remains of a call that failed half way: it must not be
    // launched (its descriptors are not offsets) - drop it.
    abandon_context(h, c);
    return SRSRAN_CUDA_OK;
  }
  for (size_t i = 0; i != c.cb_meta.size(); ++i) {
    cb_desc& d = c.h_desc.p[i];
    d.llr      = c.d_llr.p + reinterpret_cast<size_t>(d.llr);
    if (d.bits_out != nullptr) {
      d.bits_out = c.d_bits.p + i * BITS_STRIDE;
    }
  }
  return launch_context(h, ci);
}

int srsran_cuda_pusch_dec_dequeue(srsran_cuda_pusch_dec_t* h, uint32_t cb_index, uint8_t* bits, uint32_t bits_size,
                                  int8_t* softbuf_out, uint32_t N)
{
  if (h == nullptr || bits == nullptr || cb_index >= SRSRAN_CUDA_MAX_NOF_SEGMENTS) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  hal_op& op = h->hal[cb_index];
  if (!op.enqueued) {
    h->last_error = "dequeue of an operation that was not enqueued";
    return SRSRAN_CUDA_ERR_STATE;
  }
  cudaSetDevice(h->device);
  batch_context& c = h->ctx[op.ctx];
  if (c.generation != op.generation) {
    h->last_error = "operation result was overwritten";
    return SRSRAN_CUDA_ERR_STATE;
  }
  if (c.open) {
    int r = fixup_and_launch(h, op.ctx);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
  }
  cudaError_t q = cudaEventQuery(c.done);
  if (q == cudaErrorNotReady) {
    return 0;
  }
  if (q != cudaSuccess) {
    h->last_error = std::string("batch failed: ") + cudaGetErrorString(q);
    return SRSRAN_CUDA_ERR_CUDA;
  }
  const cb_host_meta& m      = c.cb_meta[op.idx];
  uint32_t            nbytes = (m.K + 7) / 8;
  if (bits_size < nbytes) {
    h->last_error = "bits buffer too small";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  const uint8_t* src = c.h_bits.p + static_cast<size_t>(op.idx) * BITS_STRIDE;
  std::memcpy(bits, src, m.K / 8);
  if (m.K % 8 != 0) {
    // bit_buffer semantics: only the first K bits belong to the message.
    uint8_t mask    = static_cast<uint8_t>(0xff00U >> (m.K % 8));
    bits[m.K / 8]   = static_cast<uint8_t>((bits[m.K / 8] & ~mask) | (src[m.K / 8] & mask));
  }
  if (softbuf_out != nullptr && N != 0) {
    CUDA_TRY(h, cudaMemcpy(softbuf_out, h->d_soft.p + static_cast<size_t>(m.slot) * SOFT_STRIDE,
                           std::min(N, SOFT_STRIDE), cudaMemcpyDeviceToHost));
  }
  return 1;
}

int srsran_cuda_pusch_dec_read_outputs(srsran_cuda_pusch_dec_t* h, uint32_t cb_index, int* crc_pass,
                                       uint32_t* nof_ldpc_iterations)
{
  if (h == nullptr || cb_index >= SRSRAN_CUDA_MAX_NOF_SEGMENTS) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  hal_op& op = h->hal[cb_index];
  if (!op.enqueued || h->ctx[op.ctx].generation != op.generation || h->ctx[op.ctx].open) {
    h->last_error = "read_outputs before a successful dequeue";
    return SRSRAN_CUDA_ERR_STATE;
  }
  batch_context& c = h->ctx[op.ctx];
  if (cudaEventQuery(c.done) != cudaSuccess) {
    h->last_error = "read_outputs before completion";
    return SRSRAN_CUDA_ERR_STATE;
  }
  const cb_result& r = c.r_res[op.idx];
  if (r.status == 1) {
    h->last_error = "internal: shared-memory layer capacity exceeded";
    return SRSRAN_CUDA_ERR_STATE;
  }
  if (crc_pass != nullptr) {
    *crc_pass = r.crc_ok ? 1 : 0;
  }
  if (nof_ldpc_iterations != nullptr) {
    // Like the software path's statistics (pusch_decoder_impl.cpp:357-363): iterations on success, else the maximum.
    *nof_ldpc_iterations = (r.iters >= 0) ? static_cast<uint32_t>(r.iters) : c.cb_meta[op.idx].max_it;
  }
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_free_harq(srsran_cuda_pusch_dec_t* h, uint32_t absolute_cb_id)
{
  if (h == nullptr || absolute_cb_id >= h->nof_slots) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_is_external_harq_supported(const srsran_cuda_pusch_dec_t*)
{
  return 1;
}

// ---- TB-level -------------------------------------------------------------------------------------------------------

int srsran_cuda_pusch_dec_segment(uint32_t tbs_bits, uint32_t base_graph, uint32_t modulation, uint32_t nof_layers,
                                  uint32_t nof_llrs, srsran_cuda_pusch_dec_cb_meta* out)
{
  if (out == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  return segment(tbs_bits, base_graph, modulation == 0 ? 1 : modulation, nof_layers, nof_llrs, out);
}

namespace {
/// Where the LLRs of one transport block of a batch come from and which HARQ slots it uses.
struct tb_source {
  const int8_t*   llrs;         ///< host memory (copied inside the batch) or device memory
  uint32_t        nof_llrs;
  bool            device;       ///< `llrs` is a device address
  const uint32_t* cb_slots;     ///< absolute code-block ids, or null: harq_first_slot + i
  uint32_t        nof_cb_slots;
};
} // namespace

/// One batch: every transport block becomes one ticket; one set of launches.
static int submit_batch(srsran_cuda_pusch_dec_t* h, uint32_t nof_tbs, const srsran_cuda_pusch_dec_tb_config* configs,
                        const tb_source* src, int* tickets, const std::vector<cudaEvent_t>* waits = nullptr,
                        int* ctx_out = nullptr)
{
  PROF_T(p0);
  cudaSetDevice(h->device);
  if (h->open_ctx >= 0) {
    // A HAL-style batch is still open: launch it first to keep HARQ ordering.
    int r = fixup_and_launch(h, h->open_ctx);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
  }
  uint32_t total_cbs = 0;
  size_t   tb_bytes  = 0;
  for (uint32_t i = 0; i != nof_tbs; ++i) {
    uint32_t B = configs[i].tbs_bits + ((configs[i].tbs_bits <= 3824) ? 16 : 24);
    uint32_t m = (configs[i].base_graph == 1) ? 8448 : 3840;
    total_cbs += (B <= m) ? 1 : (B + m - 25) / (m - 24);
    tb_bytes += configs[i].tbs_bits / 8 + 32;
  }
  int ci = open_context(h, total_cbs);
  if (ci < 0) {
    return ci;
  }
  batch_context& c = h->ctx[ci];
  int            r = prepare_tb_buffers(h, c, nof_tbs, tb_bytes);
  if (r != SRSRAN_CUDA_OK) {
    abandon_context(h, c);
    return r;
  }
  // A small batch (the latency case: one transport block, or a slot of small ones) skips the host -> device copy of its soft
  // bits: the rate dematcher reads a code block's soft bits exactly once, with aligned 16-byte loads, on its way to shared
  // memory - it reads them from the page-locked host memory itself (mapped into the device's address space), so that the
  // transfer and the dematching are one stage instead of two (one config-2 TB: 26 + 14 us -> 28 us). Larger batches stay on
  // the copy engine, which is faster per byte (53.6 vs 49.5 GB/s) and needs no SM.
  bool direct_in = h->direct_in;
  {
    size_t host_bytes = 0;
    for (uint32_t i = 0; direct_in && i != nof_tbs; ++i) {
      if (src[i].device) {
        continue;
      }
      const uint32_t B = configs[i].tbs_bits + ((configs[i].tbs_bits <= 3824) ? 16 : 24);
      const uint32_t m = (configs[i].base_graph == 1) ? 8448 : 3840;
      const uint32_t C = (B <= m) ? 1 : (B + m - 25) / (m - 24);
      host_bytes += src[i].nof_llrs;
      // Every code block must go through the dematcher's shared-memory stage (the other form reads with a stride).
      direct_in = src[i].nof_llrs / C + configs[i].modulation * configs[i].nof_layers <= DM_STAGE_CAP;
    }
    direct_in = direct_in && host_bytes != 0 && host_bytes <= DIRECT_IN_MAX_BYTES;
  }
  // Stage first (the device staging buffer may be reallocated while growing), then build the descriptors.
  std::vector<size_t>        offs(nof_tbs, 0);
  std::vector<const int8_t*> mapped(nof_tbs, nullptr);
  for (uint32_t i = 0; i != nof_tbs; ++i) {
    if (src[i].device) {
      continue;
    }
    const void* dev_src = nullptr;
    if (direct_in && is_pinned(src[i].llrs, &dev_src) && dev_src != nullptr) {
      mapped[i] = static_cast<const int8_t*>(dev_src);
      continue;
    }
    r = stage_llrs(h, c, src[i].llrs, src[i].nof_llrs, &offs[i]);
    if (r != SRSRAN_CUDA_OK) {
      abandon_context(h, c);
      return r;
    }
  }
  if (direct_in) {
    // Pageable sources have been copied into the context's page-locked staging buffer: the dematcher reads that.
    bool all_staged = true;
    for (const batch_context::copy_job& j : c.copies) {
      all_staged = all_staged && j.src == nullptr;
    }
    if (all_staged) {
      for (uint32_t i = 0; i != nof_tbs; ++i) {
        if (!src[i].device && mapped[i] == nullptr) {
          mapped[i] = c.h_llr.p + offs[i];
        }
      }
      c.copies.clear();
    }
  }
  // The packed decoders group CONSECUTIVE code blocks of the same shape (four or two per CTA). A slot of many small
  // transport blocks of mixed shapes (BASELINE config 3) is therefore processed in shape order (stable: a uniform batch keeps
  // its order); the order inside a batch has no meaning, a ticket names its transport block.
  thread_local std::vector<uint32_t> order;
  order.resize(nof_tbs);
  bool uniform = true;
  auto shape_key = [&](uint32_t i) {
    const srsran_cuda_pusch_dec_tb_config& t = configs[i];
    return std::make_tuple(t.base_graph, t.tbs_bits, t.modulation, t.nof_layers, src[i].nof_llrs, t.Nref, t.rv, t.new_data,
                           t.nof_ldpc_iterations, t.use_early_stop);
  };
  for (uint32_t i = 0; i != nof_tbs; ++i) {
    order[i] = i;
    uniform  = uniform && (i == 0 || shape_key(i) == shape_key(i - 1));
  }
  if (!uniform) {
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return shape_key(a) < shape_key(b); });
  }
  for (uint32_t j = 0; j != nof_tbs; ++j) {
    const uint32_t i   = order[j];
    const int8_t*  dev = src[i].device ? src[i].llrs : (mapped[i] != nullptr ? mapped[i] : c.d_llr.p + offs[i]);
    r                  = add_tb(h, c, configs[i], dev, src[i].nof_llrs, src[i].cb_slots, src[i].nof_cb_slots);
    if (r != SRSRAN_CUDA_OK) {
      abandon_context(h, c);
      return r;
    }
    tickets[i] = make_ticket(ci, static_cast<uint32_t>(c.tb_meta.size() - 1), c.generation);
  }
  if (waits != nullptr) {
    c.wait_events = *waits;
  }
  if (ctx_out != nullptr) {
    *ctx_out = ci;
  }
  PROF_T(p1);
  r = launch_context(h, ci);
  PROF_T(p2);
  PROF_ADD(0, p0, p1);
  PROF_ADD(1, p1, p2);
#ifdef PUSCH_DEC_HOST_PROF
  if (++g_prof_n == 8) { // the first uses of every context allocate their buffers
    for (double& v : g_prof) {
      v = 0;
    }
  }
#endif
  return r;
}

static int submit_common(srsran_cuda_pusch_dec_t* h, uint32_t nof_tbs, const srsran_cuda_pusch_dec_tb_config* configs,
                         const int8_t* const* llrs, const uint32_t* nof_llrs, int* tickets, bool device_resident,
                         const uint32_t* cb_slots = nullptr, uint32_t nof_cb_slots = 0, cudaEvent_t wait_for = nullptr,
                         int* ctx_out = nullptr)
{
  if (h == nullptr || configs == nullptr || llrs == nullptr || nof_llrs == nullptr || tickets == nullptr ||
      nof_tbs == 0 || nof_tbs > MAX_TBS_PER_CTX) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  thread_local std::vector<tb_source> src;
  src.resize(nof_tbs);
  for (uint32_t i = 0; i != nof_tbs; ++i) {
    src[i] = {llrs[i], nof_llrs[i], device_resident, cb_slots, nof_cb_slots};
  }
  std::vector<cudaEvent_t> waits;
  if (wait_for != nullptr) {
    waits.push_back(wait_for);
  }
  return submit_batch(h, nof_tbs, configs, src.data(), tickets, wait_for != nullptr ? &waits : nullptr, ctx_out);
}

int srsran_cuda_pusch_dec_submit_tbs_cb_ids(srsran_cuda_pusch_dec_t* h, uint32_t nof_tbs,
                                            const srsran_cuda_pusch_dec_tb_config* configs, const int8_t* const* llrs,
                                            const uint32_t* nof_llrs, const int* ingest_streams,
                                            const uint32_t* absolute_cb_ids, const uint32_t* nof_cb_ids, int* tickets)
{
  nvtx_scope nvtx_range("pusch_dec.submit_tbs");
  if (h == nullptr || configs == nullptr || llrs == nullptr || nof_llrs == nullptr || absolute_cb_ids == nullptr ||
      nof_cb_ids == nullptr || tickets == nullptr || nof_tbs == 0 || nof_tbs > MAX_TBS_PER_CTX) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  cudaSetDevice(h->device);
  std::vector<tb_source>   src(nof_tbs);
  std::vector<cudaEvent_t> waits;
  size_t                   id_off = 0;
  for (uint32_t i = 0; i != nof_tbs; ++i) {
    const int k = (ingest_streams != nullptr) ? ingest_streams[i] : -1;
    if (k >= 0) {
      if (k >= ingest_slot::NOF || !h->ingest[k].open) {
        h->last_error = "ingest stream not open";
        return SRSRAN_CUDA_ERR_STATE;
      }
      src[i] = {h->ingest[k].d_llr.p, h->ingest[k].used, true, absolute_cb_ids + id_off, nof_cb_ids[i]};
    } else {
      if (llrs[i] == nullptr) {
        return SRSRAN_CUDA_ERR_INVALID;
      }
      src[i] = {llrs[i], nof_llrs[i], false, absolute_cb_ids + id_off, nof_cb_ids[i]};
    }
    id_off += nof_cb_ids[i];
  }
  // The streamed transport blocks: their copies are ordered before the batch's kernels by an event each.
  for (uint32_t i = 0; i != nof_tbs; ++i) {
    const int k = (ingest_streams != nullptr) ? ingest_streams[i] : -1;
    if (k >= 0) {
      CUDA_TRY(h, cudaEventRecord(h->ingest[k].pushed, h->ingest[k].stream));
      waits.push_back(h->ingest[k].pushed);
    }
  }
  int ci = -1;
  int r  = submit_batch(h, nof_tbs, configs, src.data(), tickets, waits.empty() ? nullptr : &waits, &ci);
  if (r != SRSRAN_CUDA_OK) {
    return r; // (BUSY: the ingest streams stay open, the caller submits again)
  }
  for (uint32_t i = 0; i != nof_tbs; ++i) {
    const int k = (ingest_streams != nullptr) ? ingest_streams[i] : -1;
    if (k >= 0) {
      h->ingest[k].open       = false;
      h->ingest[k].ctx        = ci;
      h->ingest[k].generation = h->ctx[ci].generation;
    }
  }
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_wait_ticket(srsran_cuda_pusch_dec_t* h, int ticket)
{
  nvtx_scope nvtx_range("pusch_dec.wait_ticket");
  if (h == nullptr || ticket < 0) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  const int ci = (ticket >> 16) & 0xf;
  if (ci >= NOF_CONTEXTS) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  // Touches nothing but the completion event of the ticket's batch (which exists from create to destroy and is not
  // re-recorded before the ticket has been polled): safe beside any other call on the handle.
  cudaSetDevice(h->device);
  return cudaEventSynchronize(h->ctx[ci].done) == cudaSuccess ? SRSRAN_CUDA_OK : SRSRAN_CUDA_ERR_CUDA;
}

int srsran_cuda_pusch_dec_submit_tbs(srsran_cuda_pusch_dec_t* h, uint32_t nof_tbs,
                                     const srsran_cuda_pusch_dec_tb_config* configs, const int8_t* const* llrs,
                                     const uint32_t* nof_llrs, int* tickets)
{
  return submit_common(h, nof_tbs, configs, llrs, nof_llrs, tickets, false);
}

int srsran_cuda_pusch_dec_submit_tbs_device(srsran_cuda_pusch_dec_t* h, uint32_t nof_tbs,
                                            const srsran_cuda_pusch_dec_tb_config* configs,
                                            const int8_t* const* llrs_dev, const uint32_t* nof_llrs, int* tickets)
{
  return submit_common(h, nof_tbs, configs, llrs_dev, nof_llrs, tickets, true);
}

int srsran_cuda_pusch_dec_submit_tb(srsran_cuda_pusch_dec_t* h, const srsran_cuda_pusch_dec_tb_config* config,
                                    const int8_t* llrs, uint32_t nof_llrs)
{
  int ticket = -1;
  int r      = submit_common(h, 1, config, &llrs, &nof_llrs, &ticket, false);
  return (r == SRSRAN_CUDA_OK) ? ticket : r;
}

int srsran_cuda_pusch_dec_submit_tb_cb_ids(srsran_cuda_pusch_dec_t* h, const srsran_cuda_pusch_dec_tb_config* config,
                                           const int8_t* llrs, uint32_t nof_llrs, const uint32_t* absolute_cb_ids,
                                           uint32_t nof_cb_ids)
{
  if (absolute_cb_ids == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  int ticket = -1;
  int r      = submit_common(h, 1, config, &llrs, &nof_llrs, &ticket, false, absolute_cb_ids, nof_cb_ids);
  return (r == SRSRAN_CUDA_OK) ? ticket : r;
}

int srsran_cuda_pusch_dec_stream_begin(srsran_cuda_pusch_dec_t* h, uint32_t max_nof_llrs)
{
  if (h == nullptr || max_nof_llrs == 0) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  cudaSetDevice(h->device);
  for (int i = 0; i != ingest_slot::NOF; ++i) {
    ingest_slot& g = h->ingest[i];
    if (g.open) {
      continue;
    }
    if (g.ctx >= 0) {
      // The device buffer is free again once the batch that read it has completed (or its context was recycled).
      batch_context& c = h->ctx[g.ctx];
      if (c.in_flight && c.generation == g.generation && cudaEventQuery(c.done) == cudaErrorNotReady) {
        continue;
      }
      g.ctx = -1;
    }
    if (g.stream == nullptr) {
      CUDA_TRY(h, cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
      CUDA_TRY(h, cudaEventCreateWithFlags(&g.pushed, cudaEventDisableTiming));
    }
    CUDA_TRY(h, g.d_llr.reserve((static_cast<size_t>(max_nof_llrs) + 15) & ~size_t(15)));
    g.used     = 0;
    g.capacity = max_nof_llrs;
    g.open     = true;
    return i;
  }
  h->last_error = "all ingest streams are busy";
  return SRSRAN_CUDA_ERR_STATE;
}

int srsran_cuda_pusch_dec_stream_push(srsran_cuda_pusch_dec_t* h, int stream, const int8_t* llrs, uint32_t nof_llrs)
{
  if (h == nullptr || stream < 0 || stream >= ingest_slot::NOF || (llrs == nullptr && nof_llrs != 0)) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  ingest_slot& g = h->ingest[stream];
  if (!g.open || g.used + nof_llrs > g.capacity) {
    h->last_error = "ingest stream not open or more LLRs than announced";
    return SRSRAN_CUDA_ERR_STATE;
  }
  if (nof_llrs != 0) {
    cudaSetDevice(h->device);
    CUDA_TRY(h, cudaMemcpyAsync(g.d_llr.p + g.used, llrs, nof_llrs, cudaMemcpyHostToDevice, g.stream));
    g.used += nof_llrs;
  }
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_stream_submit(srsran_cuda_pusch_dec_t* h, int stream, const srsran_cuda_pusch_dec_tb_config* config,
                                        const uint32_t* absolute_cb_ids, uint32_t nof_cb_ids)
{
  nvtx_scope nvtx_range("pusch_dec.stream_submit");
  if (h == nullptr || stream < 0 || stream >= ingest_slot::NOF || config == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  ingest_slot& g = h->ingest[stream];
  if (!g.open) {
    h->last_error = "ingest stream not open";
    return SRSRAN_CUDA_ERR_STATE;
  }
  cudaSetDevice(h->device);
  g.open = false;
  CUDA_TRY(h, cudaEventRecord(g.pushed, g.stream));
  const int8_t* dev    = g.d_llr.p;
  uint32_t      n      = g.used;
  int           ticket = -1, ci = -1;
  int r = submit_common(h, 1, config, &dev, &n, &ticket, true, absolute_cb_ids, absolute_cb_ids ? nof_cb_ids : 0, g.pushed,
                        &ci);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  g.ctx        = ci;
  g.generation = h->ctx[ci].generation;
  return ticket;
}

static int poll_tb_impl(srsran_cuda_pusch_dec_t* h, int ticket, int block, uint8_t* tb,
                        srsran_cuda_pusch_dec_tb_result* result, bool consume);

int srsran_cuda_pusch_dec_poll_tb(srsran_cuda_pusch_dec_t* h, int ticket, int block, uint8_t* tb,
                                  srsran_cuda_pusch_dec_tb_result* result)
{
  nvtx_scope nvtx_range("pusch_dec.poll_tb");
  return poll_tb_impl(h, ticket, block, tb, result, true);
}

int srsran_cuda_pusch_dec_peek_tb(srsran_cuda_pusch_dec_t* h, int ticket, srsran_cuda_pusch_dec_tb_result* result)
{
  return poll_tb_impl(h, ticket, 0, nullptr, result, false);
}

static int poll_tb_impl(srsran_cuda_pusch_dec_t* h, int ticket, int block, uint8_t* tb,
                        srsran_cuda_pusch_dec_tb_result* result, bool consume)
{
  if (h == nullptr || ticket < 0) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  int      ci  = (ticket >> 16) & 0xf;
  uint32_t ti  = static_cast<uint32_t>(ticket) & 0xffff;
  uint32_t gen = (static_cast<uint32_t>(ticket) >> 20) & 0x3ff;
  if (ci >= NOF_CONTEXTS) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  batch_context& c = h->ctx[ci];
  if ((c.generation & 0x3ff) != gen || ti >= c.tb_meta.size() || !c.in_flight) {
    h->last_error = "stale or unknown ticket";
    return SRSRAN_CUDA_ERR_STATE;
  }
  cudaSetDevice(h->device);
  if (block) {
    CUDA_TRY(h, cudaEventSynchronize(c.done));
  } else {
    cudaError_t q = cudaEventQuery(c.done);
    if (q == cudaErrorNotReady) {
      return 0;
    }
    if (q != cudaSuccess) {
      h->last_error = std::string("batch failed: ") + cudaGetErrorString(q);
      return SRSRAN_CUDA_ERR_CUDA;
    }
  }
  tb_host_meta&        m  = c.tb_meta[ti];
  const tb_result_dev& tr = c.r_tbres[ti];
  if (tb != nullptr && tr.written && c.tb_on_host) {
    std::memcpy(tb, c.r_tbout + m.out_offset, m.tbs_bits / 8);
  }
  if (result != nullptr) {
    // sample_statistics<unsigned>::update (include/srsran/support/stats.h:49-53), code blocks in order.
    uint32_t nobs = 0, imin = 0xffffffffU, imax = 0;
    float    mean = 0;
    for (uint32_t i = 0; i != m.nof_cbs; ++i) {
      const cb_result& r = c.r_res[m.first_cb + i];
      if (r.status == 1) {
        h->last_error = "internal: shared-memory layer capacity exceeded";
        return SRSRAN_CUDA_ERR_STATE;
      }
      if (r.status == 2) {
        continue;
      }
      uint32_t obs   = (r.iters >= 0) ? static_cast<uint32_t>(r.iters) : m.max_it;
      float    delta = obs - mean;
      ++nobs;
      mean += delta / nobs;
      imin = std::min(imin, obs);
      imax = std::max(imax, obs);
    }
    result->tb_crc_ok            = tr.tb_crc_ok ? 1 : 0;
    result->nof_codeblocks_total = m.nof_cbs;
    result->nof_observations     = nobs;
    result->iter_min             = nobs ? imin : 0;
    result->iter_max             = nobs ? imax : 0;
    result->iter_mean            = nobs ? mean : 0;
  }
  if (consume) {
    m.polled = true;
  }
  return 1;
}

int srsran_cuda_pusch_dec_poll_tbs(srsran_cuda_pusch_dec_t* h, uint32_t nof_tickets, const int* tickets, int block,
                                   uint8_t* const* tbs, srsran_cuda_pusch_dec_tb_result* results)
{
  nvtx_scope nvtx_range("pusch_dec.poll_tbs");
  if (h == nullptr || tickets == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  // (`results` may be null: the caller has read them with peek_tb and only consumes the tickets.)
  // All or nothing: without `block`, nothing is consumed unless every ticket's batch has completed.
  if (!block) {
    for (uint32_t i = 0; i != nof_tickets; ++i) {
      int ci = (tickets[i] >> 16) & 0xf;
      if (tickets[i] < 0 || ci >= NOF_CONTEXTS) {
        return SRSRAN_CUDA_ERR_INVALID;
      }
      if (h->ctx[ci].in_flight && cudaEventQuery(h->ctx[ci].done) == cudaErrorNotReady) {
        return 0;
      }
    }
  }
  for (uint32_t i = 0; i != nof_tickets; ++i) {
    int r = srsran_cuda_pusch_dec_poll_tb(h, tickets[i], 1, tbs != nullptr ? tbs[i] : nullptr,
                                          results != nullptr ? &results[i] : nullptr);
    if (r < 0) {
      return r;
    }
  }
  return static_cast<int>(nof_tickets);
}

int srsran_cuda_pusch_dec_tb_data(srsran_cuda_pusch_dec_t* h, int ticket, const uint8_t** data)
{
  if (h == nullptr || ticket < 0 || data == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  int      ci  = (ticket >> 16) & 0xf;
  uint32_t ti  = static_cast<uint32_t>(ticket) & 0xffff;
  uint32_t gen = (static_cast<uint32_t>(ticket) >> 20) & 0x3ff;
  if (ci >= NOF_CONTEXTS) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  batch_context& c = h->ctx[ci];
  if ((c.generation & 0x3ff) != gen || ti >= c.tb_meta.size() || !c.in_flight || cudaEventQuery(c.done) != cudaSuccess) {
    h->last_error = "stale, unknown or unfinished ticket";
    return SRSRAN_CUDA_ERR_STATE;
  }
  *data = (c.r_tbres[ti].written && c.tb_on_host) ? c.r_tbout + c.tb_meta[ti].out_offset : nullptr;
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_tb_data_device(srsran_cuda_pusch_dec_t* h, int ticket, const uint8_t** data)
{
  if (h == nullptr || ticket < 0 || data == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  int      ci  = (ticket >> 16) & 0xf;
  uint32_t ti  = static_cast<uint32_t>(ticket) & 0xffff;
  uint32_t gen = (static_cast<uint32_t>(ticket) >> 20) & 0x3ff;
  if (ci >= NOF_CONTEXTS) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  batch_context& c = h->ctx[ci];
  if ((c.generation & 0x3ff) != gen || ti >= c.tb_meta.size() || !c.in_flight || cudaEventQuery(c.done) != cudaSuccess) {
    h->last_error = "stale, unknown or unfinished ticket";
    return SRSRAN_CUDA_ERR_STATE;
  }
  *data = c.r_tbres[ti].written ? c.r_tbout_dev + c.tb_meta[ti].out_offset : nullptr;
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_set_tb_host_copy(srsran_cuda_pusch_dec_t* h, int enable)
{
  if (h == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  h->tb_host_copy = enable != 0;
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_set_direct_io(srsran_cuda_pusch_dec_t* h, int direct_in, int direct_out)
{
  if (h == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  h->direct_in  = direct_in != 0;
  h->direct_out = direct_out != 0;
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_set_h2d_gather(srsran_cuda_pusch_dec_t* h, uint32_t nof_ctas)
{
  if (h == nullptr || nof_ctas > 4096) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  h->h2d_gather_ctas = nof_ctas;
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_tb_cb_outputs(srsran_cuda_pusch_dec_t* h, int ticket, uint8_t* crc_ok, uint32_t* nof_iterations,
                                        uint32_t nof_cbs)
{
  if (h == nullptr || ticket < 0 || crc_ok == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  int      ci  = (ticket >> 16) & 0xf;
  uint32_t ti  = static_cast<uint32_t>(ticket) & 0xffff;
  uint32_t gen = (static_cast<uint32_t>(ticket) >> 20) & 0x3ff;
  if (ci >= NOF_CONTEXTS) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  batch_context& c = h->ctx[ci];
  if ((c.generation & 0x3ff) != gen || ti >= c.tb_meta.size() || !c.in_flight || cudaEventQuery(c.done) != cudaSuccess) {
    h->last_error = "stale, unknown or unfinished ticket";
    return SRSRAN_CUDA_ERR_STATE;
  }
  const tb_host_meta& m = c.tb_meta[ti];
  if (nof_cbs < m.nof_cbs) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  bool all_ok = true;
  for (uint32_t i = 0; i != m.nof_cbs; ++i) {
    const cb_result& r = c.r_res[m.first_cb + i];
    crc_ok[i]          = r.crc_ok ? 1 : 0;
    all_ok             = all_ok && crc_ok[i];
    if (nof_iterations != nullptr) {
      nof_iterations[i] = (r.status == 2) ? 0xffffffffU : ((r.iters >= 0) ? static_cast<uint32_t>(r.iters) : m.max_it);
    }
  }
  // A TB whose code blocks all pass but whose own CRC fails resets every flag (pusch_decoder_impl.cpp:425-428).
  if (all_ok && m.nof_cbs > 1 && !c.r_tbres[ti].tb_crc_ok) {
    std::memset(crc_ok, 0, m.nof_cbs);
  }
  return static_cast<int>(m.nof_cbs);
}

int srsran_cuda_pusch_dec_submit_tbs_symbols(srsran_cuda_pusch_dec_t* h, uint32_t nof_tbs,
                                             const srsran_cuda_pusch_dec_tb_config* configs,
                                             const srsran_cuda_pusch_demod_config* demod_configs,
                                             const float* const* symbols, const float* const* noise_vars, int* tickets,
                                             int device_resident)
{
  nvtx_scope nvtx_range("pusch_dec.submit_tbs_symbols");
  if (h == nullptr || configs == nullptr || demod_configs == nullptr || symbols == nullptr || noise_vars == nullptr ||
      tickets == nullptr || nof_tbs == 0 || nof_tbs > MAX_TBS_PER_CTX) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  cudaSetDevice(h->device);
  int r = init_demod(h);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  if (h->open_ctx >= 0) {
    r = fixup_and_launch(h, h->open_ctx); // a HAL-style batch is still open: launch it first to keep HARQ ordering
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
  }
  uint32_t              total_cbs = 0;
  size_t                tb_bytes = 0, llr_bytes = 0, in_bytes = 0, scr_words = 0, sym_total = 0;
  std::vector<uint32_t> nsym(nof_tbs);
  for (uint32_t i = 0; i != nof_tbs; ++i) {
    nsym[i] = demod_nof_symbols(demod_configs[i]);
    if (nsym[i] == 0 || symbols[i] == nullptr || noise_vars[i] == nullptr ||
        demod_configs[i].modulation != (configs[i].modulation == 0 ? 1 : configs[i].modulation) ||
        demod_configs[i].nof_layers != configs[i].nof_layers) {
      h->last_error = "invalid demodulation configuration";
      return SRSRAN_CUDA_ERR_INVALID;
    }
    uint32_t B = configs[i].tbs_bits + ((configs[i].tbs_bits <= 3824) ? 16 : 24);
    uint32_t m = (configs[i].base_graph == 1) ? 8448 : 3840;
    total_cbs += (B <= m) ? 1 : (B + m - 25) / (m - 24);
    tb_bytes += configs[i].tbs_bits / 8 + 32;
    llr_bytes += (static_cast<size_t>(nsym[i]) * demod_configs[i].modulation + 15) & ~size_t(15);
    in_bytes += (static_cast<size_t>(nsym[i]) * 12 + 15) & ~size_t(15);
    sym_total += static_cast<size_t>(nsym[i]) * 8;
    scr_words += (static_cast<size_t>(nsym[i]) * demod_configs[i].modulation + 31) / 32 + 1;
  }
  int ci = open_context(h, total_cbs);
  if (ci < 0) {
    return ci;
  }
  batch_context& c    = h->ctx[ci];
  auto           fail = [&](int code) {
    abandon_context(h, c);
    return code;
  };
  r = prepare_tb_buffers(h, c, nof_tbs, tb_bytes);
  if (r != SRSRAN_CUDA_OK) {
    return fail(r);
  }
  // All device regions are sized before any address is taken (the buffers may be reallocated while growing).
  if (c.d_llr.reserve(llr_bytes) != cudaSuccess || c.d_scr.reserve(scr_words) != cudaSuccess ||
      c.h_dm.reserve(MAX_TBS_PER_CTX) != cudaSuccess || c.d_dm.reserve(MAX_TBS_PER_CTX) != cudaSuccess ||
      (!device_resident && c.d_dm_in.reserve(in_bytes) != cudaSuccess)) {
    cudaGetLastError();
    h->last_error = "demodulation staging allocation failed";
    return fail(SRSRAN_CUDA_ERR_NO_MEMORY);
  }
  // Staging layout: all symbols back to back, then all noise variances (both 4-byte units, sym_total is a multiple of 8), so
  // that the arrays of a slot that lie back to back in host memory leave in ONE copy each instead of two per transport block.
  size_t llr_off = 0, sym_off = 0, nv_off = sym_total, scr_off = 0;
  auto   push_raw = [&c](const void* src, size_t dst_off, size_t bytes) {
    if (!c.raw_copies.empty() && static_cast<const uint8_t*>(c.raw_copies.back().src) + c.raw_copies.back().bytes == src &&
        c.raw_copies.back().dst_off + c.raw_copies.back().bytes == dst_off) {
      c.raw_copies.back().bytes += bytes;
    } else {
      c.raw_copies.push_back({src, dst_off, bytes});
    }
  };
  for (uint32_t i = 0; i != nof_tbs; ++i) {
    const srsran_cuda_pusch_demod_config& d = demod_configs[i];
    const float*                          sym_dev = symbols[i];
    const float*                          nv_dev  = noise_vars[i];
    if (!device_resident) {
      const size_t sym_bytes = static_cast<size_t>(nsym[i]) * 8, nv_bytes = static_cast<size_t>(nsym[i]) * 4;
      push_raw(symbols[i], sym_off, sym_bytes);
      sym_dev = reinterpret_cast<const float*>(c.d_dm_in.p + sym_off);
      nv_dev  = reinterpret_cast<const float*>(c.d_dm_in.p + nv_off);
      sym_off += sym_bytes;
      nv_off += nv_bytes;
    }
    const uint32_t nllr = nsym[i] * d.modulation;
    add_demod(h, c, d, nsym[i], sym_dev, nv_dev, c.d_llr.p + llr_off, scr_off);
    r = add_tb(h, c, configs[i], c.d_llr.p + llr_off, nllr);
    if (r != SRSRAN_CUDA_OK) {
      return fail(r);
    }
    llr_off += (static_cast<size_t>(nllr) + 15) & ~size_t(15);
    scr_off += (static_cast<size_t>(nllr) + 31) / 32 + 1;
    tickets[i] = make_ticket(ci, static_cast<uint32_t>(c.tb_meta.size() - 1), c.generation);
  }
  if (!device_resident) {
    size_t off = sym_total;
    for (uint32_t i = 0; i != nof_tbs; ++i) {
      push_raw(noise_vars[i], off, static_cast<size_t>(nsym[i]) * 4);
      off += static_cast<size_t>(nsym[i]) * 4;
    }
  }
  c.llr_used = llr_off;
  return launch_context(h, ci);
}

/// One synchronous demodulation of `nof_jobs` codewords with host buffers (unit-level entry points).
static int demod_sync(srsran_cuda_pusch_dec_t* h, int8_t* llrs, const float* symbols, const float* noise_vars,
                      const srsran_cuda_pusch_demod_config& d, uint32_t nsym, uint32_t max_block_subc_override)
{
  cudaSetDevice(h->device);
  int r = init_demod(h);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  r = srsran_cuda_pusch_dec_synchronize(h);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  batch_context& c    = h->ctx[0];
  const size_t   nllr = static_cast<size_t>(nsym) * d.modulation;
  CUDA_TRY(h, c.d_llr.reserve((nllr + 15) & ~size_t(15)));
  CUDA_TRY(h, c.d_scr.reserve(nllr / 32 + 2));
  CUDA_TRY(h, c.d_dm_in.reserve(static_cast<size_t>(nsym) * 12 + 16));
  CUDA_TRY(h, c.h_dm.reserve(MAX_TBS_PER_CTX));
  CUDA_TRY(h, c.d_dm.reserve(MAX_TBS_PER_CTX));
  cudaStream_t s = c.stream;
  CUDA_TRY(h, cudaMemcpyAsync(c.d_dm_in.p, symbols, static_cast<size_t>(nsym) * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(c.d_dm_in.p + static_cast<size_t>(nsym) * 8, noise_vars, static_cast<size_t>(nsym) * 4,
                              cudaMemcpyHostToDevice, s));
  c.ndm = 0;
  c.dm_max_sym = c.dm_max_words = c.dm_qm_mask = 0;
  add_demod(h, c, d, nsym, reinterpret_cast<const float*>(c.d_dm_in.p),
            reinterpret_cast<const float*>(c.d_dm_in.p + static_cast<size_t>(nsym) * 8), c.d_llr.p, 0);
  if (max_block_subc_override != 0) {
    c.h_dm.p[0].max_block_subc = max_block_subc_override;
    c.h_dm.p[0].mbs_magic      = static_cast<uint32_t>((0x100000000ULL + max_block_subc_override - 1) / max_block_subc_override);
  }
  CUDA_TRY(h, cudaMemcpyAsync(c.d_dm.p, c.h_dm.p, sizeof(demod_desc), cudaMemcpyHostToDevice, s));
  scr_seq_kernel<<<dim3((c.dm_max_words + SCR_CHUNK_WORDS - 1) / SCR_CHUNK_WORDS, 1), 32, 0, s>>>(c.d_dm.p, h->d_scr_x1.p,
                                                                                                 h->d_scr_x2.p);
  ++h->launches;
  CUDA_TRY(h, launch_pusch_demod(c.d_dm.p, 1, nsym, c.dm_qm_mask, s, &h->launches));
  CUDA_TRY(h, cudaMemcpyAsync(llrs, c.d_llr.p, nllr, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));
  c.ndm = 0;
  return static_cast<int>(nllr);
}

int srsran_cuda_pusch_demodulate(srsran_cuda_pusch_dec_t* h, int8_t* llrs, const float* symbols, const float* noise_vars,
                                 const srsran_cuda_pusch_demod_config* config)
{
  if (h == nullptr || llrs == nullptr || symbols == nullptr || noise_vars == nullptr || config == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  uint32_t nsym = demod_nof_symbols(*config);
  if (nsym == 0) {
    h->last_error = "invalid demodulation configuration";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  return demod_sync(h, llrs, symbols, noise_vars, *config, nsym, 0);
}

int srsran_cuda_demodulate_soft(srsran_cuda_pusch_dec_t* h, int8_t* llrs, const float* symbols, const float* noise_vars,
                                uint32_t nof_symbols, uint32_t modulation, uint32_t pi2_bpsk)
{
  if (h == nullptr || llrs == nullptr || symbols == nullptr || noise_vars == nullptr || nof_symbols == 0) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  // One demodulation_mapper call = one block of one layer; the scrambling of the codeword-level path is undone below
  // by demodulating with the all-zero sequence (c_init whose x2 cancels x1 does not exist, so the unit-level path
  // re-applies the sequence on the host side of this function instead).
  srsran_cuda_pusch_demod_config d = {};
  d.modulation                     = modulation;
  d.pi2_bpsk                       = pi2_bpsk;
  d.nof_layers                     = 1;
  d.nof_ofdm_symbols               = 1;
  d.re_per_symbol[0]               = nof_symbols;
  uint32_t nsym                    = demod_nof_symbols(d);
  if (nsym == 0) {
    h->last_error = "invalid demodulation configuration";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  int n = demod_sync(h, llrs, symbols, noise_vars, d, nsym, nof_symbols); // one block: the whole input
  if (n < 0) {
    return n;
  }
  // Undo the descrambling (c_init = 0: c(n) = x1(n + 1600)) - the demapper itself does not descramble.
  std::vector<uint32_t> x1((static_cast<size_t>(n) + 31) / 32);
  lfsr_words(1U, 0x9U, x1.data(), static_cast<uint32_t>(x1.size()));
  for (int i = 0; i != n; ++i) {
    if ((x1[i >> 5] >> (i & 31)) & 1U) {
      llrs[i] = static_cast<int8_t>(-llrs[i]);
    }
  }
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_ticket_demod_ms(srsran_cuda_pusch_dec_t* h, int ticket, float* ms)
{
  if (h == nullptr || ticket < 0 || ms == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  int      ci  = (ticket >> 16) & 0xf;
  uint32_t gen = (static_cast<uint32_t>(ticket) >> 20) & 0x3ff;
  if (ci >= NOF_CONTEXTS || (h->ctx[ci].generation & 0x3ff) != gen) {
    h->last_error = "stale or unknown ticket";
    return SRSRAN_CUDA_ERR_STATE;
  }
  batch_context& c = h->ctx[ci];
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaEventSynchronize(c.done));
  *ms = 0.0F;
  if (c.ndm != 0) {
    CUDA_TRY(h, cudaEventElapsedTime(ms, c.dm_ev[0], c.dm_ev[1]));
  }
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_ticket_timing(srsran_cuda_pusch_dec_t* h, int ticket, float* stage_ms)
{
  if (h == nullptr || ticket < 0 || stage_ms == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  int      ci  = (ticket >> 16) & 0xf;
  uint32_t gen = (static_cast<uint32_t>(ticket) >> 20) & 0x3ff;
  if (ci >= NOF_CONTEXTS || (h->ctx[ci].generation & 0x3ff) != gen) {
    h->last_error = "stale or unknown ticket";
    return SRSRAN_CUDA_ERR_STATE;
  }
  batch_context& c = h->ctx[ci];
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaEventSynchronize(c.done));
  cudaEvent_t ev[6] = {c.stage[0], c.stage[1], c.stage[2], c.stage[3], c.kernels, c.done};
  for (int i = 0; i != 5; ++i) {
    CUDA_TRY(h, cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]));
  }
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_last_unit_timing(srsran_cuda_pusch_dec_t* h, float* stage_ms)
{
  if (h == nullptr || stage_ms == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  if (h->last_unit_ctx < 0 || h->ctx[h->last_unit_ctx].in_flight || h->ctx[h->last_unit_ctx].open) {
    h->last_error = "no completed unit-level batch";
    return SRSRAN_CUDA_ERR_STATE;
  }
  batch_context& c = h->ctx[h->last_unit_ctx];
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaEventSynchronize(c.done));
  cudaEvent_t ev[6] = {c.stage[0], c.stage[1], c.stage[2], c.stage[3], c.kernels, c.done};
  for (int i = 0; i != 5; ++i) {
    CUDA_TRY(h, cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]));
  }
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_timer_start(srsran_cuda_pusch_dec_t* h)
{
  if (h == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  h->timer_armed = true;
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_timer_stop(srsran_cuda_pusch_dec_t* h, float* elapsed_ms)
{
  if (h == nullptr || elapsed_ms == nullptr || h->timer_armed || h->last_launched < 0) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  cudaSetDevice(h->device);
  if (h->open_ctx >= 0) {
    int r = fixup_and_launch(h, h->open_ctx);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
  }
  // Everything launched so far: the streams of all contexts in flight join the last launched one.
  cudaStream_t s = h->ctx[h->last_launched].stream;
  for (batch_context& c : h->ctx) {
    if (c.in_flight && c.stream != s) {
      CUDA_TRY(h, cudaStreamWaitEvent(s, c.done, 0));
    }
  }
  CUDA_TRY(h, cudaEventRecord(h->timer_end, s));
  CUDA_TRY(h, cudaEventSynchronize(h->timer_end));
  CUDA_TRY(h, cudaEventElapsedTime(elapsed_ms, h->timer_begin, h->timer_end));
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_synchronize(srsran_cuda_pusch_dec_t* h)
{
  nvtx_scope nvtx_range("pusch_dec.synchronize");
  if (h == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  cudaSetDevice(h->device);
  if (h->open_ctx >= 0) {
    int r = fixup_and_launch(h, h->open_ctx);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
  }
  for (batch_context& c : h->ctx) {
    if (c.in_flight) {
      CUDA_TRY(h, cudaEventSynchronize(c.done));
    }
  }
  return SRSRAN_CUDA_OK;
}

// ---- unit-level -----------------------------------------------------------------------------------------------------

/// Runs one standalone batch (waits for everything before and after).
static int run_unit_batch(srsran_cuda_pusch_dec_t* h, int ci)
{
  int r = launch_context(h, ci);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  CUDA_TRY(h, cudaEventSynchronize(h->ctx[ci].done));
  h->ctx[ci].in_flight = false;
  h->last_unit_ctx     = ci;
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_ldpc_rate_dematch(srsran_cuda_pusch_dec_t* h, int8_t* softbuf, uint32_t N, const int8_t* llrs, uint32_t E,
                                  int new_data, uint32_t rv, uint32_t modulation, uint32_t Nref, uint32_t nof_filler_bits)
{
  if (h == nullptr || softbuf == nullptr || llrs == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  uint32_t bg = (N % 66 == 0) ? 1 : ((N % 50 == 0) ? 2 : 0);
  if (bg == 0) {
    h->last_error = "invalid soft buffer length";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  uint32_t Z = N / ((bg == 1) ? 66 : 50);
  int      r = srsran_cuda_pusch_dec_synchronize(h);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  // The unit-level interface works on a caller-owned buffer: it is staged through a hidden scratch slot.
  uint32_t slot = h->nof_slots;
  CUDA_TRY(h, cudaMemcpy(h->d_soft.p + static_cast<size_t>(slot) * SOFT_STRIDE, softbuf, N, cudaMemcpyHostToDevice));
  int ci = open_context(h, 1);
  if (ci < 0) {
    return ci;
  }
  batch_context& c = h->ctx[ci];
  size_t         off;
  r = stage_llrs(h, c, llrs, E, &off);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  cb_params p = {};
  p.bg        = bg;
  p.Z         = Z;
  p.E         = E;
  p.rv        = rv;
  p.Qm        = modulation == 0 ? 1 : modulation;
  p.Nref      = Nref;
  p.F         = nof_filler_bits;
  p.max_it    = 1;
  p.new_data  = new_data;
  p.slot      = slot;
  p.flags     = FLAG_DEMATCH;
  p.scaling   = 0.8F;
  uint32_t idx;
  r = add_cb(h, c, p, c.d_llr.p + off, false, &idx);
  if (r != SRSRAN_CUDA_OK) {
    abandon_context(h, c);
    return r;
  }
  r = run_unit_batch(h, ci);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  CUDA_TRY(h, cudaMemcpy(softbuf, h->d_soft.p + static_cast<size_t>(slot) * SOFT_STRIDE, N, cudaMemcpyDeviceToHost));
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_ldpc_decode_batch(srsran_cuda_pusch_dec_t* h, uint8_t* bits, const int8_t* llrs, uint32_t nof_cbs,
                                  uint32_t nof_llrs, uint32_t base_graph, uint32_t lifting_size,
                                  uint32_t nof_filler_bits, uint32_t crc_poly, uint32_t max_iterations,
                                  float scaling_factor, int* nof_iterations)
{
  nvtx_scope nvtx_range("pusch_dec.ldpc_decode_batch");
  if (h == nullptr || bits == nullptr || llrs == nullptr || nof_cbs == 0 || (base_graph != 1 && base_graph != 2)) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  if (!(scaling_factor > 0 && scaling_factor <= 1)) {
    h->last_error = "scaling factor out of range";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  int r = srsran_cuda_pusch_dec_synchronize(h);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  uint32_t K      = ((base_graph == 1) ? 22 : 10) * lifting_size;
  uint32_t nbytes = (K + 7) / 8;
  uint32_t done   = 0;
  while (done != nof_cbs) {
    uint32_t n  = std::min(nof_cbs - done, std::max<uint32_t>(h->max_cbs, 1));
    int      ci = open_context(h, n);
    if (ci < 0) {
      return ci;
    }
    batch_context& c = h->ctx[ci];
    r                = ensure_bits_buffers(h, c);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
    // The decoder reads and (conditionally) writes the caller's bit buffer: seed scratch data-bit slots [0, n) with it.
    CUDA_TRY(h, c.d_unit_bits.reserve(static_cast<size_t>(c.h_desc.cap) * BITS_STRIDE + 16));
    c.bits_base = c.d_unit_bits.p;
    std::vector<size_t> offs(n);
    for (uint32_t i = 0; i != n; ++i) {
      r = stage_llrs(h, c, llrs + static_cast<size_t>(done + i) * nof_llrs, nof_llrs, &offs[i]);
      if (r != SRSRAN_CUDA_OK) {
        return r;
      }
    }
    // One strided copy seeds all the scratch data-bit slots (ordered before the kernels on the batch stream).
    CUDA_TRY(h, cudaMemcpy2DAsync(c.d_unit_bits.p, BITS_STRIDE, bits + static_cast<size_t>(done) * nbytes, nbytes, nbytes, n,
                                  cudaMemcpyHostToDevice, c.stream));
    for (uint32_t i = 0; i != n; ++i) {
      cb_params p = {};
      p.bg        = base_graph;
      p.Z         = lifting_size;
      p.F         = nof_filler_bits;
      p.crc_poly  = crc_poly;
      p.max_it    = max_iterations;
      p.mode      = (crc_poly == SRSRAN_CUDA_CRC_NONE) ? MODE_NO_CRC : MODE_EARLY_STOP;
      p.slot      = i;
      p.flags     = FLAG_DECODE;
      p.n_in      = nof_llrs;
      p.scaling   = scaling_factor;
      uint32_t idx;
      r = add_cb(h, c, p, c.d_llr.p + offs[i], true, &idx);
      if (r != SRSRAN_CUDA_OK) {
        abandon_context(h, c);
        return r;
      }
    }
    r = run_unit_batch(h, ci);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
    for (uint32_t i = 0; i != n; ++i) {
      const cb_result& res = c.r_res[i];
      if (res.status == 1) {
        h->last_error = "internal: shared-memory layer capacity exceeded";
        return SRSRAN_CUDA_ERR_STATE;
      }
      uint8_t*       dst = bits + static_cast<size_t>(done + i) * nbytes;
      const uint8_t* src = c.h_bits.p + static_cast<size_t>(i) * BITS_STRIDE;
      std::memcpy(dst, src, K / 8);
      if (K % 8 != 0) {
        uint8_t mask = static_cast<uint8_t>(0xff00U >> (K % 8));
        dst[K / 8]   = static_cast<uint8_t>((dst[K / 8] & ~mask) | (src[K / 8] & mask));
      }
      if (nof_iterations != nullptr) {
        nof_iterations[done + i] = res.iters;
      }
    }
    done += n;
  }
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_ldpc_decode(srsran_cuda_pusch_dec_t* h, uint8_t* bits, const int8_t* llrs, uint32_t nof_llrs,
                            uint32_t base_graph, uint32_t lifting_size, uint32_t nof_filler_bits, uint32_t crc_poly,
                            uint32_t max_iterations, float scaling_factor, int* nof_iterations)
{
  return srsran_cuda_ldpc_decode_batch(h, bits, llrs, 1, nof_llrs, base_graph, lifting_size, nof_filler_bits, crc_poly,
                                       max_iterations, scaling_factor, nof_iterations);
}

int srsran_cuda_crc_calculate(srsran_cuda_pusch_dec_t* h, uint32_t crc_poly, const uint8_t* packed, uint32_t nof_bits,
                              uint32_t* checksum)
{
  if (h == nullptr || packed == nullptr || checksum == nullptr || crc_poly < 1 || crc_poly > 3) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  cudaSetDevice(h->device);
  size_t nbytes = (nof_bits + 7) / 8;
  CUDA_TRY(h, h->d_crc_msg.reserve(nbytes + 16));
  CUDA_TRY(h, h->d_crc_jobs.reserve(1));
  CUDA_TRY(h, h->d_crc_out.reserve(1));
  cudaStream_t s = h->ctx[0].stream;
  crc_job      job{h->d_crc_msg.p, nof_bits, crc_poly};
  CUDA_TRY(h, cudaMemcpyAsync(h->d_crc_msg.p, packed, nbytes, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_crc_jobs.p, &job, sizeof(job), cudaMemcpyHostToDevice, s));
  crc_kernel<<<1, CRC_THREADS, 0, s>>>(h->d_crc_jobs.p, h->d_crc_out.p);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  CUDA_TRY(h, cudaMemcpyAsync(checksum, h->d_crc_out.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));
  return SRSRAN_CUDA_OK;
}

// ---- inspection -----------------------------------------------------------------------------------------------------

int srsran_cuda_pusch_dec_read_softbuffer(srsran_cuda_pusch_dec_t* h, uint32_t absolute_cb_id, int8_t* out, uint32_t N)
{
  if (h == nullptr || out == nullptr || absolute_cb_id >= h->nof_slots || N > SOFT_STRIDE) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  int r = srsran_cuda_pusch_dec_synchronize(h);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  CUDA_TRY(h, cudaMemcpy(out, h->d_soft.p + static_cast<size_t>(absolute_cb_id) * SOFT_STRIDE, N, cudaMemcpyDeviceToHost));
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_write_softbuffer(srsran_cuda_pusch_dec_t* h, uint32_t absolute_cb_id, const int8_t* in,
                                           uint32_t N)
{
  if (h == nullptr || in == nullptr || absolute_cb_id >= h->nof_slots || N > SOFT_STRIDE) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  int r = srsran_cuda_pusch_dec_synchronize(h);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  CUDA_TRY(h, cudaMemcpy(h->d_soft.p + static_cast<size_t>(absolute_cb_id) * SOFT_STRIDE, in, N, cudaMemcpyHostToDevice));
  h->extent[absolute_cb_id] = std::max(h->extent[absolute_cb_id], N);
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pusch_dec_read_cb_crc(srsran_cuda_pusch_dec_t* h, uint32_t absolute_cb_id, int* crc_ok)
{
  if (h == nullptr || crc_ok == nullptr || absolute_cb_id >= h->nof_slots) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  int r = srsran_cuda_pusch_dec_synchronize(h);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  uint32_t v = 0;
  CUDA_TRY(h, cudaMemcpy(&v, h->d_crc_flags.p + absolute_cb_id, sizeof(v), cudaMemcpyDeviceToHost));
  *crc_ok = v ? 1 : 0;
  return SRSRAN_CUDA_OK;
}

} // extern "C"

// ---- PDSCH encoding accelerator: same translation unit, shares the constant tables -----------------
#include "pdsch_enc_api.cuh"
