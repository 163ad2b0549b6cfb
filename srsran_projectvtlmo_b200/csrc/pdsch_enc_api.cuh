// C ABI (include/srsran_cuda_pdsch_enc.h) of the PDSCH encoding accelerator: operation table of the hal seam, batch
// staging, one encoder launch per batch. Part of the translation unit pusch_dec_api.cu (shares its base-graph / CRC constant
// tables, buffer helpers and host segmentation arithmetic). No CPU fallback.
#pragma once
#include "../../include/srsran_cuda_pdsch_enc.h"
#include "pdsch_enc.cuh"

namespace {
/// A run of output bits (one per byte, in the batch's bit buffer) that is also wanted packed: the code word of a whole TB,
/// whose code blocks do not end on byte boundaries.
struct enc_out_run {
  uint64_t first_bit, nof_bits, packed_off;
};
} // namespace

struct srsran_cuda_pdsch_enc {
  int          device = 0;
  std::string  last_error;
  cudaStream_t stream = nullptr;
  cudaEvent_t  ev[4]  = {nullptr, nullptr, nullptr, nullptr}; // begin, copies in, kernels, copies out
  uint32_t     max_ops  = 0;
  uint64_t     launches = 0;
  bool         timing_valid = false;
  int          max_smem_optin = 0;

  // hal seam: the operations between reserve_queue() and free_queue().
  enum op_state : uint8_t { EMPTY = 0, CONFIGURED, ENQUEUED, DONE };
  struct op {
    srsran_cuda_pdsch_enc_config cfg      = {};
    op_state                     state    = EMPTY;
    size_t                       in_off   = 0;
    size_t                       bits_off = 0, packed_off = 0; // outputs in the batch buffers
    uint64_t                     nof_bits = 0;
  };
  std::vector<op>       ops;
  std::vector<uint32_t> pending;      // enqueued, not launched yet
  uint32_t              nof_done = 0; // launched and complete, not dequeued yet: their outputs live in h_bits / h_packed

  // The batch being built / last launched. Descriptors hold OFFSETS into the staging / output buffers until the launch,
  // because the buffers may still grow.
  pinned_buf<uint8_t>      h_in;
  device_buf<uint8_t>      d_in;
  size_t                   in_used = 0;
  pinned_buf<enc_cb_desc>  h_desc;
  device_buf<enc_cb_desc>  d_desc;
  uint32_t                 ndesc = 0;
  pinned_buf<crc_job>      h_jobs;
  device_buf<crc_job>      d_jobs;
  device_buf<uint32_t>     d_crcs;
  uint32_t                 njobs = 0;
  std::vector<enc_out_run> runs;
  pinned_buf<uint64_t>     h_runs;
  device_buf<uint64_t>     d_runs;
  size_t                   bits_used = 0, packed_used = 0;
  uint32_t                 max_z[2]  = {0, 0}; // largest lifting size among the code blocks of the byte-per-bit kernel
  uint32_t                 n_packed = 0, n_bytewise = 0; // code blocks per kernel (Z % 32 == 0: packed words)
  device_buf<uint8_t>      d_bits, d_packed;
  pinned_buf<uint8_t>      h_bits, h_packed;
};

namespace {

thread_local std::string g_enc_create_error;

/// Packs runs of bits given one per byte into bytes, MSB first. blockIdx.y = run: (first bit, bits, packed offset).
__global__ void __launch_bounds__(256) pack_bits_kernel(const uint8_t* __restrict__ bits, uint8_t* __restrict__ packed,
                                                        const uint64_t* __restrict__ runs)
{
  const uint64_t first = runs[3 * blockIdx.y], n = runs[3 * blockIdx.y + 1], out0 = runs[3 * blockIdx.y + 2];
  for (uint64_t b = blockIdx.x * 256ULL + threadIdx.x; b * 8 < n; b += (uint64_t)gridDim.x * 256ULL) {
    uint32_t byte = 0;
    if (b * 8 + 8 <= n && ((first + b * 8) & 7ULL) == 0) {
      const uint2 v = *reinterpret_cast<const uint2*>(bits + first + b * 8); // eight bits, one per byte
      // byte k of (v.x, v.y) holds bit k of the group: (b0 b1 b2 b3) -> a nibble by one multiplication (distinct powers,
      // no carries: b0 2^27 + b1 2^26 + b2 2^25 + b3 2^24 in the top byte)
      byte = ((((v.x & 0x01010101U) * 0x08040201U) >> 24) & 0xfU) << 4 | ((((v.y & 0x01010101U) * 0x08040201U) >> 24) & 0xfU);
    } else {
#pragma unroll
      for (uint32_t k = 0; k != 8; ++k) {
        const uint64_t i = b * 8 + k;
        byte |= (i < n ? (uint32_t)(bits[first + i] & 1U) : 0U) << (7 - k);
      }
    }
    packed[out0 + b] = (uint8_t)byte;
  }
}

int enc_upload_core_tables(srsran_cuda_pdsch_enc* h)
{
  enc_core_desc core[2][8];
  for (int bg = 1; bg <= 2; ++bg) {
    const uint16_t* rp  = (bg == 1) ? NR_BG1_ROW_PTR : NR_BG2_ROW_PTR;
    const uint8_t*  col = (bg == 1) ? NR_BG1_COL : NR_BG2_COL;
    const uint32_t  kb  = (bg == 1) ? 22 : 10;
    for (int ils = 0; ils != 8; ++ils) {
      const uint16_t* sh = (bg == 1) ? NR_BG1_SHIFT[ils] : NR_BG2_SHIFT[ils];
      enc_core_desc&  c  = core[bg - 1][ils];
      int             vals[4] = {0, 0, 0, 0}, nvals = 0;
      for (int r = 0; r != 4; ++r) {
        c.ent[r] = -1;
        for (uint32_t e = rp[r]; e != rp[r + 1]; ++e) {
          if (col[e] == kb && nvals < 4) {
            c.ent[r]      = static_cast<int16_t>(sh[e]);
            vals[nvals++] = sh[e];
          }
        }
      }
      // Three entries, two with equal shifts: the odd one out is y (summing the four core rows cancels the equal pair).
      const bool e01 = vals[0] == vals[1], e02 = vals[0] == vals[2], e12 = vals[1] == vals[2];
      if (nvals != 3 || !(e01 || e02 || e12)) {
        h->last_error = "unexpected core parity structure in the base-graph tables";
        return SRSRAN_CUDA_ERR_STATE;
      }
      c.y = static_cast<int16_t>(e01 ? vals[2] : (e02 ? vals[1] : vals[0]));
    }
  }
  CUDA_TRY(h, cudaMemcpyToSymbol(c_enc_core, core, sizeof(core)));
  return SRSRAN_CUDA_OK;
}

template <typename T>
cudaError_t enc_grow_pinned(pinned_buf<T>& b, size_t need, size_t keep)
{
  if (need <= b.cap) {
    return cudaSuccess;
  }
  pinned_buf<T> nb;
  cudaError_t   e = nb.reserve(std::max(need, b.cap * 2 + 1024));
  if (e != cudaSuccess) {
    return e;
  }
  if (b.p != nullptr && keep != 0) {
    std::memcpy(nb.p, b.p, keep * sizeof(T));
  }
  b.release();
  b = nb;
  return cudaSuccess;
}

/// Starts a new batch (descriptors and outputs; the staged inputs are kept: the hal seam stages them at enqueue time). The
/// output regions restart from zero unless results of an earlier launch are still waiting to be dequeued: then the new
/// batch is appended behind them.
void enc_begin(srsran_cuda_pdsch_enc* h)
{
  if (h->nof_done == 0) {
    h->bits_used = h->packed_used = 0;
  }
  h->ndesc = h->njobs = 0;
  h->max_z[0] = h->max_z[1] = 0;
  h->n_packed = h->n_bytewise = 0;
  h->runs.clear();
}

int enc_stage_input(srsran_cuda_pdsch_enc* h, const uint8_t* data, size_t nbytes, size_t* off)
{
  const size_t padded = (nbytes + 15) & ~size_t(15);
  CUDA_TRY(h, enc_grow_pinned(h->h_in, h->in_used + padded, h->in_used));
  std::memcpy(h->h_in.p + h->in_used, data, nbytes);
  *off = h->in_used;
  h->in_used += padded;
  return SRSRAN_CUDA_OK;
}

/// Geometry of one code block: fills the shape fields of a descriptor (input / output addresses are offsets for now).
int enc_fill_desc(srsran_cuda_pdsch_enc* h, enc_cb_desc& d, uint32_t bg, uint32_t Z, uint32_t F, uint32_t E, uint32_t rv,
                  uint32_t Qm, uint32_t Nref)
{
  const int ils = ls_index(Z);
  if ((bg != 1 && bg != 2) || ils < 0 || rv > 3 || Qm == 0 || Qm > 8 || E == 0 || E % Qm != 0) {
    h->last_error = "invalid code-block configuration";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  const uint32_t N = ((bg == 1) ? 66U : 50U) * Z, K = ((bg == 1) ? 22U : 10U) * Z;
  if (F >= K - 2 * Z) {
    h->last_error = "invalid number of filler bits";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  d            = {};
  d.E          = E;
  d.Ncb        = (Nref != 0) ? std::min(Nref, N) : N;
  d.k0         = compute_k0(bg, rv, d.Ncb, N, Z);
  d.nof_filler = F;
  d.Z          = static_cast<uint16_t>(Z);
  d.bg         = static_cast<uint8_t>(bg);
  d.ils        = static_cast<uint8_t>(ils);
  d.Qm         = static_cast<uint8_t>(Qm);
  if (Z % 32 == 0) {
    ++h->n_packed;
  } else {
    ++h->n_bytewise;
    h->max_z[bg - 1] = std::max(h->max_z[bg - 1], Z);
  }
  return SRSRAN_CUDA_OK;
}

int enc_push_desc(srsran_cuda_pdsch_enc* h, const enc_cb_desc& d)
{
  CUDA_TRY(h, enc_grow_pinned(h->h_desc, h->ndesc + 1, h->ndesc));
  h->h_desc.p[h->ndesc++] = d;
  return SRSRAN_CUDA_OK;
}

/// One code block given by its own message bits (CB mode of the hal seam).
int enc_add_cb(srsran_cuda_pdsch_enc* h, const srsran_cuda_pdsch_enc_config& cfg, size_t in_off, size_t* bits_off,
               size_t* packed_off)
{
  enc_cb_desc d;
  int         r = enc_fill_desc(h, d, cfg.base_graph, cfg.lifting_size, cfg.nof_filler_bits, cfg.rm_length, cfg.rv,
                                cfg.modulation == 0 ? 1 : cfg.modulation, cfg.Nref);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  d.info_bits  = ((cfg.base_graph == 1) ? 22U : 10U) * cfg.lifting_size - cfg.nof_filler_bits;
  d.tb_mode    = 0;
  d.src        = reinterpret_cast<const uint8_t*>(in_off);
  *bits_off    = (h->bits_used + 7) & ~size_t(7);
  *packed_off  = h->packed_used;
  d.out_bits   = reinterpret_cast<uint8_t*>(*bits_off);
  d.out_packed = reinterpret_cast<uint8_t*>(*packed_off);
  h->bits_used   = *bits_off + cfg.rm_length;
  h->packed_used = *packed_off + (cfg.rm_length + 7) / 8;
  return enc_push_desc(h, d);
}

/// All code blocks of one transport block whose packed bytes are staged at `in_off`: stream = TB | TB CRC | zero padding,
/// `info` bits of it per code block (ldpc_segmenter_tx_impl.cpp), CRC24B appended when there are several.
/// tb_crc_job: >= 0 = index of the device CRC job that computes the TB checksum, < 0: `tb_crc` is its value.
int enc_add_tb(srsran_cuda_pdsch_enc* h, uint32_t tbs_bits, uint32_t tb_crc_len, uint32_t bg, uint32_t Z, uint32_t F,
               uint32_t C, uint32_t nshort, uint32_t Ea, uint32_t Eb, uint32_t info, uint32_t rv, uint32_t Qm, uint32_t Nref,
               size_t in_off, int tb_crc_job, uint32_t tb_crc, size_t* bits_off, size_t* packed_off, uint64_t* nof_bits)
{
  const uint32_t K = ((bg == 1) ? 22U : 10U) * Z, cb_crc = (C > 1) ? 24U : 0U;
  if (C == 0 || C > SRSRAN_CUDA_MAX_NOF_SEGMENTS || info + cb_crc + F != K || (tb_crc_len != 16 && tb_crc_len != 24) ||
      static_cast<uint64_t>(info) * C < static_cast<uint64_t>(tbs_bits) + tb_crc_len) {
    h->last_error = "inconsistent transport-block segmentation";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  *bits_off     = (h->bits_used + 7) & ~size_t(7);
  *packed_off   = h->packed_used;
  uint64_t nout = 0;
  for (uint32_t r = 0; r != C; ++r) {
    enc_cb_desc d;
    const uint32_t E = (r < nshort) ? Ea : Eb;
    int            st = enc_fill_desc(h, d, bg, Z, F, E, rv, Qm, Nref);
    if (st != SRSRAN_CUDA_OK) {
      return st;
    }
    d.src        = reinterpret_cast<const uint8_t*>(in_off);
    d.out_bits   = reinterpret_cast<uint8_t*>(*bits_off + nout);
    d.out_packed = nullptr; // packed as a whole below
    d.src_bit0   = r * info;
    d.info_bits  = info;
    d.tbs_bits   = tbs_bits;
    d.tb_mode    = (tb_crc_job >= 0) ? 2 : 1;
    d.tb_crc     = (tb_crc_job >= 0) ? static_cast<uint32_t>(tb_crc_job) : tb_crc;
    d.tb_crc_len = static_cast<uint8_t>(tb_crc_len);
    d.cb_crc_len = static_cast<uint8_t>(cb_crc);
    st           = enc_push_desc(h, d);
    if (st != SRSRAN_CUDA_OK) {
      return st;
    }
    nout += E;
  }
  h->runs.push_back({*bits_off, nout, *packed_off});
  h->bits_used   = *bits_off + nout;
  h->packed_used = *packed_off + (nout + 7) / 8;
  *nof_bits      = nout;
  return SRSRAN_CUDA_OK;
}

/// Copies the staged inputs and descriptors to the device and launches the batch (TB CRC jobs, encoder, whole-TB packing).
int enc_launch(srsran_cuda_pdsch_enc* h)
{
  cudaStream_t s = h->stream;
  cudaSetDevice(h->device);
  CUDA_TRY(h, h->d_in.reserve(h->in_used + 16)); // (the kernels read whole words: up to four bytes past a message)
  CUDA_TRY(h, h->d_bits.reserve(h->bits_used + 16));
  CUDA_TRY(h, h->d_packed.reserve(h->packed_used + 16));
  CUDA_TRY(h, h->d_desc.reserve(h->ndesc));
  for (uint32_t i = 0; i != h->ndesc; ++i) {
    enc_cb_desc& d = h->h_desc.p[i];
    d.src          = h->d_in.p + reinterpret_cast<uintptr_t>(d.src);
    d.out_bits     = h->d_bits.p + reinterpret_cast<uintptr_t>(d.out_bits);
    if (d.tb_mode == 0) {
      d.out_packed = h->d_packed.p + reinterpret_cast<uintptr_t>(d.out_packed);
    }
  }
  for (uint32_t i = 0; i != h->njobs; ++i) {
    h->h_jobs.p[i].msg = h->d_in.p + reinterpret_cast<uintptr_t>(h->h_jobs.p[i].msg);
  }
  CUDA_TRY(h, cudaEventRecord(h->ev[0], s));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_in.p, h->h_in.p, h->in_used, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_desc.p, h->h_desc.p, h->ndesc * sizeof(enc_cb_desc), cudaMemcpyHostToDevice, s));
  if (h->njobs != 0) {
    CUDA_TRY(h, h->d_jobs.reserve(h->njobs));
    CUDA_TRY(h, h->d_crcs.reserve(h->njobs));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_jobs.p, h->h_jobs.p, h->njobs * sizeof(crc_job), cudaMemcpyHostToDevice, s));
  }
  if (!h->runs.empty()) {
    CUDA_TRY(h, enc_grow_pinned(h->h_runs, h->runs.size() * 3, 0));
    CUDA_TRY(h, h->d_runs.reserve(h->runs.size() * 3));
    for (size_t i = 0; i != h->runs.size(); ++i) {
      h->h_runs.p[3 * i]     = h->runs[i].first_bit;
      h->h_runs.p[3 * i + 1] = h->runs[i].nof_bits;
      h->h_runs.p[3 * i + 2] = h->runs[i].packed_off;
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->d_runs.p, h->h_runs.p, h->runs.size() * 3 * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  }
  CUDA_TRY(h, cudaEventRecord(h->ev[1], s));
  if (h->njobs != 0) {
    crc_kernel<<<h->njobs, CRC_THREADS, 0, s>>>(h->d_jobs.p, h->d_crcs.p);
    ++h->launches;
    CUDA_TRY(h, cudaGetLastError());
  }
  // Both kernels walk the same descriptor array and skip the code blocks of the other (Z % 32 == 0: packed words).
  if (h->n_packed != 0) {
    pdsch_encode_packed_kernel<<<h->ndesc, ENC_THREADS, 0, s>>>(h->d_desc.p, h->d_crcs.p);
    ++h->launches;
    CUDA_TRY(h, cudaGetLastError());
  }
  if (h->n_bytewise != 0) {
    const uint32_t smem = std::max(h->max_z[0] ? enc_smem_bytes(1, h->max_z[0]) : 0U, h->max_z[1] ? enc_smem_bytes(2, h->max_z[1]) : 0U);
    pdsch_encode_kernel<<<h->ndesc, ENC_THREADS, smem, s>>>(h->d_desc.p, h->d_crcs.p);
    ++h->launches;
    CUDA_TRY(h, cudaGetLastError());
  }
  if (!h->runs.empty()) {
    uint64_t maxbits = 0;
    for (const enc_out_run& r : h->runs) {
      maxbits = std::max(maxbits, r.nof_bits);
    }
    const uint32_t gx = static_cast<uint32_t>(std::min<uint64_t>(maxbits / (8 * 256) + 1, 512));
    pack_bits_kernel<<<dim3(gx, static_cast<uint32_t>(h->runs.size())), 256, 0, s>>>(h->d_bits.p, h->d_packed.p, h->d_runs.p);
    ++h->launches;
    CUDA_TRY(h, cudaGetLastError());
  }
  CUDA_TRY(h, cudaEventRecord(h->ev[2], s));
  return SRSRAN_CUDA_OK;
}

/// Launches the pending operations of the hal seam and brings all their outputs to pinned host memory.
int enc_flush_ops(srsran_cuda_pdsch_enc* h)
{
  enc_begin(h);
  const size_t bits0 = h->bits_used, packed0 = h->packed_used; // results not dequeued yet end here
  for (uint32_t i : h->pending) {
    srsran_cuda_pdsch_enc::op&          o = h->ops[i];
    const srsran_cuda_pdsch_enc_config& c = o.cfg;
    int                                 r;
    if (c.cb_mode) {
      r          = enc_add_cb(h, c, o.in_off, &o.bits_off, &o.packed_off);
      o.nof_bits = c.rm_length;
    } else {
      const uint32_t tb_crc = (c.nof_tb_crc_bits == 16) ? ((static_cast<uint32_t>(c.tb_crc[0]) << 8) | c.tb_crc[1])
                                                        : ((static_cast<uint32_t>(c.tb_crc[0]) << 16) |
                                                           (static_cast<uint32_t>(c.tb_crc[1]) << 8) | c.tb_crc[2]);
      r = enc_add_tb(h, c.nof_tb_bits, c.nof_tb_crc_bits, c.base_graph, c.lifting_size, c.nof_filler_bits, c.nof_segments,
                     c.nof_short_segments, c.cw_length_a, c.cw_length_b, c.nof_segment_bits, c.rv,
                     c.modulation == 0 ? 1 : c.modulation, c.Nref, o.in_off, -1, tb_crc, &o.bits_off, &o.packed_off, &o.nof_bits);
    }
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
  }
  int r = enc_launch(h);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  CUDA_TRY(h, enc_grow_pinned(h->h_bits, h->bits_used + 16, bits0));
  CUDA_TRY(h, enc_grow_pinned(h->h_packed, h->packed_used + 16, packed0));
  CUDA_TRY(h, cudaMemcpyAsync(h->h_bits.p + bits0, h->d_bits.p + bits0, h->bits_used - bits0, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->h_packed.p + packed0, h->d_packed.p + packed0, h->packed_used - packed0, cudaMemcpyDeviceToHost,
                              h->stream));
  CUDA_TRY(h, cudaEventRecord(h->ev[3], h->stream));
  CUDA_TRY(h, cudaEventSynchronize(h->ev[3]));
  h->timing_valid = true;
  for (uint32_t i : h->pending) {
    h->ops[i].state = srsran_cuda_pdsch_enc::DONE;
  }
  h->nof_done += static_cast<uint32_t>(h->pending.size());
  h->pending.clear();
  return SRSRAN_CUDA_OK;
}

/// Segmentation of the batch entry points (the same host arithmetic as the receive side: ldpc_segmenter_impl.cpp).
int enc_add_tb_from_config(srsran_cuda_pdsch_enc* h, const srsran_cuda_pdsch_enc_tb_config& c, const uint8_t* tb, size_t* bits_off,
                           size_t* packed_off, uint64_t* nof_bits)
{
  if (tb == nullptr || c.tbs_bits == 0 || c.tbs_bits % 8 != 0) {
    h->last_error = "invalid transport block";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  const uint32_t                Qm = c.modulation == 0 ? 1 : c.modulation;
  srsran_cuda_pusch_dec_cb_meta metas[SRSRAN_CUDA_MAX_NOF_SEGMENTS];
  const int                     C = segment(c.tbs_bits, c.base_graph, Qm, c.nof_layers, c.nof_ch_symbols * Qm, metas);
  if (C <= 0) {
    h->last_error = "segmentation failed (inconsistent TBS / number of channel symbols)";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  size_t in_off;
  int    r = enc_stage_input(h, tb, c.tbs_bits / 8, &in_off);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  const uint32_t tb_crc_len = (c.tbs_bits <= 3824) ? 16 : 24;
  CUDA_TRY(h, enc_grow_pinned(h->h_jobs, h->njobs + 1, h->njobs));
  h->h_jobs.p[h->njobs] = {reinterpret_cast<const uint8_t*>(in_off), c.tbs_bits,
                           static_cast<uint32_t>(tb_crc_len == 16 ? SRSRAN_CUDA_CRC16 : SRSRAN_CUDA_CRC24A)};
  const int      job    = static_cast<int>(h->njobs++);
  const uint32_t Z      = metas[0].lifting_size, K = ((c.base_graph == 1) ? 22U : 10U) * Z, F = metas[0].nof_filler_bits;
  const uint32_t cb_crc = (C > 1) ? 24U : 0U;
  uint32_t       nshort = 0;
  while (nshort != static_cast<uint32_t>(C) && metas[nshort].rm_length == metas[0].rm_length) {
    ++nshort;
  }
  return enc_add_tb(h, c.tbs_bits, tb_crc_len, c.base_graph, Z, F, static_cast<uint32_t>(C), nshort, metas[0].rm_length,
                    metas[C - 1].rm_length, K - F - cb_crc, c.rv, Qm, c.Nref, in_off, job, 0, bits_off, packed_off, nof_bits);
}

} // namespace

extern "C" {

const char* srsran_cuda_pdsch_enc_create_error(void)
{
  return g_enc_create_error.c_str();
}

srsran_cuda_pdsch_enc_t* srsran_cuda_pdsch_enc_create(int device, uint32_t max_ops)
{
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    g_enc_create_error = "no usable CUDA device (there is no CPU fallback)";
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    g_enc_create_error = "cudaSetDevice failed";
    return nullptr;
  }
  srsran_cuda_pdsch_enc* h = new (std::nothrow) srsran_cuda_pdsch_enc();
  if (h == nullptr) {
    g_enc_create_error = "out of host memory";
    return nullptr;
  }
  h->device  = device;
  h->max_ops = std::max<uint32_t>(max_ops, 1);
  h->ops.resize(h->max_ops);
  cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
  for (cudaEvent_t& e : h->ev) {
    ok = ok && cudaEventCreate(&e) == cudaSuccess;
  }
  ok = ok && upload_tables(h) == SRSRAN_CUDA_OK && enc_upload_core_tables(h) == SRSRAN_CUDA_OK;
  ok = ok && cudaFuncSetAttribute(pdsch_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(enc_smem_bytes(1, MAX_Z))) == cudaSuccess;
  if (!ok) {
    g_enc_create_error = h->last_error.empty() ? "CUDA resource creation failed" : h->last_error;
    srsran_cuda_pdsch_enc_destroy(h);
    return nullptr;
  }
  return h;
}

void srsran_cuda_pdsch_enc_destroy(srsran_cuda_pdsch_enc_t* h)
{
  if (h == nullptr) {
    return;
  }
  cudaSetDevice(h->device);
  if (h->stream != nullptr) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  for (cudaEvent_t e : h->ev) {
    if (e != nullptr) {
      cudaEventDestroy(e);
    }
  }
  h->h_in.release();
  h->d_in.release();
  h->h_desc.release();
  h->d_desc.release();
  h->h_jobs.release();
  h->d_jobs.release();
  h->d_crcs.release();
  h->h_runs.release();
  h->d_runs.release();
  h->d_bits.release();
  h->d_packed.release();
  h->h_bits.release();
  h->h_packed.release();
  delete h;
}

const char* srsran_cuda_pdsch_enc_last_error(const srsran_cuda_pdsch_enc_t* h)
{
  return h == nullptr ? "null handle" : h->last_error.c_str();
}

int srsran_cuda_pdsch_enc_configure(srsran_cuda_pdsch_enc_t* h, uint32_t cb_index, const srsran_cuda_pdsch_enc_config* config)
{
  if (h == nullptr || config == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  if (cb_index >= h->max_ops) {
    h->last_error = "operation index beyond the queue size";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  srsran_cuda_pdsch_enc::op& o = h->ops[cb_index];
  if (o.state == srsran_cuda_pdsch_enc::ENQUEUED) {
    h->last_error = "operation reconfigured while enqueued";
    return SRSRAN_CUDA_ERR_STATE;
  }
  if (o.state == srsran_cuda_pdsch_enc::DONE) {
    --h->nof_done; // its result is dropped
  }
  o.cfg   = *config;
  o.state = srsran_cuda_pdsch_enc::CONFIGURED;
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pdsch_enc_enqueue(srsran_cuda_pdsch_enc_t* h, uint32_t cb_index, const uint8_t* data, uint32_t nof_bytes)
{
  if (h == nullptr || data == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  if (cb_index >= h->max_ops) {
    return 0; // queue full: the caller dequeues and retries (pdsch_encoder_hw_impl.cpp:104-109)
  }
  srsran_cuda_pdsch_enc::op& o = h->ops[cb_index];
  if (o.state != srsran_cuda_pdsch_enc::CONFIGURED) {
    h->last_error = "enqueue of an operation that is not configured (or already enqueued)";
    return SRSRAN_CUDA_ERR_STATE;
  }
  const srsran_cuda_pdsch_enc_config& c = o.cfg;
  const uint32_t need = c.cb_mode ? (((c.base_graph == 1 ? 22U : 10U) * c.lifting_size - c.nof_filler_bits + 7) / 8) : c.nof_tb_bits / 8;
  if (nof_bytes < need) {
    h->last_error = "input shorter than the configured message";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  if (h->pending.empty()) {
    h->in_used = 0; // first operation of a new batch: the previous batch's inputs are no longer needed
  }
  int r = enc_stage_input(h, data, need, &o.in_off);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  o.state = srsran_cuda_pdsch_enc::ENQUEUED;
  h->pending.push_back(cb_index);
  return 1;
}

int srsran_cuda_pdsch_enc_dequeue(srsran_cuda_pdsch_enc_t* h, uint32_t cb_index, uint8_t* bits, uint32_t nof_bits, uint8_t* packed,
                                  uint32_t nof_packed_bytes)
{
  if (h == nullptr || bits == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  if (cb_index >= h->max_ops) {
    return 0;
  }
  srsran_cuda_pdsch_enc::op& o = h->ops[cb_index];
  if (o.state == srsran_cuda_pdsch_enc::ENQUEUED) {
    nvtx_scope nvtx_range("pdsch_enc.launch_batch");
    int        r = enc_flush_ops(h);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
  }
  if (o.state != srsran_cuda_pdsch_enc::DONE) {
    return 0;
  }
  if (nof_bits != o.nof_bits || (packed != nullptr && nof_packed_bytes < (o.nof_bits + 7) / 8)) {
    h->last_error = "output spans do not match the configured rate-matched length";
    return SRSRAN_CUDA_ERR_INVALID;
  }
  std::memcpy(bits, h->h_bits.p + o.bits_off, o.nof_bits);
  if (packed != nullptr) {
    std::memcpy(packed, h->h_packed.p + o.packed_off, (o.nof_bits + 7) / 8);
  }
  o.state = srsran_cuda_pdsch_enc::EMPTY;
  --h->nof_done;
  return 1;
}

static int enc_batch_common(srsran_cuda_pdsch_enc_t* h, uint32_t nof_tbs, const srsran_cuda_pdsch_enc_tb_config* configs,
                            const uint8_t* const* tbs, std::vector<size_t>& boff, std::vector<size_t>& poff,
                            std::vector<uint64_t>& nbits)
{
  if (h == nullptr || configs == nullptr || tbs == nullptr || nof_tbs == 0) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  if (!h->pending.empty() || h->nof_done != 0) {
    h->last_error = "operations of the hal seam are pending or not dequeued yet";
    return SRSRAN_CUDA_ERR_STATE;
  }
  enc_begin(h);
  h->in_used = 0;
  boff.resize(nof_tbs);
  poff.resize(nof_tbs);
  nbits.resize(nof_tbs);
  for (uint32_t i = 0; i != nof_tbs; ++i) {
    int r = enc_add_tb_from_config(h, configs[i], tbs[i], &boff[i], &poff[i], &nbits[i]);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
  }
  return enc_launch(h);
}

int srsran_cuda_pdsch_enc_encode_tbs(srsran_cuda_pdsch_enc_t* h, uint32_t nof_tbs, const srsran_cuda_pdsch_enc_tb_config* configs,
                                     const uint8_t* const* tbs, uint8_t* const* codewords, uint8_t* const* packed)
{
  nvtx_scope            nvtx_range("pdsch_enc.encode_tbs");
  std::vector<size_t>   boff, poff;
  std::vector<uint64_t> nbits;
  int                   r = enc_batch_common(h, nof_tbs, configs, tbs, boff, poff, nbits);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  // Outputs that are adjacent on both sides (the caller laid the code words of a slot back to back) leave in one copy.
  auto copy_out = [&](uint8_t* const* dst, const uint8_t* dev, const std::vector<size_t>& off, bool pk) -> int {
    uint32_t i = 0;
    while (i != nof_tbs) {
      if (dst[i] == nullptr) {
        ++i;
        continue;
      }
      auto   len = [&](uint32_t k) { return pk ? static_cast<size_t>((nbits[k] + 7) / 8) : static_cast<size_t>(nbits[k]); };
      size_t n   = len(i);
      uint32_t j = i + 1;
      while (j != nof_tbs && dst[j] == dst[i] + n && off[j] == off[i] + n) {
        n += len(j);
        ++j;
      }
      CUDA_TRY(h, cudaMemcpyAsync(dst[i], dev + off[i], n, cudaMemcpyDeviceToHost, h->stream));
      i = j;
    }
    return SRSRAN_CUDA_OK;
  };
  if (codewords != nullptr) {
    r = copy_out(codewords, h->d_bits.p, boff, false);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
  }
  if (packed != nullptr) {
    r = copy_out(packed, h->d_packed.p, poff, true);
    if (r != SRSRAN_CUDA_OK) {
      return r;
    }
  }
  CUDA_TRY(h, cudaEventRecord(h->ev[3], h->stream));
  CUDA_TRY(h, cudaEventSynchronize(h->ev[3]));
  h->timing_valid = true;
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pdsch_enc_encode_tbs_resident(srsran_cuda_pdsch_enc_t* h, uint32_t nof_tbs,
                                              const srsran_cuda_pdsch_enc_tb_config* configs, const uint8_t* const* tbs,
                                              const uint8_t** dev_bits, uint64_t* offsets)
{
  nvtx_scope            nvtx_range("pdsch_enc.encode_tbs_resident");
  std::vector<size_t>   boff, poff;
  std::vector<uint64_t> nbits;
  if (dev_bits == nullptr || offsets == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  int r = enc_batch_common(h, nof_tbs, configs, tbs, boff, poff, nbits);
  if (r != SRSRAN_CUDA_OK) {
    return r;
  }
  CUDA_TRY(h, cudaEventRecord(h->ev[3], h->stream));
  CUDA_TRY(h, cudaEventSynchronize(h->ev[3]));
  h->timing_valid = true;
  *dev_bits       = h->d_bits.p;
  for (uint32_t i = 0; i != nof_tbs; ++i) {
    offsets[i] = boff[i];
  }
  return SRSRAN_CUDA_OK;
}

int srsran_cuda_pdsch_enc_last_timing(srsran_cuda_pdsch_enc_t* h, float* stage_ms)
{
  if (h == nullptr || stage_ms == nullptr) {
    return SRSRAN_CUDA_ERR_INVALID;
  }
  if (!h->timing_valid) {
    h->last_error = "no completed batch";
    return SRSRAN_CUDA_ERR_STATE;
  }
  for (int i = 0; i != 3; ++i) {
    CUDA_TRY(h, cudaEventElapsedTime(&stage_ms[i], h->ev[i], h->ev[i + 1]));
  }
  return SRSRAN_CUDA_OK;
}

uint64_t srsran_cuda_pdsch_enc_launch_count(const srsran_cuda_pdsch_enc_t* h)
{
  return h == nullptr ? 0 : h->launches;
}

} // extern "C"
