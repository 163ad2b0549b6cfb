/*
 * srsran_cuda_pusch_dec.h - C ABI of the B200 (sm_100a) PUSCH channel-decoding accelerator.
 *
 * Drop-in boundary for srsRAN Project's du_low PUSCH decoding path: LDPC rate dematching with HARQ soft combining,
 * layered normalized min-sum LDPC decoding (BG1/BG2, 51 lifting sizes) and the code-block / transport-block CRC check.
 * Plain pointers and sizes only; the library owns all device memory; the caller owns every pointer it passes. No
 * exceptions cross this boundary: every function returns a status (SRSRAN_CUDA_*) or a documented value.
 * A handle is thread-compatible (one thread at a time), like one hal::hw_accelerator_pusch_dec instance
 * (reference: include/srsran/hal/phy/upper/channel_processors/pusch/hw_accelerator_pusch_dec.h:36-115).
 *
 * There is no CPU fallback: every entry point fails with SRSRAN_CUDA_ERR_NO_DEVICE if no CUDA device is usable.
 *
 * Arithmetic contract: bit-exact to the reference's AVX2/AVX-512 flavour (ldpc_decoder_avx512.cpp,
 * ldpc_rate_dematcher_avx512_impl.cpp, crc_calculator_*): decoded bits, CRC verdict, iteration count and the combined
 * soft-buffer bytes. Valid LLR domain: [-120, 120] and +-127 (include/srsran/phy/upper/log_likelihood_ratio.h:46-51).
 */
#ifndef SRSRAN_CUDA_PUSCH_DEC_H
#define SRSRAN_CUDA_PUSCH_DEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes. */
#define SRSRAN_CUDA_OK 0
#define SRSRAN_CUDA_ERR_NO_DEVICE (-1)   /* no usable CUDA device / driver */
#define SRSRAN_CUDA_ERR_INVALID (-2)     /* invalid argument or configuration */
#define SRSRAN_CUDA_ERR_NO_MEMORY (-3)   /* device or pinned host allocation failed */
#define SRSRAN_CUDA_ERR_CUDA (-4)        /* CUDA runtime error (see srsran_cuda_pusch_dec_last_error) */
#define SRSRAN_CUDA_ERR_STATE (-5)       /* call sequence violated (e.g. dequeue of an operation never enqueued) */
#define SRSRAN_CUDA_ERR_BUSY (-6)        /* every batch context holds transport blocks nobody has polled yet: consume older
                                            tickets (poll_tb / poll_tbs) and submit again - like a full bbdev queue, not fatal */

/* CRC polynomial selectors (crc_generator_poly of include/srsran/phy/upper/channel_coding/crc_calculator.h:31-39). */
#define SRSRAN_CUDA_CRC_NONE 0
#define SRSRAN_CUDA_CRC24A 1
#define SRSRAN_CUDA_CRC24B 2
#define SRSRAN_CUDA_CRC16 3

/* Code-block CRC type as hal::hw_dec_cb_crc_type (hw_accelerator_pusch_dec.h:39). */
#define SRSRAN_CUDA_CB_CRC16 0
#define SRSRAN_CUDA_CB_CRC24B 1
#define SRSRAN_CUDA_CB_CRC24A 2

#define SRSRAN_CUDA_MAX_NOF_SEGMENTS 162 /* MAX_NOF_SEGMENTS, include/srsran/ran/sch/sch_constants.h:33-44 */
#define SRSRAN_CUDA_MAX_CB_LENGTH 25344  /* 66 * 384 */

typedef struct srsran_cuda_pusch_dec srsran_cuda_pusch_dec_t;

/* Field-for-field image of hal::hw_pusch_decoder_configuration (hw_accelerator_pusch_dec.h:42-77). */
typedef struct {
  uint32_t base_graph;              /* 1 = BG1, 2 = BG2 */
  uint32_t modulation;              /* bits per symbol: 1, 2, 4, 6, 8 (0 = pi/2-BPSK, handled as 1) */
  uint32_t nof_segments;            /* code blocks in the TB */
  uint32_t rv;                      /* redundancy version 0..3 */
  uint32_t cw_length;               /* rate-matched length E of this code block */
  uint32_t lifting_size;            /* Z */
  uint32_t Ncb;                     /* 66 Z (BG1) or 50 Z (BG2) */
  uint32_t Nref;                    /* limited-buffer rate matching length, 0 = unlimited */
  uint32_t nof_segment_bits;        /* payload bits of this code block that go into the TB */
  uint32_t nof_filler_bits;
  uint32_t max_nof_ldpc_iterations;
  uint32_t use_early_stop;
  uint32_t new_data;
  uint32_t cb_crc_len;              /* 16 or 24 */
  uint32_t cb_crc_type;             /* SRSRAN_CUDA_CB_CRC16 / CRC24B / CRC24A */
  uint32_t absolute_cb_id;          /* HARQ soft-buffer slot, < nof_harq_cb_slots */
} srsran_cuda_pusch_dec_cb_config;

/* Image of pusch_decoder::configuration (include/srsran/phy/upper/channel_processors/pusch/pusch_decoder.h:57-77) plus
 * what new_data()/on_end_softbits() derive from their arguments. */
typedef struct {
  uint32_t tbs_bits;                /* transport block size in bits (multiple of 8) */
  uint32_t base_graph;              /* 1 or 2 */
  uint32_t rv;
  uint32_t modulation;              /* bits per symbol */
  uint32_t Nref;
  uint32_t nof_layers;
  uint32_t nof_ldpc_iterations;
  uint32_t use_early_stop;
  uint32_t new_data;
  uint32_t harq_first_slot;         /* absolute_cb_id of code block 0; code block i uses harq_first_slot + i */
} srsran_cuda_pusch_dec_tb_config;

/* Image of pusch_decoder_result (pusch_decoder_result.h:30-41) + per-TB bookkeeping. */
typedef struct {
  int32_t  tb_crc_ok;
  uint32_t nof_codeblocks_total;
  uint32_t nof_observations;        /* code blocks decoded in this transmission (those not already CRC-ok) */
  uint32_t iter_min;
  uint32_t iter_max;
  float    iter_mean;
} srsran_cuda_pusch_dec_tb_result;

/* Segmentation metadata of one code block (codeblock_metadata, include/srsran/phy/upper/codeblock_metadata.h). */
typedef struct {
  uint32_t base_graph, lifting_size, full_length, rm_length, nof_filler_bits, cw_offset, nof_crc_bits;
} srsran_cuda_pusch_dec_cb_meta;

/* ---- Lifetime ---------------------------------------------------------------------------------------------------- */

/* Creates an accelerator on CUDA device `device` with room for `max_cbs_in_flight` queued code-block operations and
 * `nof_harq_cb_slots` HBM-resident HARQ code-block slots (25344 soft bits + 1056 data bytes each, zero-initialised and
 * never cleared afterwards, like rx_buffer_pool: include/srsran/phy/upper/rx_buffer_pool.h:62-63).
 * Replaces: hal::create_hw_accelerator_pusch_dec_factory(...)->create() (lib/hal/.../hw_accelerator_factories.cpp:32-84). */
int srsran_cuda_pusch_dec_create(int device, uint32_t max_cbs_in_flight, uint32_t nof_harq_cb_slots,
                                 srsran_cuda_pusch_dec_t** handle);
void srsran_cuda_pusch_dec_destroy(srsran_cuda_pusch_dec_t* handle);
/* Text of the last CUDA/runtime error seen by this handle (or by create when handle is NULL). */
const char* srsran_cuda_pusch_dec_last_error(const srsran_cuda_pusch_dec_t* handle);
/* Number of kernels this handle has launched so far (evidence for bench.py's gpu_launches). */
uint64_t srsran_cuda_pusch_dec_launch_count(const srsran_cuda_pusch_dec_t* handle);

/* Selects which flavour of the reference's HARQ combine is reproduced for NON-FINITE inputs (finite LLRs combine
 * identically in all flavours): 64 = AVX-512 (default; ldpc_rate_dematcher_avx512_impl.cpp:29-64), 32 = AVX2
 * (ldpc_rate_dematcher_avx2_impl.cpp), 0 = generic (ldpc_rate_dematcher_impl.cpp:116-126). */
int srsran_cuda_pusch_dec_set_combine_flavour(srsran_cuda_pusch_dec_t* handle, uint32_t simd_block);
/* Selects the LDPC decoder kernel: 0 (default) = automatic - groups of four same-shape code blocks with few layers
 * (high-rate PUSCH, Z >= 144) run on the packed kernel (four code blocks per CTA in the two binary16 lanes of two
 * registers, one thread per lifted check), single code blocks on the intra-code-block packed kernel (four lifted checks
 * of one code block per thread), the rest on the general kernel; 1 = general kernel only; 2 = packed groups with the
 * messages in shared memory instead of tensor memory (the round-1 kernel, one CTA per SM); 3 = intra-code-block packed
 * kernel wherever it fits; 4 = packed groups of TWO code blocks
 * per CTA, two CTAs per SM (what a small batch uses anyway); 5 = without the many-layer form (pairs of code blocks with up
 * to 46 layers per CTA, messages in tensor memory); 6 = the many-layer form also where the shared-memory pair form of variant 4
 * fits; 7 = the forms that run one CTA per SM stage their inputs with cp.async.bulk + mbarrier instead of
 * 128-bit loads. 2-7 exist for A/B measurements.
 * Results are identical: all variants are bit-exact to the reference (ldpc_decoder_avx512.cpp). */
int srsran_cuda_pusch_dec_set_decoder_variant(srsran_cuda_pusch_dec_t* handle, uint32_t variant);
/* Page-locked host memory for LLR buffers (what pusch_decoder_buffer::get_next_block_view hands to the demodulator,
 * include/srsran/phy/upper/channel_processors/pusch/pusch_decoder_buffer.h:47): LLRs passed from such memory are copied
 * to the device without an intermediate staging copy. */
void* srsran_cuda_pusch_dec_host_alloc(size_t bytes);
void  srsran_cuda_pusch_dec_host_free(void* ptr);

/* ---- hal::hw_accelerator_pusch_dec, one call per method ------------------------------------------------------------ */

/* hw_accelerator_pusch_dec::reserve_queue / free_queue (hw_accelerator_pusch_dec.h:87-90). */
int srsran_cuda_pusch_dec_reserve_queue(srsran_cuda_pusch_dec_t* handle);
int srsran_cuda_pusch_dec_free_queue(srsran_cuda_pusch_dec_t* handle);
/* hw_accelerator_pusch_dec::configure_operation (:95). */
int srsran_cuda_pusch_dec_configure(srsran_cuda_pusch_dec_t* handle, uint32_t cb_index,
                                    const srsran_cuda_pusch_dec_cb_config* config);
/* hw_accelerator<int8_t,uint8_t>::enqueue_operation (include/srsran/hal/hw_accelerator.h:47). `softbuf` is ignored
 * (may be NULL): HARQ soft bits live in HBM (is_external_harq_supported() == true). Returns 1 if enqueued, 0 if the
 * queue is full (caller retries after dequeuing), < 0 on error. */
int srsran_cuda_pusch_dec_enqueue(srsran_cuda_pusch_dec_t* handle, uint32_t cb_index, const int8_t* llrs, uint32_t E,
                                  const int8_t* softbuf, uint32_t N);
/* hw_accelerator::dequeue_operation (:56). The first dequeue after a run of enqueues launches everything queued as one
 * batch. Returns 1 and fills `bits` (K_bg * Z bits, MSB-first, `bits_size` >= ceil(K/8) bytes) when the operation has
 * completed, 0 if it is not ready yet, < 0 on error. If `softbuf_out` is not NULL it receives the N combined soft bits. */
int srsran_cuda_pusch_dec_dequeue(srsran_cuda_pusch_dec_t* handle, uint32_t cb_index, uint8_t* bits, uint32_t bits_size,
                                  int8_t* softbuf_out, uint32_t N);
/* hw_accelerator_pusch_dec::read_operation_outputs (:100-101). Valid after a successful dequeue of cb_index. */
int srsran_cuda_pusch_dec_read_outputs(srsran_cuda_pusch_dec_t* handle, uint32_t cb_index, int* crc_pass,
                                       uint32_t* nof_ldpc_iterations);
/* hw_accelerator_pusch_dec::free_harq_context_entry (:105). The slot content is kept (the reference never clears it). */
int srsran_cuda_pusch_dec_free_harq(srsran_cuda_pusch_dec_t* handle, uint32_t absolute_cb_id);
/* hw_accelerator_pusch_dec::is_external_harq_supported (:109): always 1. */
int srsran_cuda_pusch_dec_is_external_harq_supported(const srsran_cuda_pusch_dec_t* handle);

/* ---- pusch_decoder, one transport block per call ------------------------------------------------------------------- */

/* ldpc_segmenter_rx::segment metadata (lib/phy/upper/channel_coding/ldpc/ldpc_segmenter_impl.cpp:254-331). Writes up to
 * SRSRAN_CUDA_MAX_NOF_SEGMENTS entries; returns the number of code blocks or < 0. Host-only arithmetic. */
int srsran_cuda_pusch_dec_segment(uint32_t tbs_bits, uint32_t base_graph, uint32_t modulation, uint32_t nof_layers,
                                  uint32_t nof_llrs, srsran_cuda_pusch_dec_cb_meta* out);

/* pusch_decoder::new_data + on_new_softbits + on_end_softbits for a whole TB (pusch_decoder_impl.cpp:89-307): copies
 * `nof_llrs` LLRs from HOST memory, runs dematch + LDPC + CB CRC for every code block, assembles the TB and checks its
 * CRC on the device. Asynchronous: returns a ticket >= 0 (or < 0 on error); the CRC flags of the HARQ slots are kept by
 * the handle across transmissions (rx_buffer::get_codeblocks_crc). */
int srsran_cuda_pusch_dec_submit_tb(srsran_cuda_pusch_dec_t* handle, const srsran_cuda_pusch_dec_tb_config* config,
                                    const int8_t* llrs, uint32_t nof_llrs);
/* The same with the HARQ slot of every code block given explicitly: `absolute_cb_ids[i]` is the rx_buffer's absolute
 * code-block id of code block i (rx_buffer::get_absolute_codeblock_id, rx_buffer.h:50-53; the pool hands out ids that
 * are NOT consecutive), `config->harq_first_slot` is ignored. This is the call a `pusch_decoder` implementation makes
 * from on_end_softbits (pusch_decoder_impl.cpp:238-307). */
int srsran_cuda_pusch_dec_submit_tb_cb_ids(srsran_cuda_pusch_dec_t* handle, const srsran_cuda_pusch_dec_tb_config* config,
                                           const int8_t* llrs, uint32_t nof_llrs, const uint32_t* absolute_cb_ids,
                                           uint32_t nof_cb_ids);
/* Streaming ingestion of ONE transport block, for a `pusch_decoder` whose soft bits arrive block by block
 * (pusch_decoder_buffer::on_new_softbits, pusch_decoder_impl.cpp:159-236; pusch_decoder::set_nof_softbits announces the
 * total): `stream_begin` reserves device space for up to `max_nof_llrs` LLRs and returns a stream id >= 0; every
 * `stream_push` appends a block and starts its host -> device copy at once (the host memory must stay valid until the
 * ticket completes; page-locked memory from srsran_cuda_pusch_dec_host_alloc makes the copy asynchronous);
 * `stream_submit` closes the stream and queues dematch + decode + TB assembly behind the copies: at on_end_softbits only
 * the last block is still in flight. `absolute_cb_ids` may be NULL (then config->harq_first_slot + i). Returns a ticket
 * for srsran_cuda_pusch_dec_poll_tb. */
int srsran_cuda_pusch_dec_stream_begin(srsran_cuda_pusch_dec_t* handle, uint32_t max_nof_llrs);
int srsran_cuda_pusch_dec_stream_push(srsran_cuda_pusch_dec_t* handle, int stream, const int8_t* llrs, uint32_t nof_llrs);
int srsran_cuda_pusch_dec_stream_submit(srsran_cuda_pusch_dec_t* handle, int stream,
                                        const srsran_cuda_pusch_dec_tb_config* config, const uint32_t* absolute_cb_ids,
                                        uint32_t nof_cb_ids);
/* Completion of a ticket: returns 1 and fills `tb` (tbs_bits / 8 bytes; written only when the reference writes it) and
 * `result` when done, 0 if `block` is 0 and the TB is still in flight, < 0 on error. Replaces
 * pusch_decoder_notifier::on_sch_data (pusch_decoder_notifier.h:38). */
int srsran_cuda_pusch_dec_poll_tb(srsran_cuda_pusch_dec_t* handle, int ticket, int block, uint8_t* tb,
                                  srsran_cuda_pusch_dec_tb_result* result);

/* The same for `nof_tickets` tickets in one call (a slot's worth of TBs): `tbs` may be NULL or hold NULL entries.
 * Returns nof_tickets when all are complete (always, with `block`), 0 if `block` is 0 and any is still in flight - in
 * which case nothing is consumed - and < 0 on error. */
int srsran_cuda_pusch_dec_poll_tbs(srsran_cuda_pusch_dec_t* handle, uint32_t nof_tickets, const int* tickets, int block,
                                   uint8_t* const* tbs, srsran_cuda_pusch_dec_tb_result* results);

/* Per code block outputs of a completed TB: CRC flags as the reference leaves them in rx_buffer::get_codeblocks_crc()
 * (pusch_decoder_impl.cpp:346-356,425-428: all reset when the TB CRC fails although every code block passed) and, when
 * `nof_iterations` is not NULL, the value the reference feeds to its LDPC statistics for that code block
 * (pusch_decoder_impl.cpp:357-363: the iteration count on success, else the configured maximum; 0xffffffff = no
 * observation: the code block was already ok and only dematched). Returns the number of code blocks. */
int srsran_cuda_pusch_dec_tb_cb_outputs(srsran_cuda_pusch_dec_t* handle, int ticket, uint8_t* crc_ok,
                                        uint32_t* nof_iterations, uint32_t nof_cbs);

/* Zero-copy access to the bytes of a completed TB: `*data` points into the batch's page-locked result buffer (tbs_bits / 8
 * bytes, valid until the ticket's batch context is reused, i.e. for at least the next two submissions); NULL if the
 * reference would not have written the TB (a code-block CRC failed). The ticket must have completed (poll_tb returned 1). */
int srsran_cuda_pusch_dec_tb_data(srsran_cuda_pusch_dec_t* handle, int ticket, const uint8_t** data);
/* Where the decoded transport blocks go. 1 (default): copied to the library's page-locked result buffer (poll_tb,
 * tb_data). 0: left in HBM for a consumer on the device side of the link (tb_data_device: a MAC PDU assembler or a NIC
 * reading GPU memory) - results, statistics and CRC verdicts are still copied; poll_tb then leaves its `tb` argument
 * untouched. At eight GPUs the return path of the decoded bits (16 GB/s per GPU at 130 Gbit/s) is what the box's PCIe
 * fabric limits first. Applies to batches submitted after the call. */
int srsran_cuda_pusch_dec_set_tb_host_copy(srsran_cuda_pusch_dec_t* handle, int enable);
/* Small batches (the latency case: one transport block, a slot of small ones) cross the link without copy-engine jobs.
 * direct_in (default 1): when a batch holds at most 4 MB of host soft bits, the rate dematcher reads them from page-locked
 * host memory itself (the caller's, or the library's staging buffer for pageable memory) instead of a host -> device copy
 * followed by the dematcher. direct_out (default 1): when results + transport-block bytes of a batch are at most 512 KB,
 * the kernels write them straight into the page-locked result buffer instead of a device -> host copy. Results are
 * identical either way; 0 / 0 is the copy-engine path (A/B measurements, tests). */
int srsran_cuda_pusch_dec_set_direct_io(srsran_cuda_pusch_dec_t* handle, int direct_in, int direct_out);
/* Soft bits that arrive in four or more separate page-locked pieces (buffers that are not adjacent in host memory: one per
 * decoder instance behind the plugin interface, pusch_decoder_buffer::get_next_block_view) are read over the link by a
 * gather kernel of `nof_ctas` CTAs instead of one copy-engine job per piece. Default 32; 0 = copy-engine jobs only.
 * Pieces that are adjacent in host memory are merged into one copy before this applies. */
int srsran_cuda_pusch_dec_set_h2d_gather(srsran_cuda_pusch_dec_t* handle, uint32_t nof_ctas);
/* Device address of a completed transport block (valid until its batch context is reused), NULL if the reference would
 * not have written it. */
int srsran_cuda_pusch_dec_tb_data_device(srsran_cuda_pusch_dec_t* handle, int ticket, const uint8_t** data);

/* Same as submit_tb for a batch of TBs whose LLRs are ALREADY RESIDENT in device memory (`llrs_dev[i]` points to
 * `nof_llrs[i]` int8 LLRs in HBM); used when the demodulator runs on the device, and by the benchmark's device-resident
 * leg. One ticket per TB is written to `tickets`. `stream` is a cudaStream_t (NULL = the handle's own stream). */
int srsran_cuda_pusch_dec_submit_tbs_device(srsran_cuda_pusch_dec_t* handle, uint32_t nof_tbs,
                                            const srsran_cuda_pusch_dec_tb_config* configs,
                                            const int8_t* const* llrs_dev, const uint32_t* nof_llrs, int* tickets);
/* Same as submit_tb for a batch of TBs with HOST LLRs: one staging copy, one set of launches. */
int srsran_cuda_pusch_dec_submit_tbs(srsran_cuda_pusch_dec_t* handle, uint32_t nof_tbs,
                                     const srsran_cuda_pusch_dec_tb_config* configs, const int8_t* const* llrs,
                                     const uint32_t* nof_llrs, int* tickets);
/* A batch of TBs whose HARQ slots are the rx buffers' absolute code-block ids (rx_buffer::get_absolute_codeblock_id,
 * include/srsran/phy/upper/rx_buffer.h:50-53; not consecutive): `absolute_cb_ids` holds the ids of all TBs back to back,
 * `nof_cb_ids[i]` of them for TB i. `ingest_streams` (may be NULL): >= 0 for a TB whose LLRs were streamed to the device with
 * stream_begin / stream_push (llrs[i] / nof_llrs[i] are then ignored and the stream is closed by this call), -1 for host
 * LLRs. This is what a slot aggregator above many pusch_decoder instances calls once per batch. SRSRAN_CUDA_ERR_BUSY: nothing
 * was submitted (ingest streams stay open) - poll older tickets and call again. */
int srsran_cuda_pusch_dec_submit_tbs_cb_ids(srsran_cuda_pusch_dec_t* handle, uint32_t nof_tbs,
                                            const srsran_cuda_pusch_dec_tb_config* configs, const int8_t* const* llrs,
                                            const uint32_t* nof_llrs, const int* ingest_streams,
                                            const uint32_t* absolute_cb_ids, const uint32_t* nof_cb_ids, int* tickets);
/* poll_tb without consuming the ticket: returns 1 and the result if the TB has completed, 0 if not. The batch context (and
 * with it the buffer srsran_cuda_pusch_dec_tb_data points into) stays reserved until poll_tb / poll_tbs consume the ticket. */
int srsran_cuda_pusch_dec_peek_tb(srsran_cuda_pusch_dec_t* handle, int ticket, srsran_cuda_pusch_dec_tb_result* result);
/* Blocks until the batch of `ticket` has completed on the device. Unlike every other function it may be called while
 * another thread uses the handle (it only waits on the batch's completion event): a completion thread waits here, then
 * takes the handle for the short poll_tb calls. */
int srsran_cuda_pusch_dec_wait_ticket(srsran_cuda_pusch_dec_t* handle, int ticket);

/* ---- Soft demodulation + descrambling + UL-SCH demultiplexing on the device ------------------------
 * The soft bits are born in HBM: the caller hands over what the channel equalizer produced - equalized symbols (complex
 * binary32, re / im interleaved) and post-equalization noise variances, in the order the reference's demodulator consumes
 * them ([OFDM symbol][RE][layer]) - and the library reproduces, bit for bit (x86 flavour of the reference):
 *   demodulation_mapper::demodulate_soft   lib/phy/upper/channel_modulation/demodulation_mapper_impl.cpp:76-106 (+ _qpsk/_qam16/_qam64/_qam256.cpp)
 *   descrambling with c_init = rnti * 2^15 + n_id  lib/phy/upper/channel_processors/pusch/pusch_demodulator_impl.cpp:38-127,139-140,287-293
 *   the block partition of the demodulator loop (one demapper call per <= MAX_BLOCK_SIZE / bits-per-RE subcarriers of an OFDM
 *   symbol; the SIMD / scalar-tail split of a demapper call depends on it)  pusch_demodulator_impl.cpp:163-279
 *   ulsch_demultiplex without UCI (soft bits pass through in order)  ulsch_demultiplex_impl.cpp:253-275
 * PUSCH with multiplexed UCI is not handled here (SRSRAN_CUDA_ERR_INVALID is not returned for it because the configuration
 * below cannot express it: keep the host demultiplexer for those slots). */
typedef struct {
  uint32_t modulation;        /* bits per symbol: 1 (BPSK), 2, 4, 6, 8 */
  uint32_t pi2_bpsk;          /* 1: pi/2-BPSK (modulation == 1) */
  uint32_t rnti;              /* pusch_demodulator::configuration::rnti */
  uint32_t n_id;              /* pusch_demodulator::configuration::n_id */
  uint32_t nof_layers;        /* nof_tx_layers */
  uint32_t nof_ofdm_symbols;  /* OFDM symbols of the allocation (<= 14): entries of re_per_symbol */
  uint32_t re_per_symbol[14]; /* data REs PER LAYER in each of them (0: a DM-RS symbol without data) */
} srsran_cuda_pusch_demod_config;

/* demodulation_mapper::demodulate_soft on ONE block of `nof_symbols` symbols (unit-level interface, synchronous):
 * `symbols` = nof_symbols x (re, im), host memory; writes nof_symbols * modulation soft bits to `llrs` (host). */
int srsran_cuda_demodulate_soft(srsran_cuda_pusch_dec_t* handle, int8_t* llrs, const float* symbols, const float* noise_vars,
                                uint32_t nof_symbols, uint32_t modulation, uint32_t pi2_bpsk);
/* pusch_demodulator_impl::demodulate minus the equalizer + UL-SCH demultiplexing (no UCI) of one codeword (synchronous,
 * host buffers): writes sum(re_per_symbol) * nof_layers * modulation descrambled soft bits to `llrs`. Returns their number. */
int srsran_cuda_pusch_demodulate(srsran_cuda_pusch_dec_t* handle, int8_t* llrs, const float* symbols, const float* noise_vars,
                                 const srsran_cuda_pusch_demod_config* config);
/* submit_tbs with the soft bits produced on the device: per TB, `symbols[i]` / `noise_vars[i]` hold
 * sum(re_per_symbol) * nof_layers entries, in host memory (`device_resident` = 0: copied to the device inside the call's
 * batch) or already in device memory (1: the upstream equalizer ran on the GPU). One set of launches: scrambling
 * sequences, demodulation, rate dematching, LDPC decoding, TB assembly. One ticket per TB. */
int srsran_cuda_pusch_dec_submit_tbs_symbols(srsran_cuda_pusch_dec_t* handle, uint32_t nof_tbs,
                                             const srsran_cuda_pusch_dec_tb_config* configs,
                                             const srsran_cuda_pusch_demod_config* demod_configs,
                                             const float* const* symbols, const float* const* noise_vars, int* tickets,
                                             int device_resident);
/* Device-side duration in milliseconds of the demodulation stage (scrambling sequences + demodulation kernels) of the batch
 * a ticket belongs to; 0 if the batch had none. Waits for the batch to complete. */
int srsran_cuda_pusch_dec_ticket_demod_ms(srsran_cuda_pusch_dec_t* handle, int ticket, float* ms);

/* Device-side duration (CUDA events on the stream the batch ran on) of the five stages of the batch a ticket belongs
 * to, in milliseconds: [0] host->device copies, [1] rate-dematch kernel, [2] LDPC decode kernels, [3] TB assembly + CRC
 * kernel, [4] device->host copies. Waits for the batch to complete. Used by the benchmark's roofline accounting. */
int srsran_cuda_pusch_dec_ticket_timing(srsran_cuda_pusch_dec_t* handle, int ticket, float* stage_ms);
/* The same five durations for the last batch run through the unit-level interfaces (srsran_cuda_ldpc_decode_batch,
 * srsran_cuda_ldpc_rate_dematch), which are synchronous and hand out no ticket. */
int srsran_cuda_pusch_dec_last_unit_timing(srsran_cuda_pusch_dec_t* handle, float* stage_ms);
/* Device-side stopwatch over several batches: timer_start arms an event that is recorded on the stream of the NEXT batch
 * launched, in front of its first copy; timer_stop records an event behind everything launched so far, waits for it and
 * returns the elapsed milliseconds between the two (CUDA events on the streams the work runs on). */
int srsran_cuda_pusch_dec_timer_start(srsran_cuda_pusch_dec_t* handle);
int srsran_cuda_pusch_dec_timer_stop(srsran_cuda_pusch_dec_t* handle, float* elapsed_ms);
/* Blocks until everything submitted on this handle has completed. */
int srsran_cuda_pusch_dec_synchronize(srsran_cuda_pusch_dec_t* handle);

/* ---- unit-level interfaces (synchronous, host buffers) -------------------------------------------------------------- */

/* ldpc_rate_dematcher::rate_dematch (include/srsran/phy/upper/channel_coding/ldpc/ldpc_rate_dematcher.h:52-55):
 * `softbuf` (N = 66Z / 50Z soft bits) is read and updated in place. */
int srsran_cuda_ldpc_rate_dematch(srsran_cuda_pusch_dec_t* handle, int8_t* softbuf, uint32_t N, const int8_t* llrs,
                                  uint32_t E, int new_data, uint32_t rv, uint32_t modulation, uint32_t Nref,
                                  uint32_t nof_filler_bits);
/* ldpc_decoder::decode (include/srsran/phy/upper/channel_coding/ldpc/ldpc_decoder.h:73-75). `crc_poly` =
 * SRSRAN_CUDA_CRC_NONE mirrors crc == nullptr. `bits` (ceil(K/8) bytes) is read and written like the reference's
 * bit_buffer (left untouched when the reference leaves it untouched). `*nof_iterations` = iteration count, or -1 for
 * std::nullopt. LLRs past the end of a partially filled last node are taken as zero (a fresh reference instance). */
int srsran_cuda_ldpc_decode(srsran_cuda_pusch_dec_t* handle, uint8_t* bits, const int8_t* llrs, uint32_t nof_llrs,
                            uint32_t base_graph, uint32_t lifting_size, uint32_t nof_filler_bits, uint32_t crc_poly,
                            uint32_t max_iterations, float scaling_factor, int* nof_iterations);
/* Batched variant on identical shapes (BASELINE config 1): `nof_cbs` code blocks of `nof_llrs` LLRs each, contiguous. */
int srsran_cuda_ldpc_decode_batch(srsran_cuda_pusch_dec_t* handle, uint8_t* bits, const int8_t* llrs, uint32_t nof_cbs,
                                  uint32_t nof_llrs, uint32_t base_graph, uint32_t lifting_size,
                                  uint32_t nof_filler_bits, uint32_t crc_poly, uint32_t max_iterations,
                                  float scaling_factor, int* nof_iterations);
/* crc_calculator::calculate / calculate_byte (include/srsran/phy/upper/channel_coding/crc_calculator.h:70-80) over the
 * first `nof_bits` bits of an MSB-first packed buffer. */
int srsran_cuda_crc_calculate(srsran_cuda_pusch_dec_t* handle, uint32_t crc_poly, const uint8_t* packed,
                              uint32_t nof_bits, uint32_t* checksum);

/* ---- inspection (tests, HARQ migration) ----------------------------------------------------------------------------- */

/* Copies `N` soft bits of HARQ slot `absolute_cb_id` to / from host memory. */
int srsran_cuda_pusch_dec_read_softbuffer(srsran_cuda_pusch_dec_t* handle, uint32_t absolute_cb_id, int8_t* out,
                                          uint32_t N);
int srsran_cuda_pusch_dec_write_softbuffer(srsran_cuda_pusch_dec_t* handle, uint32_t absolute_cb_id, const int8_t* in,
                                           uint32_t N);
/* CRC flag of a HARQ slot as kept by the TB-level API (rx_buffer::get_codeblocks_crc). */
int srsran_cuda_pusch_dec_read_cb_crc(srsran_cuda_pusch_dec_t* handle, uint32_t absolute_cb_id, int* crc_ok);

#ifdef __cplusplus
}
#endif
#endif /* SRSRAN_CUDA_PUSCH_DEC_H */
