/*
 * C ABI of the B200 PDSCH encoding accelerator - the downlink mirror of srsran_cuda_pusch_dec.h:
 * code-block CRC attachment + LDPC encoding + rate matching (bit selection and interleaving) in one kernel launch per
 * batch, behind the reference's hal::hw_accelerator_pdsch_enc seam.
 *
 * Each entry point names the reference interface it replaces; the C++ adapter a maintainer adds on the reference side is
 * srsran_projectvtlmo_b200/host/hw_accelerator_pdsch_enc_cuda_impl.{h,cpp}.
 *
 * Conventions (same as srsran_cuda_pusch_dec.h): plain pointers and sizes, the caller owns every host pointer, the library
 * owns all device memory, int status (SRSRAN_CUDA_OK = 0, negative = error, text via ..._last_error), no exceptions across
 * the boundary, one thread at a time per handle. There is no CPU fallback.
 */
#ifndef SRSRAN_CUDA_PDSCH_ENC_H
#define SRSRAN_CUDA_PDSCH_ENC_H

#include "srsran_cuda_pusch_dec.h" /* status codes, SRSRAN_CUDA_MAX_NOF_SEGMENTS */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct srsran_cuda_pdsch_enc srsran_cuda_pdsch_enc_t;

/* hal::hw_pdsch_encoder_configuration (hw_accelerator_pdsch_enc.h:36-73), field by field. `modulation` = bits per symbol
 * (get_bits_per_symbol of the reference's modulation_scheme), `base_graph` = 1 or 2. */
typedef struct {
  uint32_t nof_tb_bits;
  uint32_t nof_tb_crc_bits;
  uint32_t base_graph;
  uint32_t modulation;
  uint32_t nof_segments;
  uint32_t nof_short_segments;
  uint32_t rv;
  uint32_t cw_length_a;
  uint32_t cw_length_b;
  uint32_t lifting_size;
  uint32_t Ncb;
  uint32_t Nref;
  uint32_t nof_segment_bits;
  uint32_t nof_filler_bits;
  uint32_t rm_length;
  uint8_t  tb_crc[3]; /* the TB checksum bytes as the reference passes them (2 bytes used when nof_tb_crc_bits == 16) */
  uint8_t  cb_mode;   /* 1: one operation per code block (input = its K - F message bits incl. CRCs), 0: per TB */
} srsran_cuda_pdsch_enc_config;

/* pdsch_encoder::configuration (include/srsran/phy/upper/channel_processors/pdsch_encoder.h) for the batch entry point. */
typedef struct {
  uint32_t tbs_bits;       /* multiple of 8 */
  uint32_t base_graph;     /* 1 or 2 */
  uint32_t rv;             /* 0..3 */
  uint32_t modulation;     /* bits per symbol: 2, 4, 6, 8 (1 for pi/2-BPSK) */
  uint32_t Nref;           /* limited-buffer rate matching length, 0 = unlimited */
  uint32_t nof_layers;
  uint32_t nof_ch_symbols; /* the code word has nof_ch_symbols * modulation bits */
} srsran_cuda_pdsch_enc_tb_config;

/* hw_accelerator_pdsch_enc_factory::create (hw_accelerator_pdsch_enc_factory.h): one accelerator object. `max_ops` bounds
 * the operations between reserve_queue() and free_queue() (>= 162 = MAX_NOF_SEGMENTS for CB mode). NULL on failure
 * (srsran_cuda_pdsch_enc_create_error() tells why). */
srsran_cuda_pdsch_enc_t* srsran_cuda_pdsch_enc_create(int device, uint32_t max_ops);
void                     srsran_cuda_pdsch_enc_destroy(srsran_cuda_pdsch_enc_t* handle);
const char*              srsran_cuda_pdsch_enc_last_error(const srsran_cuda_pdsch_enc_t* handle);
const char*              srsran_cuda_pdsch_enc_create_error(void);

/* hw_accelerator_pdsch_enc::configure_operation (hw_accelerator_pdsch_enc.h:86-89). */
int srsran_cuda_pdsch_enc_configure(srsran_cuda_pdsch_enc_t* handle, uint32_t cb_index, const srsran_cuda_pdsch_enc_config* config);
/* hw_accelerator::enqueue_operation (hal/hw_accelerator.h:44-50): `data` = the packed message bits of the code block (CB
 * mode: ceil((K - F) / 8) bytes) or the transport block (TB mode). Returns 1 = enqueued, 0 = queue full (the caller retries
 * after dequeuing, pdsch_encoder_hw_impl.cpp:104-109), < 0 error. Nothing runs yet: the batch is launched by the first
 * dequeue, so that a whole TB (or several) is ONE kernel launch. */
int srsran_cuda_pdsch_enc_enqueue(srsran_cuda_pdsch_enc_t* handle, uint32_t cb_index, const uint8_t* data, uint32_t nof_bytes);
/* hw_accelerator::dequeue_operation (hal/hw_accelerator.h:52-59): `bits` receives the rate-matched bits one per byte
 * (nof_bits of them: rm_length in CB mode, the whole code word in TB mode), `packed` the same bits packed MSB first
 * (ceil(nof_bits / 8) bytes; may be NULL). Returns 1 = dequeued, 0 = nothing enqueued under that index, < 0 error. */
int srsran_cuda_pdsch_enc_dequeue(srsran_cuda_pdsch_enc_t* handle, uint32_t cb_index, uint8_t* bits, uint32_t nof_bits,
                                  uint8_t* packed, uint32_t nof_packed_bytes);

/* pdsch_encoder::encode for a batch of transport blocks in one launch (segmentation metadata on the host, TB CRC, code-block
 * CRCs, encoding and rate matching on the device): tbs[i] = packed TB (tbs_bits / 8 bytes), codewords[i] receives
 * nof_ch_symbols * modulation bits one per byte (may be NULL), packed[i] the same packed MSB first (may be NULL).
 * Synchronous. */
int srsran_cuda_pdsch_enc_encode_tbs(srsran_cuda_pdsch_enc_t* handle, uint32_t nof_tbs, const srsran_cuda_pdsch_enc_tb_config* configs,
                                     const uint8_t* const* tbs, uint8_t* const* codewords, uint8_t* const* packed);
/* The same with the outputs left in device memory (the modulation mapper's input when it runs on the GPU too): returns
 * device pointers valid until the next call on the handle; offsets[i] = first bit of TB i in `*dev_bits` (one bit per
 * byte). */
int srsran_cuda_pdsch_enc_encode_tbs_resident(srsran_cuda_pdsch_enc_t* handle, uint32_t nof_tbs,
                                              const srsran_cuda_pdsch_enc_tb_config* configs, const uint8_t* const* tbs,
                                              const uint8_t** dev_bits, uint64_t* offsets);
/* Device-side durations of the last launched batch in milliseconds: [0] host->device copies, [1] kernels, [2] device->host
 * copies (CUDA events on the handle's stream). */
int srsran_cuda_pdsch_enc_last_timing(srsran_cuda_pdsch_enc_t* handle, float* stage_ms);
/* Number of kernels launched so far. */
uint64_t srsran_cuda_pdsch_enc_launch_count(const srsran_cuda_pdsch_enc_t* handle);

#ifdef __cplusplus
}
#endif
#endif /* SRSRAN_CUDA_PDSCH_ENC_H */
