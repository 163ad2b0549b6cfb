// Host/device shared descriptors of the B200 PUSCH channel-decoding path.
#pragma once
#include <stdint.h>

namespace pusch_dec {

/// HBM layout of the HARQ state: one slot per absolute code-block id (reference: rx_buffer_codeblock_pool,
/// lib/phy/upper/rx_buffer_codeblock_pool.h:38-113 - soft bits + packed data bits + CRC flag per code block).
constexpr uint32_t SOFT_STRIDE = 25344; // 66 * 384 int8 soft bits, 16-byte aligned (25344 = 16 * 1584)
constexpr uint32_t BITS_STRIDE = 1056;  // 8448 / 8 decoded bytes, 16-byte aligned
constexpr uint32_t MAX_Z       = 384;
constexpr uint32_t MAX_EDGES   = 316;
constexpr uint32_t MAX_DEG     = 19;

enum cb_mode : uint8_t {
  MODE_NO_CRC     = 0, ///< ldpc_decoder::decode with crc == nullptr: max_it iterations, returns nullopt.
  MODE_EARLY_STOP = 1, ///< CRC checked after every iteration (ldpc_decoder_impl.cpp:126-134).
  MODE_CRC_AT_END = 2, ///< max_it iterations, then CRC (pusch_codeblock_decoder.cpp:61-70).
};

enum cb_flags : uint8_t {
  FLAG_DEMATCH    = 1, ///< run rate dematching into the HARQ slot
  FLAG_DECODE     = 2, ///< run the LDPC decoder
  FLAG_USE_HARQ   = 4, ///< decoder input is the HARQ slot (else: `llr` is the decoder input itself, unit-level decode)
  FLAG_TRACK_CRC  = 8, ///< honour / update the per-slot CRC flag (TB-level path)
};

/// One code-block operation. 64 bytes.
struct cb_desc {
  const int8_t* llr;      ///< rate-matched LLRs (E of them) in device memory, or the decoder input (unit-level)
  uint8_t*      bits_out; ///< where the K/8 decoded bytes are also copied for the batch D2H (may be null)
  uint32_t      E;
  uint32_t      slot;     ///< HARQ slot
  uint32_t      N;        ///< full code-block length 66Z / 50Z
  uint32_t      Ncb;      ///< circular buffer length: Nref ? min(Nref, N) : N
  uint32_t      k0;       ///< rate-matching start position
  uint32_t      nof_filler;
  uint32_t      n_in;     ///< number of LLRs the decoder reads
  uint32_t      scan_len; ///< upper bound on the non-zero extent of the decoder input (everything beyond is zero)
  uint16_t      Z;
  uint16_t      scale_mult; ///< (uint16)(scaling * 65536), 0 = no scaling
  uint8_t       bg;         ///< 1 or 2
  uint8_t       ils;        ///< lifting-set index 0..7
  uint8_t       Qm;
  uint8_t       new_data;
  uint8_t       crc_poly;
  uint8_t       max_it;
  uint8_t       mode;
  uint8_t       flags;
  uint32_t      layer_cap;  ///< shared-memory capacity, in layers, the launch was sized for
};
static_assert(sizeof(cb_desc) == 64, "cb_desc must stay 64 bytes");

struct cb_result {
  int32_t  iters;      ///< iteration count, -1 = nullopt
  uint32_t crc_ok;     ///< CRC verdict of this operation (also when the code block was skipped because already ok)
  uint32_t nof_layers; ///< layers actually processed (0 if not decoded)
  uint32_t status;     ///< 0 ok, 1 = shared-memory capacity exceeded (host bound wrong), 2 = skipped (CRC already ok)
};

/// One transport-block assembly operation.
struct tb_desc {
  uint32_t first_cb;      ///< index of code block 0 in the batch's cb_desc / cb_result arrays
  uint32_t nof_cbs;
  uint32_t first_slot;    ///< HARQ slot of code block 0 (informative: the kernels take every slot from cb_desc::slot)
  uint32_t tbs_bits;
  uint32_t cb_data_bits;  ///< payload bits per code block (K - crc - filler)
  uint32_t out_offset;    ///< byte offset of this TB in the batch TB output buffer
  uint32_t pad[2];
};

struct tb_result_dev {
  uint32_t tb_crc_ok;
  uint32_t written; ///< 1 if the TB bytes were written (reference writes them only then)
};

} // namespace pusch_dec
