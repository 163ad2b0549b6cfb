/// \file
/// \brief hal::hw_accelerator_pdsch_enc over the B200 C ABI (include/srsran_cuda_pdsch_enc.h) - the "cuda" sibling of
/// hw_accelerator_pdsch_enc_acc100_impl (lib/hal/phy/upper/channel_processors/hw_accelerator_pdsch_enc_acc100_impl.h).
///
/// Driven unchanged by pdsch_encoder_hw_impl::encode (lib/phy/upper/channel_processors/pdsch_encoder_hw_impl.cpp:34-176):
/// reserve_queue, then per code block (CB mode) or once per transport block (TB mode) configure_operation +
/// enqueue_operation, then dequeue_operation, free_queue. The operations of a transport block are launched together by the
/// first dequeue: one kernel launch per TB instead of one per code block.
#pragma once

#include "srsran/hal/phy/upper/channel_processors/hw_accelerator_pdsch_enc.h"
#include "srsran_cuda_pdsch_enc.h"
#include <memory>
#include <string>

namespace srsran {
namespace hal {

/// Configuration of the CUDA PDSCH encoder accelerator (the counterpart of bbdev_hwacc_pdsch_enc_configuration).
struct cuda_hwacc_pdsch_enc_configuration {
  /// CUDA device ordinal.
  int device = 0;
  /// Operation mode (CB = true [default], TB = false).
  bool cb_mode = true;
  /// Maximum supported TB size in bytes (TB mode; larger TBs fall back to CB mode in pdsch_encoder_hw_impl.cpp:38-42).
  unsigned max_tb_size = 1U << 20;
  /// Operations between reserve_queue() and free_queue() (>= MAX_NOF_SEGMENTS for CB mode).
  unsigned max_nof_operations = 256;
};

class hw_accelerator_pdsch_enc_cuda_impl : public hw_accelerator_pdsch_enc
{
public:
  /// Throws std::runtime_error if no CUDA device is usable (there is no CPU fallback).
  explicit hw_accelerator_pdsch_enc_cuda_impl(const cuda_hwacc_pdsch_enc_configuration& config);
  ~hw_accelerator_pdsch_enc_cuda_impl() override;

  // See hw_accelerator_pdsch_enc for the documentation.
  void     reserve_queue() override {}
  void     free_queue() override {}
  bool     enqueue_operation(span<const uint8_t> data, span<const uint8_t> aux_data = {}, unsigned cb_index = 0) override;
  bool     dequeue_operation(span<uint8_t> data, span<uint8_t> packed_data = {}, unsigned segment_index = 0) override;
  void     configure_operation(const hw_pdsch_encoder_configuration& config, unsigned cb_index = 0) override;
  bool     get_cb_mode() const override { return cfg.cb_mode; }
  unsigned get_max_tb_size() const override { return cfg.max_tb_size; }

private:
  cuda_hwacc_pdsch_enc_configuration cfg;
  srsran_cuda_pdsch_enc_t*           handle = nullptr;
};

} // namespace hal
} // namespace srsran
