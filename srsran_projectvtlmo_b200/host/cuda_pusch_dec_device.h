/// \file
/// \brief Shared ownership of one srsran_cuda_pusch_dec handle (one B200, its HBM-resident HARQ slots and batch contexts).
///
/// Host side of the B200 PUSCH channel-decoding path, to be dropped into lib/hal of srsRAN Project next to the ACC100
/// implementation (lib/hal/phy/upper/channel_processors/pusch/). Everything below this layer is the C ABI of
/// include/srsran_cuda_pusch_dec.h; nothing here computes on the CPU.
#pragma once

#include "srsran_cuda_pusch_dec.h"
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>

namespace srsran {
namespace hal {

/// Parameters of the CUDA accelerator (the counterpart of bbdev_hwacc_pusch_dec_factory_configuration).
struct cuda_hwacc_pusch_dec_configuration {
  /// CUDA device ordinal.
  int device = 0;
  /// Code-block operations that can be queued before the first dequeue launches them as one batch.
  unsigned max_cbs_in_flight = 4096;
  /// Number of HARQ code-block slots kept in HBM: rx_buffer_pool_config::max_nof_codeblocks of a pool created with
  /// external_soft_bits = true (include/srsran/phy/upper/rx_buffer_pool.h:90-103).
  unsigned nof_harq_cb_slots = 4096;
};

/// One CUDA device context shared by every accelerator instance a factory creates: the absolute code-block identifiers
/// of the rx_buffer_pool index the same HBM HARQ slots no matter which pusch_decoder_hw_impl instance decodes.
class cuda_pusch_dec_device
{
public:
  explicit cuda_pusch_dec_device(const cuda_hwacc_pusch_dec_configuration& cfg)
  {
    int st = srsran_cuda_pusch_dec_create(cfg.device, cfg.max_cbs_in_flight, cfg.nof_harq_cb_slots, &handle);
    if (st != SRSRAN_CUDA_OK) {
      // No CPU fallback: the caller gets a null factory (like the ACC100 factory without DPDK).
      throw std::runtime_error(std::string("srsran_cuda_pusch_dec_create failed: ") +
                               srsran_cuda_pusch_dec_last_error(nullptr));
    }
  }
  ~cuda_pusch_dec_device() { srsran_cuda_pusch_dec_destroy(handle); }
  cuda_pusch_dec_device(const cuda_pusch_dec_device&)            = delete;
  cuda_pusch_dec_device& operator=(const cuda_pusch_dec_device&) = delete;

  srsran_cuda_pusch_dec_t* get() { return handle; }
  /// A handle is thread-compatible: accelerator instances sharing it serialise their calls.
  std::mutex& mutex() { return mtx; }

private:
  srsran_cuda_pusch_dec_t* handle = nullptr;
  std::mutex               mtx;
};

} // namespace hal
} // namespace srsran
