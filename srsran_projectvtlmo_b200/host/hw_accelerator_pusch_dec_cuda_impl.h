/// \file
/// \brief hal::hw_accelerator_pusch_dec over the B200 C ABI - the "cuda" sibling of hw_accelerator_pusch_dec_acc100_impl
/// (lib/hal/phy/upper/channel_processors/pusch/hw_accelerator_pusch_dec_acc100_impl.h).
///
/// Driven unchanged by pusch_decoder_hw_impl::on_end_softbits (lib/phy/upper/channel_processors/pusch/
/// pusch_decoder_hw_impl.cpp:132-342): reserve_queue, then per code block configure_operation + enqueue_operation, then
/// dequeue_operation + read_operation_outputs, free_harq_context_entry on TB success, free_queue. HARQ soft bits never
/// leave HBM (is_external_harq_supported() == true), so the rx_buffer_pool is created with external_soft_bits = true.
#pragma once

#include "cuda_pusch_dec_device.h"
#include "srsran/hal/phy/upper/channel_processors/pusch/hw_accelerator_pusch_dec.h"

namespace srsran {
namespace hal {

class hw_accelerator_pusch_dec_cuda_impl : public hw_accelerator_pusch_dec
{
public:
  explicit hw_accelerator_pusch_dec_cuda_impl(std::shared_ptr<cuda_pusch_dec_device> device_) : device(std::move(device_)) {}

  // See hw_accelerator_pusch_dec for the documentation.
  void reserve_queue() override;
  void free_queue() override;
  bool enqueue_operation(span<const int8_t> data, span<const int8_t> soft_data = {}, unsigned cb_index = 0) override;
  bool dequeue_operation(span<uint8_t> data, span<int8_t> soft_data = {}, unsigned segment_index = 0) override;
  void configure_operation(const hw_pusch_decoder_configuration& config, unsigned cb_index = 0) override;
  void read_operation_outputs(hw_pusch_decoder_outputs& out, unsigned cb_index = 0, unsigned absolute_cb_id = 0) override;
  void free_harq_context_entry(unsigned absolute_cb_id) override;
  bool is_external_harq_supported() const override { return true; }

private:
  std::shared_ptr<cuda_pusch_dec_device> device;
  /// Held from reserve_queue() to free_queue(): one transport block owns the handle's operation table at a time. The lock is
  /// owned by this object (released by free_queue, or by the destructor if the owner never got there), and the device mutex
  /// is recursive: pusch_decoder_hw_impl recomputes CRC16 / CRC24A between reserve_queue and free_queue
  /// (pusch_decoder_hw_impl.cpp:353-360,384), which may be "cuda" crc_calculators of the same device.
  std::unique_lock<std::recursive_mutex> queue_lock;
};

} // namespace hal
} // namespace srsran
