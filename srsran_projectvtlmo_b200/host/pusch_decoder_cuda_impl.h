/// \file
/// \brief CUDA implementation of the reference's pusch_decoder interface
/// (include/srsran/phy/upper/channel_processors/pusch/pusch_decoder.h:54-99), the sibling of pusch_decoder_impl (software)
/// and pusch_decoder_hw_impl (ACC100): one transport block per instance at a time, soft bits streamed to the B200 as the
/// UL-SCH demultiplexer delivers them, one device-side pass (dematch, LDPC, CB CRC, TB assembly, TB CRC) at
/// on_end_softbits, completion through pusch_decoder_notifier::on_sch_data.
#pragma once

#include "cuda_pusch_dec_device.h"
#include "srsran/phy/upper/channel_processors/pusch/factories.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_buffer.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_notifier.h"
#include "srsran/phy/upper/unique_rx_buffer.h"
#include "srsran/support/executors/task_executor.h"
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <optional>
#include <vector>

namespace srsran {

class pusch_decoder_cuda_impl : public pusch_decoder, private pusch_decoder_buffer
{
public:
  /// \param[in] device_     CUDA device context shared with the other decoder instances (HBM HARQ slots indexed by the
  ///                        rx_buffer_pool's absolute code-block identifiers; pool created with external_soft_bits).
  /// \param[in] executor_   Optional executor that waits for the device and notifies (asynchronous completion, like the
  ///                        software decoder's code-block executor); if null, on_end_softbits blocks until the TB is done.
  /// \param[in] nof_prb     Maximum number of PRB, \param[in] nof_layers maximum number of layers (soft-bit staging size,
  ///                        pusch_decoder_impl.h:86).
  pusch_decoder_cuda_impl(std::shared_ptr<hal::cuda_pusch_dec_device> device_,
                          task_executor*                              executor_,
                          unsigned                                    nof_prb,
                          unsigned                                    nof_layers);
  /// Several GPUs: a transport block is decoded on devices_[first absolute code-block id % devices_.size()]. The rx buffer
  /// of a HARQ process keeps its code-block ids from the first transmission to its release, so every retransmission finds
  /// its soft bits on the device that holds them (sticky sharding, no device-to-device traffic).
  pusch_decoder_cuda_impl(std::vector<std::shared_ptr<hal::cuda_pusch_dec_device>> devices_,
                          task_executor*                                           executor_,
                          unsigned                                                 nof_prb,
                          unsigned                                                 nof_layers);
  ~pusch_decoder_cuda_impl() override;

  // See interface for the documentation.
  pusch_decoder_buffer& new_data(span<uint8_t>           transport_block,
                                 unique_rx_buffer        rm_buffer,
                                 pusch_decoder_notifier& notifier,
                                 const configuration&    cfg) override;

  // See interface for the documentation.
  void set_nof_softbits(units::bits nof_softbits) override;

private:
  /// Same life cycle as pusch_decoder_impl::internal_states (pusch_decoder_impl.h:113-137).
  enum class internal_states : uint8_t { idle = 0, collecting, decoding };

  // See pusch_decoder_buffer for the documentation.
  span<log_likelihood_ratio> get_next_block_view(unsigned block_size) override;
  void                       on_new_softbits(span<const log_likelihood_ratio> softbits) override;
  void                       on_end_softbits() override;

  /// Pushes the soft bits collected since the last push to the device (no-op while the total is unknown).
  void push_pending();
  /// Completion callback of the device's slot aggregator (completion thread): copies the transport block and the per
  /// code block outputs, then hands over to the executor or to the thread blocked in on_end_softbits.
  void on_device_completion(const hal::cuda_tb_completion& completion);
  /// Updates the rx buffer (CRC flags, release / unlock), returns to idle and notifies.
  void finish_and_notify();

  std::vector<std::shared_ptr<hal::cuda_pusch_dec_device>> devices;
  /// Device of the transport block in flight (chosen in new_data).
  hal::cuda_pusch_dec_device* device = nullptr;
  task_executor*              executor;
  /// Page-locked soft-bit staging (the UL-SCH demultiplexer writes into it through get_next_block_view).
  log_likelihood_ratio*        softbits_buffer = nullptr;
  unsigned                     softbits_capacity;
  unsigned                     softbits_count  = 0;
  unsigned                     softbits_pushed = 0;
  int                          ingest_stream   = -1;
  std::optional<units::bits>   nof_ulsch_softbits;
  span<uint8_t>                transport_block;
  unique_rx_buffer             unique_rm_buffer;
  pusch_decoder_notifier*      result_notifier = nullptr;
  configuration                current_config;
  unsigned                     nof_codeblocks = 0;
  std::atomic<internal_states> current_state{internal_states::idle};
  /// Result of the transport block in flight (written by the completion thread before it signals).
  srsran_cuda_pusch_dec_tb_result device_result = {};
  uint8_t                         cb_crc[SRSRAN_CUDA_MAX_NOF_SEGMENTS];
  uint32_t                        cb_iterations[SRSRAN_CUDA_MAX_NOF_SEGMENTS];
  std::mutex                      completion_mutex;
  std::condition_variable         completion_cv;
  std::atomic<bool>               completed{false};
};

/// Factory of CUDA PUSCH decoders ("cuda" flavour next to create_pusch_decoder_factory_sw / _hw, pusch/factories.h:57-80).
/// Returns nullptr if the device context is null (no usable CUDA device: there is no CPU fallback).
std::shared_ptr<pusch_decoder_factory> create_pusch_decoder_factory_cuda(std::shared_ptr<hal::cuda_pusch_dec_device> device,
                                                                         task_executor* executor,
                                                                         unsigned       nof_prb,
                                                                         unsigned       nof_layers);
/// The same over several GPUs (one device context each); nullptr if the list is empty or holds a null context.
std::shared_ptr<pusch_decoder_factory>
create_pusch_decoder_factory_cuda(std::vector<std::shared_ptr<hal::cuda_pusch_dec_device>> devices,
                                  task_executor*                                           executor,
                                  unsigned                                                 nof_prb,
                                  unsigned                                                 nof_layers);

} // namespace srsran
