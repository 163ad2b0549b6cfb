/// \file
/// \brief Shared ownership of one srsran_cuda_pusch_dec handle (one B200, its HBM-resident HARQ slots and batch contexts).
///
/// Host side of the B200 PUSCH channel-decoding path, to be dropped into lib/hal of srsRAN Project next to the ACC100
/// implementation (lib/hal/phy/upper/channel_processors/pusch/). Everything below this layer is the C ABI of
/// include/srsran_cuda_pusch_dec.h; nothing here computes on the CPU.
#pragma once

#include "srsran_cuda_pusch_dec.h"
#include "srsran/support/executors/task_executor.h"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

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

/// One transport block handed to the device's slot aggregator.
struct cuda_tb_request {
  srsran_cuda_pusch_dec_tb_config cfg = {};
  /// Page-locked soft bits (ignored when ingest_stream >= 0: the soft bits were streamed with stream_begin / stream_push).
  const int8_t* llrs          = nullptr;
  uint32_t      nof_llrs      = 0;
  int           ingest_stream = -1;
  /// HARQ slots: the rx buffer's absolute code-block identifiers.
  uint32_t cb_ids[SRSRAN_CUDA_MAX_NOF_SEGMENTS] = {};
  uint32_t nof_cbs                              = 0;
};

/// What the device reports for a completed transport block.
struct cuda_tb_completion {
  srsran_cuda_pusch_dec_tb_result result = {};
  /// Decoded transport block in the library's page-locked result buffer (valid during the callback), or nullptr if the
  /// reference would not have written it (a code-block CRC failed, pusch_decoder_impl.cpp:405-418).
  const uint8_t* tb_data = nullptr;
  uint8_t        cb_crc[SRSRAN_CUDA_MAX_NOF_SEGMENTS]        = {};
  uint32_t       cb_iterations[SRSRAN_CUDA_MAX_NOF_SEGMENTS] = {};
};

/// One CUDA device context shared by every accelerator instance a factory creates: the absolute code-block identifiers
/// of the rx_buffer_pool index the same HBM HARQ slots no matter which pusch_decoder_hw_impl instance decodes.
///
/// It is also the slot aggregator of the native decoders (pusch_decoder_cuda_impl): the transport blocks that many decoder
/// instances finish collecting within a few microseconds of each other (the PUSCH allocations of one slot) leave as ONE
/// batch - one set of copies and kernel launches, four code blocks of the same shape per CTA - instead of one launch set per
/// transport block. A flusher thread submits the open batch when it holds max_tbs_per_batch transport blocks or when its
/// oldest one has waited flush_deadline; a completion thread waits for the oldest batch in flight (outside the handle's
/// lock), collects the results and runs the callbacks. A busy accelerator (every batch context in flight and unpolled) is
/// back-pressure, not an error: the flusher waits for the completion thread and submits again.
class cuda_pusch_dec_device
{
public:
  using completion_fn = std::function<void(const cuda_tb_completion&)>;

  explicit cuda_pusch_dec_device(const cuda_hwacc_pusch_dec_configuration& cfg);
  ~cuda_pusch_dec_device();
  cuda_pusch_dec_device(const cuda_pusch_dec_device&)            = delete;
  cuda_pusch_dec_device& operator=(const cuda_pusch_dec_device&) = delete;

  srsran_cuda_pusch_dec_t* get() { return handle; }
  /// A handle is thread-compatible: everything sharing it serialises its calls. Recursive, so that a caller holding it
  /// (the hal accelerator between reserve_queue and free_queue) may use the "cuda" CRC calculators of the same device.
  std::recursive_mutex& mutex() { return mtx; }

  /// Batch size and deadline of the aggregator (defaults: 64 transport blocks, 50 microseconds).
  void set_aggregation(unsigned max_tbs_per_batch, std::chrono::microseconds flush_deadline);
  /// Queues one transport block. Once the device has decoded it, \c on_done runs - as a task of \c executor if one is
  /// given (the copy of the transport block out of the library's result buffer then happens on the executor's threads, in
  /// parallel for the transport blocks of a batch), else on the device's completion thread. Thread-safe.
  void submit(const cuda_tb_request& request, completion_fn on_done, task_executor* executor = nullptr);
  /// Number of batches submitted so far and transport blocks in them (diagnostics: mean batch size).
  std::pair<uint64_t, uint64_t> aggregation_stats() const { return {nof_batches.load(), nof_batched_tbs.load()}; }

private:
  struct pending_tb {
    cuda_tb_request                       req;
    completion_fn                         on_done;
    task_executor*                        executor;
    std::chrono::steady_clock::time_point arrival;
  };
  /// A batch in flight; shared with the callbacks running on executors: the last one to finish consumes the tickets (only
  /// then may the library reuse the batch's result buffer the callbacks copy from).
  struct flying_batch {
    std::vector<int>                tickets;
    std::vector<completion_fn>      callbacks;
    std::vector<task_executor*>     executors;
    std::vector<uint32_t>           nof_cbs;
    std::vector<cuda_tb_completion> done;
    std::atomic<unsigned>           remaining{0};
  };
  void consume(flying_batch& fb);

  void flusher_loop();
  void completer_loop();
  void start_threads();

  srsran_cuda_pusch_dec_t* handle = nullptr;
  std::recursive_mutex     mtx;

  std::mutex                            agg_mtx;
  std::condition_variable               cv_pending;  // flusher: work arrived / shutdown
  std::condition_variable               cv_flying;   // completer: batch in flight / shutdown
  std::condition_variable               cv_progress; // flusher: a batch completed (busy accelerator)
  std::deque<pending_tb>                pending;
  std::deque<std::shared_ptr<flying_batch>> flying;
  unsigned                              max_batch = 64;
  std::chrono::microseconds             deadline{50};
  bool                                  stop    = false;
  bool                                  started = false;
  std::thread                           flusher, completer;
  std::atomic<uint64_t>                 nof_batches{0}, nof_batched_tbs{0};

public:
  /// Where the two threads spent their time, in microseconds (diagnostics of the benchmark harness).
  struct thread_times {
    std::atomic<uint64_t> submit_us{0}, busy_wait_us{0}, device_wait_us{0}, collect_us{0}, callbacks_us{0}, consume_us{0};
  } times;
};

} // namespace hal
} // namespace srsran
