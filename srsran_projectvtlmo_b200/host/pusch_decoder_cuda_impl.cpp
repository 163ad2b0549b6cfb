#include "pusch_decoder_cuda_impl.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_result.h"
#include "srsran/ran/pusch/pusch_constants.h"
#include "srsran/ran/sch/sch_segmentation.h"
#include "srsran/support/srsran_assert.h"
#include <algorithm>
#include <cstring>
#include <mutex>

using namespace srsran;

namespace {
/// Blocks smaller than this are merged with the next one before they are copied to the device.
constexpr unsigned MIN_PUSH_SOFTBITS = 16384;
} // namespace

pusch_decoder_cuda_impl::pusch_decoder_cuda_impl(std::shared_ptr<hal::cuda_pusch_dec_device> device_,
                                                 task_executor*                              executor_,
                                                 unsigned                                    nof_prb,
                                                 unsigned                                    nof_layers) :
  pusch_decoder_cuda_impl(std::vector<std::shared_ptr<hal::cuda_pusch_dec_device>>{std::move(device_)},
                          executor_,
                          nof_prb,
                          nof_layers)
{
}

pusch_decoder_cuda_impl::pusch_decoder_cuda_impl(std::vector<std::shared_ptr<hal::cuda_pusch_dec_device>> devices_,
                                                 task_executor*                                           executor_,
                                                 unsigned                                                 nof_prb,
                                                 unsigned                                                 nof_layers) :
  devices(std::move(devices_)),
  executor(executor_),
  softbits_capacity(pusch_constants::get_max_codeword_size(nof_prb, nof_layers).value())
{
  srsran_assert(!devices.empty(), "No CUDA device context.");
  for (const auto& d : devices) {
    srsran_assert(d, "Invalid CUDA device context.");
  }
  static_assert(sizeof(log_likelihood_ratio) == sizeof(int8_t), "LLRs are int8");
  softbits_buffer = static_cast<log_likelihood_ratio*>(srsran_cuda_pusch_dec_host_alloc(softbits_capacity));
  report_fatal_error_if_not(softbits_buffer != nullptr, "Cannot allocate page-locked soft-bit staging memory.");
}

pusch_decoder_cuda_impl::~pusch_decoder_cuda_impl()
{
  srsran_cuda_pusch_dec_host_free(softbits_buffer);
}

pusch_decoder_buffer& pusch_decoder_cuda_impl::new_data(span<uint8_t>           transport_block_,
                                                        unique_rx_buffer        unique_rm_buffer_,
                                                        pusch_decoder_notifier& notifier,
                                                        const configuration&    cfg)
{
  internal_states previous_state = current_state.exchange(internal_states::collecting);
  srsran_assert(previous_state == internal_states::idle, "Invalid state: a transport block is already being processed.");

  transport_block  = transport_block_;
  unique_rm_buffer = std::move(unique_rm_buffer_);
  result_notifier  = &notifier;
  current_config   = cfg;
  softbits_count   = 0;
  softbits_pushed  = 0;
  ingest_stream    = -1;
  nof_ulsch_softbits.reset();

  unsigned tb_size = transport_block.size() * 8;
  nof_codeblocks   = ldpc::compute_nof_codeblocks(units::bits(tb_size), cfg.base_graph);
  srsran_assert(nof_codeblocks == unique_rm_buffer->get_nof_codeblocks(),
                "Wrong number of codeblocks {} (expected {}).",
                unique_rm_buffer->get_nof_codeblocks(),
                nof_codeblocks);

  // Sticky sharding: the HARQ process (its rx buffer's code-block ids) decides the GPU.
  device = devices[unique_rm_buffer->get_absolute_codeblock_id(0) % devices.size()].get();

  // Reset CRCs if new data is flagged (pusch_decoder_impl.cpp:131-135). The device ignores its own flags of a slot when
  // the configuration says new data, so only the host view needs clearing.
  if (cfg.new_data) {
    unique_rm_buffer->reset_codeblocks_crc();
  }
  return *this;
}

void pusch_decoder_cuda_impl::set_nof_softbits(units::bits nof_softbits)
{
  // No effect before new_data or once decoding has started (pusch_decoder.h:91-97).
  if (current_state.load() != internal_states::collecting || nof_ulsch_softbits.has_value()) {
    return;
  }
  nof_ulsch_softbits = nof_softbits;
  // From now on every block is copied to the device as it arrives.
  std::lock_guard<std::recursive_mutex> lock(device->mutex());
  ingest_stream = srsran_cuda_pusch_dec_stream_begin(device->get(), nof_softbits.value());
  if (ingest_stream < 0) {
    // Every ingest stream of the device is taken by other decoders: this transport block is copied in one piece at
    // on_end_softbits instead (same result, only the overlap with the reception is lost).
    ingest_stream = -1;
  }
}

span<log_likelihood_ratio> pusch_decoder_cuda_impl::get_next_block_view(unsigned block_size)
{
  srsran_assert(current_state.load() == internal_states::collecting, "Invalid state: not collecting soft bits.");
  srsran_assert(softbits_count + block_size <= softbits_capacity,
                "The sum of current buffer number of elements (i.e., {}) and the block size (i.e., {}), exceeds the "
                "total number of elements of the buffer (i.e., {}).",
                softbits_count,
                block_size,
                softbits_capacity);
  return span<log_likelihood_ratio>(softbits_buffer + softbits_count, block_size);
}

void pusch_decoder_cuda_impl::on_new_softbits(span<const log_likelihood_ratio> softbits)
{
  srsran_assert(current_state.load() == internal_states::collecting, "Invalid state: not collecting soft bits.");
  span<log_likelihood_ratio> block = get_next_block_view(softbits.size());
  // Copy only if the soft bits were not written through the view (pusch_decoder_impl.cpp:188-191).
  if (block.data() != softbits.data()) {
    std::copy(softbits.begin(), softbits.end(), block.begin());
  }
  softbits_count += softbits.size();
  if (ingest_stream >= 0 && softbits_count - softbits_pushed >= MIN_PUSH_SOFTBITS) {
    push_pending();
  }
}

void pusch_decoder_cuda_impl::push_pending()
{
  if (ingest_stream < 0 || softbits_pushed == softbits_count) {
    return;
  }
  std::lock_guard<std::recursive_mutex> lock(device->mutex());
  int st = srsran_cuda_pusch_dec_stream_push(device->get(),
                                             ingest_stream,
                                             reinterpret_cast<const int8_t*>(softbits_buffer + softbits_pushed),
                                             softbits_count - softbits_pushed);
  report_fatal_error_if_not(st == SRSRAN_CUDA_OK, "CUDA PUSCH decoder: {}", srsran_cuda_pusch_dec_last_error(device->get()));
  softbits_pushed = softbits_count;
}

void pusch_decoder_cuda_impl::on_end_softbits()
{
  internal_states previous_state = current_state.exchange(internal_states::decoding);
  srsran_assert(previous_state == internal_states::collecting, "Invalid state: not collecting soft bits.");
  if (nof_ulsch_softbits.has_value()) {
    srsran_assert(nof_ulsch_softbits->value() == softbits_count,
                  "The number of soft bits (i.e., {}) is not the announced one (i.e., {}).",
                  softbits_count,
                  nof_ulsch_softbits->value());
  }

  hal::cuda_tb_request req;
  req.cfg.tbs_bits            = transport_block.size() * 8;
  req.cfg.base_graph          = (current_config.base_graph == ldpc_base_graph_type::BG1) ? 1 : 2;
  req.cfg.rv                  = current_config.rv;
  req.cfg.modulation          = get_bits_per_symbol(current_config.mod);
  req.cfg.Nref                = current_config.Nref;
  req.cfg.nof_layers          = current_config.nof_layers;
  req.cfg.nof_ldpc_iterations = current_config.nof_ldpc_iterations;
  req.cfg.use_early_stop      = current_config.use_early_stop ? 1 : 0;
  req.cfg.new_data            = current_config.new_data ? 1 : 0;
  // HARQ slots: the rx buffer's absolute code-block identifiers (rx_buffer.h:50-53), not necessarily consecutive.
  req.nof_cbs = nof_codeblocks;
  for (unsigned i = 0; i != nof_codeblocks; ++i) {
    req.cb_ids[i] = unique_rm_buffer->get_absolute_codeblock_id(i);
  }
  if (ingest_stream >= 0) {
    push_pending();
  }
  req.llrs          = reinterpret_cast<const int8_t*>(softbits_buffer);
  req.nof_llrs      = softbits_count;
  req.ingest_stream = ingest_stream;
  ingest_stream     = -1;

  // The transport block joins the device's open batch (slot aggregator); the completion thread of the device copies the
  // result out and either posts the notification to the executor or wakes this thread up.
  completed.store(false);
  device->submit(req, [this](const hal::cuda_tb_completion& c) { on_device_completion(c); }, executor);
  if (executor != nullptr) {
    return;
  }
  {
    std::unique_lock<std::mutex> lock(completion_mutex);
    completion_cv.wait(lock, [this]() { return completed.load(); });
  }
  finish_and_notify();
}

void pusch_decoder_cuda_impl::on_device_completion(const hal::cuda_tb_completion& c)
{
  // Runs as a task of the executor (asynchronous decoders) or on the device's completion thread.
  if (c.tb_data != nullptr) {
    // The reference writes the transport block only when every code block is ok (pusch_decoder_impl.cpp:405-418).
    std::memcpy(transport_block.data(), c.tb_data, transport_block.size());
  }
  device_result = c.result;
  std::memcpy(cb_crc, c.cb_crc, nof_codeblocks);
  std::memcpy(cb_iterations, c.cb_iterations, nof_codeblocks * sizeof(uint32_t));
  if (executor != nullptr) {
    finish_and_notify(); // already a task of the executor (cuda_pusch_dec_device::submit)
    return;
  }
  {
    std::lock_guard<std::mutex> lock(completion_mutex);
    completed.store(true);
  }
  completion_cv.notify_one();
}

void pusch_decoder_cuda_impl::finish_and_notify()
{
  pusch_decoder_result stats;
  stats.tb_crc_ok            = device_result.tb_crc_ok != 0;
  stats.nof_codeblocks_total = nof_codeblocks;
  // sample_statistics keeps a running mean: feed the observations code block by code block, in order
  // (pusch_decoder_impl.cpp:357-363).
  span<bool> cb_crcs = unique_rm_buffer->get_codeblocks_crc();
  for (unsigned i = 0; i != nof_codeblocks; ++i) {
    if (cb_iterations[i] != 0xffffffffU) {
      stats.ldpc_decoder_stats.update(cb_iterations[i]);
    }
    cb_crcs[i] = cb_crc[i] != 0;
  }

  // Release soft buffer if the CRC is OK, otherwise unlock (pusch_decoder_impl.cpp:431-436).
  if (stats.tb_crc_ok) {
    unique_rm_buffer.release();
  } else {
    unique_rm_buffer.unlock();
  }

  pusch_decoder_notifier* notifier = result_notifier;
  internal_states previous_state   = current_state.exchange(internal_states::idle);
  srsran_assert(previous_state == internal_states::decoding, "Invalid state: not decoding.");

  // Finally report decoding result.
  notifier->on_sch_data(stats);
}

namespace {

class pusch_decoder_factory_cuda : public pusch_decoder_factory
{
public:
  pusch_decoder_factory_cuda(std::vector<std::shared_ptr<hal::cuda_pusch_dec_device>> devices_,
                             task_executor*                                           executor_,
                             unsigned                                                 nof_prb_,
                             unsigned                                                 nof_layers_) :
    devices(std::move(devices_)), executor(executor_), nof_prb(nof_prb_), nof_layers(nof_layers_)
  {
  }

  std::unique_ptr<pusch_decoder> create() override
  {
    return std::make_unique<pusch_decoder_cuda_impl>(devices, executor, nof_prb, nof_layers);
  }

private:
  std::vector<std::shared_ptr<hal::cuda_pusch_dec_device>> devices;
  task_executor*                              executor;
  unsigned                                    nof_prb;
  unsigned                                    nof_layers;
};

} // namespace

std::shared_ptr<pusch_decoder_factory>
srsran::create_pusch_decoder_factory_cuda(std::shared_ptr<hal::cuda_pusch_dec_device> device,
                                          task_executor*                              executor,
                                          unsigned                                    nof_prb,
                                          unsigned                                    nof_layers)
{
  if (!device) {
    return nullptr;
  }
  return create_pusch_decoder_factory_cuda(
      std::vector<std::shared_ptr<hal::cuda_pusch_dec_device>>{std::move(device)}, executor, nof_prb, nof_layers);
}

std::shared_ptr<pusch_decoder_factory>
srsran::create_pusch_decoder_factory_cuda(std::vector<std::shared_ptr<hal::cuda_pusch_dec_device>> devices,
                                          task_executor*                                           executor,
                                          unsigned                                                 nof_prb,
                                          unsigned                                                 nof_layers)
{
  if (devices.empty()) {
    return nullptr;
  }
  for (const auto& d : devices) {
    if (!d) {
      return nullptr;
    }
  }
  return std::make_shared<pusch_decoder_factory_cuda>(std::move(devices), executor, nof_prb, nof_layers);
}
