#include "hw_accelerator_pusch_dec_cuda_impl.h"
#include "srsran/support/srsran_assert.h"

using namespace srsran;
using namespace hal;

void hw_accelerator_pusch_dec_cuda_impl::reserve_queue()
{
  if (!queue_lock.owns_lock()) {
    queue_lock = std::unique_lock<std::recursive_mutex>(device->mutex());
  }
  int st     = srsran_cuda_pusch_dec_reserve_queue(device->get());
  srsran_assert(st == SRSRAN_CUDA_OK, "CUDA PUSCH decoder: reserve_queue failed ({}).", st);
}

void hw_accelerator_pusch_dec_cuda_impl::free_queue()
{
  srsran_cuda_pusch_dec_free_queue(device->get());
  if (queue_lock.owns_lock()) {
    queue_lock.unlock();
  }
}

void hw_accelerator_pusch_dec_cuda_impl::configure_operation(const hw_pusch_decoder_configuration& config,
                                                             unsigned                              cb_index)
{
  srsran_cuda_pusch_dec_cb_config c = {};
  c.base_graph                      = (config.base_graph_index == ldpc_base_graph_type::BG1) ? 1 : 2;
  c.modulation                      = get_bits_per_symbol(config.modulation);
  c.nof_segments                    = config.nof_segments;
  c.rv                              = config.rv;
  c.cw_length                       = config.cw_length;
  c.lifting_size                    = config.lifting_size;
  c.Ncb                             = config.Ncb;
  c.Nref                            = config.Nref;
  c.nof_segment_bits                = config.nof_segment_bits;
  c.nof_filler_bits                 = config.nof_filler_bits;
  c.max_nof_ldpc_iterations         = config.max_nof_ldpc_iterations;
  c.use_early_stop                  = config.use_early_stop ? 1 : 0;
  c.new_data                        = config.new_data ? 1 : 0;
  c.cb_crc_len                      = config.cb_crc_len;
  c.cb_crc_type                     = static_cast<uint32_t>(config.cb_crc_type);
  c.absolute_cb_id                  = config.absolute_cb_id;
  int st                            = srsran_cuda_pusch_dec_configure(device->get(), cb_index, &c);
  srsran_assert(st == SRSRAN_CUDA_OK, "CUDA PUSCH decoder: invalid configuration of code block {} ({}).", cb_index, st);
}

bool hw_accelerator_pusch_dec_cuda_impl::enqueue_operation(span<const int8_t> data,
                                                           span<const int8_t> soft_data,
                                                           unsigned           cb_index)
{
  // soft_data is empty by contract (external HARQ): the combined soft bits live in HBM.
  int st = srsran_cuda_pusch_dec_enqueue(device->get(), cb_index, data.data(), data.size(), soft_data.data(), soft_data.size());
  srsran_assert(st >= 0,
                "CUDA PUSCH decoder: enqueue of code block {} failed ({}): {}.",
                cb_index,
                st,
                srsran_cuda_pusch_dec_last_error(device->get()));
  return st == 1;
}

bool hw_accelerator_pusch_dec_cuda_impl::dequeue_operation(span<uint8_t> data, span<int8_t> soft_data, unsigned segment_index)
{
  int st = srsran_cuda_pusch_dec_dequeue(
      device->get(), segment_index, data.data(), data.size(), soft_data.empty() ? nullptr : soft_data.data(), soft_data.size());
  srsran_assert(st >= 0,
                "CUDA PUSCH decoder: dequeue of code block {} failed ({}): {}.",
                segment_index,
                st,
                srsran_cuda_pusch_dec_last_error(device->get()));
  return st == 1;
}

void hw_accelerator_pusch_dec_cuda_impl::read_operation_outputs(hw_pusch_decoder_outputs& out,
                                                                unsigned                  cb_index,
                                                                unsigned /*absolute_cb_id*/)
{
  int      crc_pass = 0;
  uint32_t iters    = 0;
  int      st       = srsran_cuda_pusch_dec_read_outputs(device->get(), cb_index, &crc_pass, &iters);
  srsran_assert(st == SRSRAN_CUDA_OK, "CUDA PUSCH decoder: read_operation_outputs of code block {} failed ({}).", cb_index, st);
  out.CRC_pass            = crc_pass != 0;
  out.nof_ldpc_iterations = iters;
}

void hw_accelerator_pusch_dec_cuda_impl::free_harq_context_entry(unsigned absolute_cb_id)
{
  srsran_cuda_pusch_dec_free_harq(device->get(), absolute_cb_id);
}
