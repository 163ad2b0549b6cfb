#include "hw_accelerator_pdsch_enc_cuda_impl.h"
#include "srsran/support/error_handling.h"
#include <stdexcept>

using namespace srsran;
using namespace hal;

hw_accelerator_pdsch_enc_cuda_impl::hw_accelerator_pdsch_enc_cuda_impl(const cuda_hwacc_pdsch_enc_configuration& config) :
  cfg(config)
{
  handle = srsran_cuda_pdsch_enc_create(cfg.device, cfg.max_nof_operations);
  if (handle == nullptr) {
    throw std::runtime_error(std::string("cuda PDSCH encoder accelerator: ") + srsran_cuda_pdsch_enc_create_error());
  }
}

hw_accelerator_pdsch_enc_cuda_impl::~hw_accelerator_pdsch_enc_cuda_impl()
{
  srsran_cuda_pdsch_enc_destroy(handle);
}

void hw_accelerator_pdsch_enc_cuda_impl::configure_operation(const hw_pdsch_encoder_configuration& config, unsigned cb_index)
{
  srsran_cuda_pdsch_enc_config c = {};
  c.nof_tb_bits                  = config.nof_tb_bits;
  c.nof_tb_crc_bits              = config.nof_tb_crc_bits;
  c.base_graph                   = (config.base_graph_index == ldpc_base_graph_type::BG1) ? 1 : 2;
  c.modulation                   = get_bits_per_symbol(config.modulation);
  c.nof_segments                 = config.nof_segments;
  c.nof_short_segments           = config.nof_short_segments;
  c.rv                           = config.rv;
  c.cw_length_a                  = config.cw_length_a;
  c.cw_length_b                  = config.cw_length_b;
  c.lifting_size                 = config.lifting_size;
  c.Ncb                          = config.Ncb;
  c.Nref                         = config.Nref;
  c.nof_segment_bits             = config.nof_segment_bits;
  c.nof_filler_bits              = config.nof_filler_bits;
  c.rm_length                    = config.rm_length;
  for (unsigned i = 0; i != config.tb_crc.size() && i != 3; ++i) {
    c.tb_crc[i] = config.tb_crc[i];
  }
  c.cb_mode = config.cb_mode ? 1 : 0;
  int st    = srsran_cuda_pdsch_enc_configure(handle, cb_index, &c);
  report_fatal_error_if_not(st == SRSRAN_CUDA_OK, "cuda PDSCH encoder: configure_operation: {}", srsran_cuda_pdsch_enc_last_error(handle));
}

bool hw_accelerator_pdsch_enc_cuda_impl::enqueue_operation(span<const uint8_t> data, span<const uint8_t> /*aux_data*/, unsigned cb_index)
{
  int st = srsran_cuda_pdsch_enc_enqueue(handle, cb_index, data.data(), static_cast<uint32_t>(data.size()));
  report_fatal_error_if_not(st >= 0, "cuda PDSCH encoder: enqueue_operation: {}", srsran_cuda_pdsch_enc_last_error(handle));
  return st == 1;
}

bool hw_accelerator_pdsch_enc_cuda_impl::dequeue_operation(span<uint8_t> data, span<uint8_t> packed_data, unsigned segment_index)
{
  int st = srsran_cuda_pdsch_enc_dequeue(handle,
                                         segment_index,
                                         data.data(),
                                         static_cast<uint32_t>(data.size()),
                                         packed_data.empty() ? nullptr : packed_data.data(),
                                         static_cast<uint32_t>(packed_data.size()));
  report_fatal_error_if_not(st >= 0, "cuda PDSCH encoder: dequeue_operation: {}", srsran_cuda_pdsch_enc_last_error(handle));
  return st == 1;
}
