/// \file
/// \brief "cuda" variants of the channel-coding factories of the PUSCH decoding path
/// (include/srsran/phy/upper/channel_coding/channel_coding_factories.h:43-77): ldpc_decoder, ldpc_rate_dematcher and
/// crc_calculator objects that compute on the B200 through the C ABI. They serve the unit-level interfaces (one
/// synchronous call per code block, host buffers); the throughput path is the hal accelerator / TB-level ABI.
#pragma once

#include "cuda_pusch_dec_device.h"
#include "srsran/phy/upper/channel_coding/channel_coding_factories.h"

namespace srsran {

/// `dec_type == "cuda"` branch of create_ldpc_decoder_factory_sw (channel_coding_factories.cpp:100-124).
std::shared_ptr<ldpc_decoder_factory> create_ldpc_decoder_factory_cuda(std::shared_ptr<hal::cuda_pusch_dec_device> device);
/// `dematcher_type == "cuda"` branch of create_ldpc_rate_dematcher_factory_sw.
std::shared_ptr<ldpc_rate_dematcher_factory>
create_ldpc_rate_dematcher_factory_cuda(std::shared_ptr<hal::cuda_pusch_dec_device> device);
/// `type == "cuda"` branch of create_crc_calculator_factory_sw (CRC24A, CRC24B and CRC16 - the PUSCH polynomials; the
/// factory returns nullptr for the others).
std::shared_ptr<crc_calculator_factory> create_crc_calculator_factory_cuda(std::shared_ptr<hal::cuda_pusch_dec_device> device);

} // namespace srsran
