/// \file
/// \brief Factory of CUDA PUSCH decoder accelerators. Its own header because the ACC100 factory header pulls DPDK
/// (include/srsran/hal/phy/upper/channel_processors/pusch/hw_accelerator_factories.h:25 -> bbdev_acc.h -> rte_bbdev.h).
#pragma once

#include "cuda_pusch_dec_device.h"
#include "hw_accelerator_pdsch_enc_cuda_impl.h"
#include "srsran/hal/phy/upper/channel_processors/hw_accelerator_pdsch_enc_factory.h"
#include "srsran/hal/phy/upper/channel_processors/pusch/hw_accelerator_pusch_dec_factory.h"
#include <memory>

namespace srsran {
namespace hal {

/// Returns a factory whose accelerators share one CUDA device context, or nullptr if no CUDA device is usable (there is
/// no CPU fallback; the caller then keeps the software pusch_decoder factory, as it does when the ACC100 is absent).
std::shared_ptr<hw_accelerator_pusch_dec_factory>
create_cuda_pusch_dec_acc_factory(const cuda_hwacc_pusch_dec_configuration& accelerator_config);

/// PDSCH mirror (include/srsran/hal/phy/upper/channel_processors/hw_accelerator_factories.h: create_bbdev_pdsch_enc_acc_factory):
/// returns a factory of CUDA PDSCH encoder accelerators (each create() opens its own handle: one per pdsch_encoder_hw_impl,
/// used from one thread at a time), or nullptr if no CUDA device is usable.
std::shared_ptr<hw_accelerator_pdsch_enc_factory>
create_cuda_pdsch_enc_acc_factory(const cuda_hwacc_pdsch_enc_configuration& accelerator_config);

} // namespace hal
} // namespace srsran
