#include "hw_accelerator_factories_cuda.h"
#include "hw_accelerator_pusch_dec_cuda_impl.h"

using namespace srsran;
using namespace hal;

namespace {

class hw_accelerator_pusch_dec_cuda_factory : public hw_accelerator_pusch_dec_factory
{
public:
  explicit hw_accelerator_pusch_dec_cuda_factory(std::shared_ptr<cuda_pusch_dec_device> device_) : device(std::move(device_)) {}

  std::unique_ptr<hw_accelerator_pusch_dec> create() override
  {
    return std::make_unique<hw_accelerator_pusch_dec_cuda_impl>(device);
  }

private:
  std::shared_ptr<cuda_pusch_dec_device> device;
};

class hw_accelerator_pdsch_enc_cuda_factory : public hw_accelerator_pdsch_enc_factory
{
public:
  explicit hw_accelerator_pdsch_enc_cuda_factory(const cuda_hwacc_pdsch_enc_configuration& cfg_) : cfg(cfg_) {}

  std::unique_ptr<hw_accelerator_pdsch_enc> create() override
  {
    try {
      return std::make_unique<hw_accelerator_pdsch_enc_cuda_impl>(cfg);
    } catch (const std::exception&) {
      return nullptr;
    }
  }

private:
  cuda_hwacc_pdsch_enc_configuration cfg;
};

} // namespace

std::shared_ptr<hw_accelerator_pdsch_enc_factory>
srsran::hal::create_cuda_pdsch_enc_acc_factory(const cuda_hwacc_pdsch_enc_configuration& accelerator_config)
{
  // Probe once: without a usable device the factory itself is refused, as for the PUSCH decoder.
  srsran_cuda_pdsch_enc_t* probe = srsran_cuda_pdsch_enc_create(accelerator_config.device, 1);
  if (probe == nullptr) {
    return nullptr;
  }
  srsran_cuda_pdsch_enc_destroy(probe);
  return std::make_shared<hw_accelerator_pdsch_enc_cuda_factory>(accelerator_config);
}

std::shared_ptr<hw_accelerator_pusch_dec_factory>
srsran::hal::create_cuda_pusch_dec_acc_factory(const cuda_hwacc_pusch_dec_configuration& accelerator_config)
{
  try {
    return std::make_shared<hw_accelerator_pusch_dec_cuda_factory>(std::make_shared<cuda_pusch_dec_device>(accelerator_config));
  } catch (const std::exception&) {
    return nullptr;
  }
}
