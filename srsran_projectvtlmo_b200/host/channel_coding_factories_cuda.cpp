#include "channel_coding_factories_cuda.h"
#include "srsran/phy/upper/channel_coding/ldpc/ldpc.h"
#include "srsran/srsvec/bit.h"
#include "srsran/support/srsran_assert.h"
#include <vector>

using namespace srsran;

namespace {

uint32_t to_cuda_poly(crc_generator_poly poly)
{
  switch (poly) {
    case crc_generator_poly::CRC24A:
      return SRSRAN_CUDA_CRC24A;
    case crc_generator_poly::CRC24B:
      return SRSRAN_CUDA_CRC24B;
    case crc_generator_poly::CRC16:
      return SRSRAN_CUDA_CRC16;
    default:
      return SRSRAN_CUDA_CRC_NONE;
  }
}

/// ldpc_decoder (include/srsran/phy/upper/channel_coding/ldpc/ldpc_decoder.h:37-75) on the GPU.
class ldpc_decoder_cuda : public ldpc_decoder
{
public:
  explicit ldpc_decoder_cuda(std::shared_ptr<hal::cuda_pusch_dec_device> device_) : device(std::move(device_)) {}

  std::optional<unsigned>
  decode(bit_buffer& output, span<const log_likelihood_ratio> input, crc_calculator* crc, const configuration& cfg) override
  {
    const auto& tb      = cfg.block_conf.tb_common;
    uint32_t    bg      = (tb.base_graph == ldpc_base_graph_type::BG1) ? 1 : 2;
    uint32_t    Z       = static_cast<uint32_t>(tb.lifting_size);
    uint32_t    poly    = (crc != nullptr) ? to_cuda_poly(crc->get_generator_poly()) : SRSRAN_CUDA_CRC_NONE;
    int         nof_its = -1;
    srsran_assert((crc == nullptr) || (poly != SRSRAN_CUDA_CRC_NONE), "CRC polynomial not used by PUSCH.");
    std::lock_guard<std::recursive_mutex> lock(device->mutex());
    int st = srsran_cuda_ldpc_decode(device->get(),
                                     output.get_buffer().data(),
                                     reinterpret_cast<const int8_t*>(input.data()),
                                     input.size(),
                                     bg,
                                     Z,
                                     cfg.block_conf.cb_specific.nof_filler_bits,
                                     poly,
                                     cfg.algorithm_conf.max_iterations,
                                     cfg.algorithm_conf.scaling_factor,
                                     &nof_its);
    srsran_assert(st == SRSRAN_CUDA_OK, "CUDA LDPC decoder failed ({}): {}", st, srsran_cuda_pusch_dec_last_error(device->get()));
    if (nof_its < 0) {
      return std::nullopt;
    }
    return static_cast<unsigned>(nof_its);
  }

private:
  std::shared_ptr<hal::cuda_pusch_dec_device> device;
};

/// ldpc_rate_dematcher (include/srsran/phy/upper/channel_coding/ldpc/ldpc_rate_dematcher.h:35-56) on the GPU.
class ldpc_rate_dematcher_cuda : public ldpc_rate_dematcher
{
public:
  explicit ldpc_rate_dematcher_cuda(std::shared_ptr<hal::cuda_pusch_dec_device> device_) : device(std::move(device_)) {}

  void rate_dematch(span<log_likelihood_ratio>       output,
                    span<const log_likelihood_ratio> input,
                    bool                             new_data,
                    const codeblock_metadata&        cfg) override
  {
    std::lock_guard<std::recursive_mutex> lock(device->mutex());
    int st = srsran_cuda_ldpc_rate_dematch(device->get(),
                                           reinterpret_cast<int8_t*>(output.data()),
                                           output.size(),
                                           reinterpret_cast<const int8_t*>(input.data()),
                                           input.size(),
                                           new_data ? 1 : 0,
                                           cfg.tb_common.rv,
                                           get_bits_per_symbol(cfg.tb_common.mod),
                                           cfg.tb_common.Nref,
                                           cfg.cb_specific.nof_filler_bits);
    srsran_assert(st == SRSRAN_CUDA_OK, "CUDA rate dematcher failed ({}): {}", st, srsran_cuda_pusch_dec_last_error(device->get()));
  }

private:
  std::shared_ptr<hal::cuda_pusch_dec_device> device;
};

/// crc_calculator (include/srsran/phy/upper/channel_coding/crc_calculator.h:62-84) on the GPU.
class crc_calculator_cuda : public crc_calculator
{
public:
  crc_calculator_cuda(std::shared_ptr<hal::cuda_pusch_dec_device> device_, crc_generator_poly poly_) :
    device(std::move(device_)), poly(poly_)
  {
  }

  crc_calculator_checksum_t calculate_byte(span<const uint8_t> data) override { return run(data.data(), data.size() * 8); }

  crc_calculator_checksum_t calculate_bit(span<const uint8_t> data) override
  {
    packed.assign((data.size() + 7) / 8, 0);
    for (size_t i = 0; i != data.size(); ++i) {
      packed[i / 8] |= static_cast<uint8_t>((data[i] & 1U) << (7 - i % 8));
    }
    return run(packed.data(), data.size());
  }

  crc_calculator_checksum_t calculate(const bit_buffer& data) override
  {
    return run(data.get_buffer().data(), data.size());
  }

  crc_generator_poly get_generator_poly() const override { return poly; }

private:
  crc_calculator_checksum_t run(const uint8_t* bytes, size_t nof_bits)
  {
    uint32_t                    checksum = 0;
    std::lock_guard<std::recursive_mutex> lock(device->mutex());
    int st = srsran_cuda_crc_calculate(device->get(), to_cuda_poly(poly), bytes, nof_bits, &checksum);
    srsran_assert(st == SRSRAN_CUDA_OK, "CUDA CRC calculator failed ({}).", st);
    return checksum;
  }

  std::shared_ptr<hal::cuda_pusch_dec_device> device;
  crc_generator_poly                          poly;
  std::vector<uint8_t>                        packed;
};

template <typename Interface, typename Impl>
class factory_cuda : public Interface
{
public:
  explicit factory_cuda(std::shared_ptr<hal::cuda_pusch_dec_device> device_) : device(std::move(device_)) {}
  auto create() -> decltype(std::declval<Interface>().create()) override { return std::make_unique<Impl>(device); }

private:
  std::shared_ptr<hal::cuda_pusch_dec_device> device;
};

class crc_calculator_factory_cuda : public crc_calculator_factory
{
public:
  explicit crc_calculator_factory_cuda(std::shared_ptr<hal::cuda_pusch_dec_device> device_) : device(std::move(device_)) {}
  std::unique_ptr<crc_calculator> create(crc_generator_poly poly) override
  {
    if (to_cuda_poly(poly) == SRSRAN_CUDA_CRC_NONE) {
      return nullptr;
    }
    return std::make_unique<crc_calculator_cuda>(device, poly);
  }

private:
  std::shared_ptr<hal::cuda_pusch_dec_device> device;
};

} // namespace

std::shared_ptr<ldpc_decoder_factory> srsran::create_ldpc_decoder_factory_cuda(std::shared_ptr<hal::cuda_pusch_dec_device> device)
{
  return device ? std::make_shared<factory_cuda<ldpc_decoder_factory, ldpc_decoder_cuda>>(std::move(device)) : nullptr;
}

std::shared_ptr<ldpc_rate_dematcher_factory>
srsran::create_ldpc_rate_dematcher_factory_cuda(std::shared_ptr<hal::cuda_pusch_dec_device> device)
{
  return device ? std::make_shared<factory_cuda<ldpc_rate_dematcher_factory, ldpc_rate_dematcher_cuda>>(std::move(device))
                : nullptr;
}

std::shared_ptr<crc_calculator_factory> srsran::create_crc_calculator_factory_cuda(std::shared_ptr<hal::cuda_pusch_dec_device> device)
{
  return device ? std::make_shared<crc_calculator_factory_cuda>(std::move(device)) : nullptr;
}
