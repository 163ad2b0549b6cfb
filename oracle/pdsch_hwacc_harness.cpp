// TEST INFRASTRUCTURE - drop-in proof of the PDSCH mirror on the reference's own classes.
//
// Built by oracle/Makefile (target hwacc) into oracle/_ref/pdsch_hwacc_parity from the UNMODIFIED reference sources where
// they lie under /root/reference plus the repo's C++ host adapter:
//
//   reference pdsch_encoder_hw_impl  +  OUR hal::hw_accelerator_pdsch_enc (CUDA, B200), CB mode and TB mode  <- under test
//   reference pdsch_encoder_impl     (ldpc_segmenter_tx + ldpc_encoder AVX2 + ldpc_rate_matcher)             <- the oracle
//
// Both encode the same transport blocks over allocations / modulations / redundancy versions / buffer limits in the manner
// of pdsch_encoder_test.cpp and must produce the same code word, bit for bit. Then a throughput figure in the shape of
// pdsch_encoder_hwacc_benchmark.cpp (one encoder instance, TB after TB through the hal seam). Exit code 0 = parity, 1 =
// mismatch, 2 = no CUDA device.
#include "hw_accelerator_factories_cuda.h"
#include "pdsch_encoder_hw_impl.h"
#include "pdsch_encoder_impl.h"
#include "srsran/phy/upper/channel_coding/channel_coding_factories.h"
#include "srsran/ran/sch/tbs_calculator.h"
#include <chrono>
#include <cstdio>
#include <random>
#include <vector>

using namespace srsran;

namespace {

std::unique_ptr<pdsch_encoder> make_reference_encoder()
{
  auto crc_f = create_crc_calculator_factory_sw("auto");
  auto seg_f = create_ldpc_segmenter_tx_factory_sw(crc_f);
  auto enc_f = create_ldpc_encoder_factory_sw("auto");
  auto rm_f  = create_ldpc_rate_matcher_factory_sw();
  return std::make_unique<pdsch_encoder_impl>(seg_f->create(), enc_f->create(), rm_f->create());
}

std::unique_ptr<pdsch_encoder> make_cuda_encoder(bool cb_mode, unsigned max_tb_size)
{
  hal::cuda_hwacc_pdsch_enc_configuration cfg;
  cfg.cb_mode     = cb_mode;
  cfg.max_tb_size = max_tb_size;
  auto acc_f      = hal::create_cuda_pdsch_enc_acc_factory(cfg);
  if (!acc_f) {
    return nullptr;
  }
  auto                           crc_f = create_crc_calculator_factory_sw("auto");
  auto                           seg_f = create_ldpc_segmenter_tx_factory_sw(crc_f);
  pdsch_encoder_hw_impl::sch_crc crcs  = {crc_f->create(crc_generator_poly::CRC16),
                                          crc_f->create(crc_generator_poly::CRC24A),
                                          crc_f->create(crc_generator_poly::CRC24B)};
  return std::make_unique<pdsch_encoder_hw_impl>(crcs, seg_f->create(), acc_f->create());
}

struct tb_case {
  unsigned prb, qm, rate, layers, bg, nref, rv;
};

unsigned tbs_of(const tb_case& c)
{
  tbs_calculator_configuration t = {};
  t.nof_symb_sh                  = 14;
  t.nof_dmrs_prb                 = 12;
  t.nof_oh_prb                   = 0;
  t.mcs_descr.modulation         = static_cast<modulation_scheme>(c.qm);
  t.mcs_descr.target_code_rate   = static_cast<float>(c.rate);
  t.nof_layers                   = c.layers;
  t.tb_scaling_field             = 0;
  t.n_prb                        = c.prb;
  return tbs_calculator_calculate(t);
}

} // namespace

int main(int argc, char** argv)
{
  auto ref = make_reference_encoder();
  auto cb  = make_cuda_encoder(true, 1U << 20);
  auto tb  = make_cuda_encoder(false, 1U << 20);
  auto tbf = make_cuda_encoder(false, 4000); // TB mode with a small limit: larger TBs are forced into CB mode
  if (!cb || !tb || !tbf) {
    std::printf("pdsch_hwacc_parity: no usable CUDA device\n");
    return 2;
  }
  const tb_case cases[] = {{52, 4, 658, 1, 1, 0, 0},     {25, 2, 120, 1, 2, 0, 2},     {24, 8, 948, 2, 1, 12611, 3},
                           {106, 6, 873, 2, 1, 0, 1},    {52, 2, 449, 1, 1, 25344, 0}, {10, 4, 490, 1, 2, 0, 3},
                           {4, 2, 308, 1, 2, 0, 0},      {1, 2, 120, 1, 2, 0, 1},      {273, 2, 308, 1, 1, 0, 2},
                           {273, 2, 193, 2, 2, 0, 0},    {133, 8, 948, 3, 1, 12611, 0}, {51, 6, 567, 3, 1, 9000, 2},
                           {273, 8, 948, 2, 1, 25223, 0}, {273, 8, 948, 3, 1, 16815, 1}, {200, 6, 666, 4, 1, 0, 3}};
  std::mt19937  rng(99);
  unsigned      checked = 0, bad = 0;
  for (const tb_case& c : cases) {
    const unsigned             tbs   = tbs_of(c);
    const unsigned             nbits = c.prb * 156 * c.qm * c.layers;
    std::vector<uint8_t>       data(tbs / 8), want(nbits), got(nbits);
    pdsch_encoder::configuration cfg;
    cfg.base_graph     = (c.bg == 1) ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
    cfg.rv             = c.rv;
    cfg.mod            = static_cast<modulation_scheme>(c.qm);
    cfg.Nref           = c.nref;
    cfg.nof_layers     = c.layers;
    cfg.nof_ch_symbols = nbits / c.qm;
    for (unsigned rep = 0; rep != 2; ++rep) {
      for (uint8_t& b : data) {
        b = static_cast<uint8_t>(rng());
      }
      ref->encode(want, data, cfg);
      pdsch_encoder* dut[3]  = {cb.get(), tb.get(), tbf.get()};
      const char*    name[3] = {"cb-mode", "tb-mode", "tb-mode(forced cb)"};
      for (unsigned k = 0; k != 3; ++k) {
        std::fill(got.begin(), got.end(), 0xAA);
        dut[k]->encode(got, data, cfg);
        ++checked;
        if (got != want) {
          ++bad;
          size_t first = 0;
          while (first != nbits && got[first] == want[first]) {
            ++first;
          }
          std::printf("MISMATCH %s prb=%u qm=%u R=%u layers=%u bg=%u nref=%u rv=%u tbs=%u first bit %zu\n", name[k], c.prb, c.qm,
                      c.rate, c.layers, c.bg, c.nref, c.rv, tbs, first);
        }
      }
    }
  }
  std::printf("pdsch_hwacc_parity: %u code words compared with the reference's pdsch_encoder_impl, %u mismatches\n", checked, bad);

  // Throughput through the reference's pdsch_encoder interface, one instance, TB after TB (pdsch_encoder_hwacc_benchmark.cpp).
  {
    const tb_case  c     = {273, 8, 948, 2, 1, 25223, 0};
    const unsigned tbs   = tbs_of(c);
    const unsigned nbits = c.prb * 156 * c.qm * c.layers;
    std::vector<uint8_t> data(tbs / 8, 0x5b), out(nbits);
    pdsch_encoder::configuration cfg;
    cfg.base_graph     = ldpc_base_graph_type::BG1;
    cfg.rv             = 0;
    cfg.mod            = modulation_scheme::QAM256;
    cfg.Nref           = c.nref;
    cfg.nof_layers     = c.layers;
    cfg.nof_ch_symbols = nbits / c.qm;
    const unsigned reps = (argc > 1) ? static_cast<unsigned>(std::atoi(argv[1])) : 200;
    pdsch_encoder* who[3]  = {ref.get(), cb.get(), tb.get()};
    const char*    name[3] = {"reference pdsch_encoder_impl (1 thread)", "pdsch_encoder_hw_impl + cuda accelerator, CB mode",
                              "pdsch_encoder_hw_impl + cuda accelerator, TB mode"};
    for (unsigned k = 0; k != 3; ++k) {
      who[k]->encode(out, data, cfg);
      auto t0 = std::chrono::steady_clock::now();
      for (unsigned r = 0; r != reps; ++r) {
        who[k]->encode(out, data, cfg);
      }
      double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      std::printf("{\"encoder\": \"%s\", \"tbs_bits\": %u, \"codeword_bits\": %u, \"us_per_tb\": %.1f, \"info_gbit_per_s\": %.3f}\n",
                  name[k], tbs, nbits, s / reps * 1e6, static_cast<double>(tbs) * reps / s / 1e9);
    }
  }
  return bad == 0 ? 0 : 1;
}
