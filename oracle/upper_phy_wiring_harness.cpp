// TEST INFRASTRUCTURE - SURVEY.md 8(f) row 3 as compiled code: what the patched upper PHY factory
// (integration/0001-pusch-decoder-type-cuda.patch, lib/phy/upper/upper_phy_factories.cpp:395-445,626) does for the PUSCH
// decoding path, executed on the reference's own classes:
//   * the PUSCH decoder factory is chosen by the configuration STRING ("sw" | "cuda"), exactly the branch of the patch;
//   * the receive buffer pool is the reference's rx_buffer_pool_impl (lib/phy/upper/rx_buffer_pool_impl.cpp:36-142) created
//     with external_soft_bits = true for "cuda": its buffers hand out EMPTY soft-bit spans (rx_buffer_impl.h), the soft bits
//     live in the GPU's HARQ slots addressed by the pool's absolute code-block identifiers;
//   * HARQ sequences run through reserve() / unique_rx_buffer lock - unlock - release, several UEs interleaved, against a
//     second pool (internal soft bits) driving the reference's software decoder: same CRC verdicts, same transport blocks,
//     same iteration statistics, same CRC flags in both pools;
//   * expiry: a HARQ process whose retransmission never comes is expired by run_slot(); its code-block identifiers go back
//     to the pool and are handed to another UE, whose NEW transmission decodes correctly in the reused GPU slots
//     (free_harq_context_entry is a no-op by design: new-data dematching overwrites, like the pool never clears).
// Built by oracle/Makefile (target hwacc) into oracle/_ref/upper_phy_wiring. Exit 0 = parity, 1 = mismatch, 2 = no CUDA device.
#include "pdsch_encoder_impl.h"
#include "pusch_decoder_cuda_impl.h"
#include "pusch_decoder_impl.h"
#include "srsran/phy/upper/channel_coding/channel_coding_factories.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_notifier.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_result.h"
#include "srsran/phy/upper/rx_buffer_pool.h"
#include "srsran/phy/upper/unique_rx_buffer.h"
#include <cmath>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

using namespace srsran;

namespace {

/// The software decoder factory (what create_pusch_decoder_factory_sw builds; constructed directly so that the harness does
/// not link the whole PUSCH processor).
class sw_decoder_factory : public pusch_decoder_factory
{
public:
  std::unique_ptr<pusch_decoder> create() override
  {
    auto crc_factory = create_crc_calculator_factory_sw("auto");
    auto dec_factory = create_ldpc_decoder_factory_sw("auto");
    auto dem_factory = create_ldpc_rate_dematcher_factory_sw("auto");
    auto seg_factory = create_ldpc_segmenter_rx_factory_sw();
    std::vector<std::unique_ptr<pusch_codeblock_decoder>> cbd(1);
    pusch_codeblock_decoder::sch_crc                      crcs;
    crcs.crc16  = crc_factory->create(crc_generator_poly::CRC16);
    crcs.crc24A = crc_factory->create(crc_generator_poly::CRC24A);
    crcs.crc24B = crc_factory->create(crc_generator_poly::CRC24B);
    cbd[0]      = std::make_unique<pusch_codeblock_decoder>(dem_factory->create(), dec_factory->create(), crcs);
    auto pool   = std::make_shared<pusch_decoder_impl::codeblock_decoder_pool>(std::move(cbd));
    pusch_decoder_impl::sch_crc sw_crcs;
    sw_crcs.crc16  = crc_factory->create(crc_generator_poly::CRC16);
    sw_crcs.crc24A = crc_factory->create(crc_generator_poly::CRC24A);
    sw_crcs.crc24B = crc_factory->create(crc_generator_poly::CRC24B);
    return std::make_unique<pusch_decoder_impl>(seg_factory->create(), pool, std::move(sw_crcs), nullptr, MAX_RB, 4);
  }
};

/// The branch of the patched upper_phy_factories.cpp: decoder factory from the configuration string.
std::shared_ptr<pusch_decoder_factory> make_decoder_factory(const std::string& type, const rx_buffer_pool_config& pool_cfg)
{
  if (type == "cuda") {
    try {
      hal::cuda_hwacc_pusch_dec_configuration cuda_config;
      cuda_config.device            = 0;
      cuda_config.nof_harq_cb_slots = pool_cfg.nof_codeblocks;
      cuda_config.max_cbs_in_flight = std::min(pool_cfg.nof_codeblocks, 16384U);
      std::vector<std::shared_ptr<hal::cuda_pusch_dec_device>> devices = {std::make_shared<hal::cuda_pusch_dec_device>(cuda_config)};
      devices[0]->set_aggregation(1, std::chrono::microseconds(0)); // one decoder, synchronous: no point in waiting
      return create_pusch_decoder_factory_cuda(devices, nullptr, MAX_RB, 4);
    } catch (const std::exception& e) {
      std::printf("upper_phy_wiring: %s\n", e.what());
      return nullptr;
    }
  }
  return std::make_shared<sw_decoder_factory>();
}

struct notifier_t : public pusch_decoder_notifier {
  void on_sch_data(const pusch_decoder_result& result) override
  {
    res  = result;
    done = true;
  }
  pusch_decoder_result res;
  bool                 done = false;
};

struct ue_t {
  uint16_t             rnti;
  uint8_t              harq_id;
  unsigned             tbs_bits, bg, Qm, nof_layers, nof_llrs, Nref;
  double               mu;
  std::vector<uint8_t> tb;
};

pusch_decoder_result decode(pusch_decoder& dec, unique_rx_buffer buffer, std::vector<uint8_t>& out, const std::vector<int8_t>& llrs,
                            const ue_t& ue, unsigned rv, bool new_data)
{
  pusch_decoder::configuration cfg;
  cfg.base_graph          = (ue.bg == 1) ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
  cfg.rv                  = rv;
  cfg.mod                 = static_cast<modulation_scheme>(ue.Qm);
  cfg.Nref                = ue.Nref;
  cfg.nof_layers          = ue.nof_layers;
  cfg.nof_ldpc_iterations = 6;
  cfg.use_early_stop      = true;
  cfg.new_data            = new_data;
  notifier_t            n;
  pusch_decoder_buffer& buf = dec.new_data(span<uint8_t>(out), std::move(buffer), n, cfg);
  buf.on_new_softbits(span<const log_likelihood_ratio>(reinterpret_cast<const log_likelihood_ratio*>(llrs.data()), llrs.size()));
  buf.on_end_softbits();
  return n.res;
}

bool same_result(const pusch_decoder_result& a, const pusch_decoder_result& b)
{
  bool same = (a.tb_crc_ok == b.tb_crc_ok) && (a.nof_codeblocks_total == b.nof_codeblocks_total) &&
              (a.ldpc_decoder_stats.get_nof_observations() == b.ldpc_decoder_stats.get_nof_observations());
  if (same && a.ldpc_decoder_stats.get_nof_observations() != 0) {
    same = (a.ldpc_decoder_stats.get_min() == b.ldpc_decoder_stats.get_min()) &&
           (a.ldpc_decoder_stats.get_max() == b.ldpc_decoder_stats.get_max()) &&
           (std::abs(a.ldpc_decoder_stats.get_mean() - b.ldpc_decoder_stats.get_mean()) < 1e-4);
  }
  return same;
}

} // namespace

int main()
{
  // Pool dimensions: small on purpose, so that identifiers are recycled and code-block runs get scattered.
  rx_buffer_pool_config pool_cfg;
  pool_cfg.max_codeblock_size   = ldpc::MAX_CODEBLOCK_SIZE;
  pool_cfg.nof_buffers          = 6;
  pool_cfg.nof_codeblocks       = 14;
  pool_cfg.expire_timeout_slots = 8;
  pool_cfg.external_soft_bits   = false;
  rx_buffer_pool_config cuda_pool_cfg = pool_cfg;
  cuda_pool_cfg.external_soft_bits    = true; // what the patched factory does for pusch_decoder_type == "cuda"

  auto cuda_factory = make_decoder_factory("cuda", cuda_pool_cfg);
  if (!cuda_factory) {
    std::printf("upper_phy_wiring: pusch_decoder_type cuda is not available (no usable CUDA device)\n");
    return 2;
  }
  auto sw_factory  = make_decoder_factory("sw", pool_cfg);
  auto cuda_dec    = cuda_factory->create();
  auto sw_dec      = sw_factory->create();
  auto sw_pool_c   = create_rx_buffer_pool(pool_cfg);
  auto cuda_pool_c = create_rx_buffer_pool(cuda_pool_cfg);
  rx_buffer_pool& sw_pool   = sw_pool_c->get_pool();
  rx_buffer_pool& cuda_pool = cuda_pool_c->get_pool();

  auto               crc_factory = create_crc_calculator_factory_sw("auto");
  auto               seg_tx      = create_ldpc_segmenter_tx_factory_sw(crc_factory);
  auto               enc_f       = create_ldpc_encoder_factory_sw("auto");
  auto               rm_f        = create_ldpc_rate_matcher_factory_sw();
  pdsch_encoder_impl encoder(seg_tx->create(), enc_f->create(), rm_f->create());

  std::mt19937 rgen(77);
  auto make_ue = [&](uint16_t rnti, uint8_t harq, unsigned tbs, unsigned bg, unsigned qm, unsigned nllr, double mu) {
    ue_t u = {rnti, harq, tbs, bg, qm, 1, nllr, 25344, mu, std::vector<uint8_t>(tbs / 8)};
    for (uint8_t& b : u.tb) {
      b = static_cast<uint8_t>(rgen());
    }
    return u;
  };
  auto transmit = [&](const ue_t& ue, unsigned rv, std::vector<int8_t>& llrs) {
    std::vector<uint8_t>         cw(ue.nof_llrs);
    pdsch_encoder::configuration ecfg;
    ecfg.base_graph     = (ue.bg == 1) ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
    ecfg.rv             = rv;
    ecfg.mod            = static_cast<modulation_scheme>(ue.Qm);
    ecfg.Nref           = ue.Nref;
    ecfg.nof_layers     = ue.nof_layers;
    ecfg.nof_ch_symbols = ue.nof_llrs / ue.Qm;
    encoder.encode(span<uint8_t>(cw), span<const uint8_t>(ue.tb), ecfg);
    std::normal_distribution<double> noise(0.0, std::sqrt(2.0 * ue.mu));
    llrs.resize(ue.nof_llrs);
    for (unsigned k = 0; k != ue.nof_llrs; ++k) {
      double x = (cw[k] ? -ue.mu : ue.mu) + noise(rgen);
      llrs[k]  = static_cast<int8_t>(std::max(-120.0, std::min(120.0, std::round(4.0 * x))));
    }
  };

  int        failures = 0, nof_ok = 0, nof_recycled = 0;
  std::vector<int> id_owner(pool_cfg.nof_codeblocks, -1); // rnti that last used a code-block identifier
  slot_point slot(1, 0);
  // One step: the same transmission through both pools / decoders, results compared.
  auto step = [&](const ue_t& ue, unsigned rv, bool new_data, const char* what, bool expect_reserved = true) {
    unsigned nof_cbs = ldpc::compute_nof_codeblocks(units::bits(ue.tbs_bits),
                                                    (ue.bg == 1) ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2);
    trx_buffer_identifier id(ue.rnti, ue.harq_id);
    unique_rx_buffer      b_sw   = sw_pool.reserve(slot, id, nof_cbs, new_data);
    unique_rx_buffer      b_cuda = cuda_pool.reserve(slot, id, nof_cbs, new_data);
    if (b_sw.is_valid() != b_cuda.is_valid() || b_sw.is_valid() != expect_reserved) {
      std::printf("%-46s reserve sw/cuda = %d/%d (expected %d) -> MISMATCH\n", what, b_sw.is_valid(), b_cuda.is_valid(), expect_reserved);
      ++failures;
      return pusch_decoder_result{};
    }
    if (!b_sw.is_valid()) {
      std::printf("%-46s reservation refused by both pools (as expected) -> ok\n", what);
      return pusch_decoder_result{};
    }
    // external_soft_bits: the pool's buffer carries no soft-bit storage at all (asking it for even one soft bit asserts,
    // rx_buffer_impl.h:189: "exceeds maximum size 0"), so nothing on the host could be combining.
    bool     no_host_soft = b_cuda->get_codeblock_soft_bits(0, 0).empty();
    unsigned first_id = b_cuda->get_absolute_codeblock_id(0), last_id = b_cuda->get_absolute_codeblock_id(nof_cbs - 1);
    std::vector<int8_t> llrs;
    transmit(ue, rv, llrs);
    std::vector<uint8_t> out_sw(ue.tb.size()), out_cuda(ue.tb.size());
    std::vector<bool>    crc_sw(nof_cbs), crc_cuda(nof_cbs);
    // The buffers are moved into the decoders (which release or unlock them); keep views of the CRC flags first.
    span<bool> f_sw = b_sw->get_codeblocks_crc(), f_cuda = b_cuda->get_codeblocks_crc();
    pusch_decoder_result r_sw   = decode(*sw_dec, std::move(b_sw), out_sw, llrs, ue, rv, new_data);
    pusch_decoder_result r_cuda = decode(*cuda_dec, std::move(b_cuda), out_cuda, llrs, ue, rv, new_data);
    bool same = same_result(r_sw, r_cuda) && no_host_soft;
    for (unsigned i = 0; same && i != nof_cbs; ++i) {
      same = (f_sw[i] == f_cuda[i]);
    }
    if (same && r_sw.tb_crc_ok) {
      same = (out_sw == out_cuda) && (out_cuda == ue.tb);
    }
    bool recycled = false;
    for (unsigned id2 = first_id; id2 <= last_id && id2 < id_owner.size(); ++id2) {
      recycled     = recycled || (id_owner[id2] >= 0 && id_owner[id2] != ue.rnti);
      id_owner[id2] = ue.rnti;
    }
    std::printf("%-46s rnti %#x harq %u rv%u cbs=%u ids=[%u..%u]%s crc sw/cuda=%d/%d -> %s\n", what, ue.rnti, ue.harq_id, rv, nof_cbs,
                first_id, last_id, recycled ? " (recycled from another UE)" : "", r_sw.tb_crc_ok, r_cuda.tb_crc_ok,
                same ? "ok" : "MISMATCH");
    failures += same ? 0 : 1;
    nof_ok += r_cuda.tb_crc_ok ? 1 : 0;
    nof_recycled += recycled ? 1 : 0;
    return r_cuda;
  };
  auto advance = [&](unsigned n) {
    for (unsigned i = 0; i != n; ++i) {
      ++slot;
      sw_pool.run_slot(slot);
      cuda_pool.run_slot(slot);
    }
  };

  // ---- 1. several UEs interleaved, HARQ retransmissions with soft combining in the GPU's slots --------------------------------
  ue_t a = make_ue(0x4601, 0, 21000, 1, 4, 32448, 2.6);  // 3 code blocks, needs a retransmission
  ue_t b = make_ue(0x4602, 3, 1928, 2, 2, 16224, 0.42);  // 1 code block BG2
  ue_t c = make_ue(0x4603, 1, 12040, 1, 4, 32448, 1.7);  // 2 code blocks
  step(a, 0, true, "UE A first transmission");
  advance(1);
  step(b, 0, true, "UE B first transmission");
  step(c, 0, true, "UE C first transmission");
  advance(2);
  // (A HARQ process whose transport block has been decoded is released by the decoder: a further "retransmission" would be
  // refused by the pools, so the sequence below follows the CRC verdicts.)
  bool a_ok = false, b_ok = false, c_ok = false;
  auto retx = [&](const ue_t& ue, bool& done, unsigned rv, const char* what) {
    if (!done) {
      done = step(ue, rv, false, what).tb_crc_ok;
    }
  };
  retx(a, a_ok, 2, "UE A retransmission (combining in HBM)");
  retx(c, c_ok, 2, "UE C retransmission");
  advance(1);
  retx(b, b_ok, 2, "UE B retransmission");
  retx(a, a_ok, 3, "UE A second retransmission");
  retx(c, c_ok, 3, "UE C second retransmission");
  retx(b, b_ok, 3, "UE B second retransmission");
  // ---- 2. a retransmission for a HARQ process the pool does not know is refused by both pools ---------------------------------
  ue_t ghost = make_ue(0x4777, 5, 12040, 1, 4, 32448, 0.7);
  step(ghost, 2, false, "retransmission without a first transmission", false);
  // ---- 3. expiry: UE D fails and never retransmits; its identifiers return to the pool and are reused -------------------------
  ue_t d = make_ue(0x4604, 2, 21000, 1, 4, 32448, 0.20); // hopeless SNR: stays locked until it expires
  step(d, 0, true, "UE D first transmission (fails, abandoned)");
  advance(pool_cfg.expire_timeout_slots + 2);
  step(d, 2, false, "UE D retransmission after expiry", false);
  // New UEs fill the pool again: their code blocks land in the GPU slots UE D (and the released UEs) used before.
  for (unsigned k = 0; k != 5; ++k) {
    ue_t e = make_ue(static_cast<uint16_t>(0x4700 + k), static_cast<uint8_t>(k), (k % 2) ? 21000 : 12040, 1, 4, 32448, 6.0);
    step(e, 0, true, "new UE in recycled code-block slots");
    advance(1);
  }
  if (nof_ok < 4 || nof_recycled < 2) {
    std::printf("scenario too weak: %d transport blocks decoded, %d reuses of another UE's identifiers\n", nof_ok, nof_recycled);
    ++failures;
  }
  std::printf("upper_phy_wiring: %s (%d mismatches; %d transport blocks decoded, %d allocations in recycled code-block slots)\n",
              failures ? "FAILED" : "PASSED", failures, nof_ok, nof_recycled);
  return failures ? 1 : 0;
}
