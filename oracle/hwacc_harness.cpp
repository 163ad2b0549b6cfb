// TEST INFRASTRUCTURE - drop-in proof on the reference's own classes.
//
// Builds (oracle/Makefile, target hwacc) into oracle/_ref/hwacc_parity from the UNMODIFIED reference sources where they
// lie under /root/reference plus the repo's C++ host adapters (srsran_projectvtlmo_b200/host/):
//
//   reference pusch_decoder_hw_impl  +  OUR hal::hw_accelerator_pusch_dec (CUDA, B200)      <- device under test
//   OUR pusch_decoder_cuda_impl (the reference's pusch_decoder interface, streamed soft bits)  <- device under test
//   reference pusch_decoder_impl     +  reference AVX-512/AVX2 ldpc_decoder / rate dematcher <- the oracle
//
// Both decode the same transport blocks (reference pdsch_encoder_impl as the transmitter, AWGN, int8 LLRs) over the HARQ
// sequence rv 0,2,3,1 like pusch_decoder_vectortest.cpp:261-397 and must agree on TB CRC, TB bytes, number of code blocks
// and LDPC iteration statistics. Exit code 0 = parity, 1 = mismatch, 2 = no CUDA device.
#include "hw_accelerator_factories_cuda.h"
#include "channel_coding_factories_cuda.h"
#include "pdsch_encoder_impl.h"
#include "pusch_decoder_cuda_impl.h"
#include "pusch_decoder_hw_impl.h"
#include "pusch_decoder_impl.h"
#include "srsran/phy/upper/channel_coding/channel_coding_factories.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_notifier.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_result.h"
#include "srsran/phy/upper/unique_rx_buffer.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

using namespace srsran;

namespace {

class harness_rx_buffer : public unique_rx_buffer::callback
{
public:
  harness_rx_buffer(unsigned nof_cbs_, unsigned first_absolute_id_, unsigned id_stride_ = 1) :
    nof_cbs(nof_cbs_),
    first_absolute_id(first_absolute_id_),
    id_stride(id_stride_),
    crcs(new bool[nof_cbs_]()),
    soft(nof_cbs_),
    data(nof_cbs_)
  {
    for (unsigned i = 0; i != nof_cbs; ++i) {
      soft[i].assign(ldpc::MAX_CODEBLOCK_SIZE, log_likelihood_ratio(0));
      data[i].assign(ldpc::MAX_CODEBLOCK_SIZE / 8 + 8, 0);
    }
  }
  unsigned   get_nof_codeblocks() const override { return nof_cbs; }
  void       reset_codeblocks_crc() override { std::fill(crcs.get(), crcs.get() + nof_cbs, false); }
  span<bool> get_codeblocks_crc() override { return span<bool>(crcs.get(), nof_cbs); }
  // The pool hands out identifiers from a free list: they are not consecutive in general (id_stride > 1 mimics that).
  unsigned get_absolute_codeblock_id(unsigned codeblock_id) const override
  {
    return first_absolute_id + codeblock_id * id_stride;
  }
  span<log_likelihood_ratio> get_codeblock_soft_bits(unsigned codeblock_id, unsigned codeblock_size) override
  {
    return span<log_likelihood_ratio>(soft[codeblock_id]).first(codeblock_size);
  }
  bit_buffer get_codeblock_data_bits(unsigned codeblock_id, unsigned data_size) override
  {
    return bit_buffer::from_bytes(span<uint8_t>(data[codeblock_id])).first(data_size);
  }
  void lock() override {}
  void unlock() override {}
  void release() override {}

  unsigned                                       nof_cbs, first_absolute_id, id_stride;
  std::unique_ptr<bool[]>                        crcs;
  std::vector<std::vector<log_likelihood_ratio>> soft;
  std::vector<std::vector<uint8_t>>              data;
};

class notifier_t : public pusch_decoder_notifier
{
public:
  void on_sch_data(const pusch_decoder_result& result) override
  {
    res  = result;
    done = true;
  }
  pusch_decoder_result res;
  bool                 done = false;
};

struct tb_case {
  const char* name;
  unsigned    tbs_bits, bg, Qm, nof_layers, nof_llrs, Nref;
  double      mu;
  bool        early_stop;
  unsigned    max_it;
};

pusch_decoder_result run(pusch_decoder&                decoder,
                         harness_rx_buffer&            buffer,
                         std::vector<uint8_t>&         tb_out,
                         const std::vector<int8_t>&    llrs,
                         const tb_case&                c,
                         unsigned                      rv,
                         bool                          new_data,
                         unsigned                      feed_mode = 0)
{
  if (new_data) {
    buffer.reset_codeblocks_crc(); // rx_buffer_pool_impl::reserve does this for new data
  }
  pusch_decoder::configuration cfg;
  cfg.base_graph          = (c.bg == 1) ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
  cfg.rv                  = rv;
  cfg.mod                 = static_cast<modulation_scheme>(c.Qm);
  cfg.Nref                = c.Nref;
  cfg.nof_layers          = c.nof_layers;
  cfg.nof_ldpc_iterations = c.max_it;
  cfg.use_early_stop      = c.early_stop;
  cfg.new_data            = new_data;
  notifier_t            notifier;
  pusch_decoder_buffer& buf = decoder.new_data(span<uint8_t>(tb_out), unique_rx_buffer(buffer), notifier, cfg);
  const log_likelihood_ratio* src = reinterpret_cast<const log_likelihood_ratio*>(llrs.data());
  if (feed_mode == 0) {
    buf.on_new_softbits(span<const log_likelihood_ratio>(src, llrs.size()));
  } else {
    // Block by block like the UL-SCH demultiplexer (ulsch_demultiplex_impl.cpp): alternately written through the
    // decoder's view and passed by pointer; feed_mode 1 announces the total first (pusch_decoder::set_nof_softbits).
    if (feed_mode == 1) {
      decoder.set_nof_softbits(units::bits(llrs.size()));
    }
    const size_t block = 20000 + 17 * rv;
    unsigned     k     = 0;
    for (size_t off = 0; off < llrs.size(); off += block, ++k) {
      size_t n = std::min(block, llrs.size() - off);
      if (k % 2 == 0) {
        span<log_likelihood_ratio> view = buf.get_next_block_view(n);
        std::copy(src + off, src + off + n, view.begin());
        buf.on_new_softbits(view);
      } else {
        buf.on_new_softbits(span<const log_likelihood_ratio>(src + off, n));
      }
    }
  }
  buf.on_end_softbits();
  return notifier.res;
}

} // namespace

int main()
{
  // ---- device under test: reference pusch_decoder_hw_impl over the CUDA accelerator ------------------------------------
  hal::cuda_hwacc_pusch_dec_configuration acc_cfg;
  acc_cfg.device            = 0;
  acc_cfg.max_cbs_in_flight = 256;
  acc_cfg.nof_harq_cb_slots = 4096;
  auto hw_factory           = hal::create_cuda_pusch_dec_acc_factory(acc_cfg);
  if (!hw_factory) {
    std::printf("hwacc_parity: no usable CUDA device (%s)\n", srsran_cuda_pusch_dec_last_error(nullptr));
    return 2;
  }
  auto crc_factory = create_crc_calculator_factory_sw("auto");
  auto seg_factory = create_ldpc_segmenter_rx_factory_sw();

  pusch_decoder_hw_impl::sch_crc hw_crcs = {crc_factory->create(crc_generator_poly::CRC16),
                                            crc_factory->create(crc_generator_poly::CRC24A),
                                            crc_factory->create(crc_generator_poly::CRC24B)};
  pusch_decoder_hw_impl          hw_decoder(seg_factory->create(), hw_crcs, hw_factory->create());

  // ---- device under test: the pusch_decoder interface implemented directly on the device (own device context, HARQ slot
  //      ids with a stride of 3 so that the code blocks of a TB do not sit in consecutive slots) --------------------------------
  hal::cuda_hwacc_pusch_dec_configuration cu_cfg = acc_cfg;
  cu_cfg.nof_harq_cb_slots                       = 8192;
  // Two device contexts (on one GPU here; one per GPU in a multi-GPU host): transport blocks are sharded by the rx buffer's
  // first code-block id, and a HARQ retransmission only decodes if it lands on the context that holds its soft bits.
  std::vector<std::shared_ptr<hal::cuda_pusch_dec_device>> cu_devices = {
      std::make_shared<hal::cuda_pusch_dec_device>(cu_cfg), std::make_shared<hal::cuda_pusch_dec_device>(cu_cfg)};
  auto cu_factory = create_pusch_decoder_factory_cuda(cu_devices, nullptr, MAX_RB, 4);
  std::unique_ptr<pusch_decoder> cu_decoder = cu_factory->create();

  // ---- oracle: reference software decoder ---------------------------------------------------------------------------------
  auto dec_factory = create_ldpc_decoder_factory_sw("auto");
  auto dem_factory = create_ldpc_rate_dematcher_factory_sw("auto");
  std::vector<std::unique_ptr<pusch_codeblock_decoder>> cb_decoders(1);
  {
    pusch_codeblock_decoder::sch_crc crcs;
    crcs.crc16     = crc_factory->create(crc_generator_poly::CRC16);
    crcs.crc24A    = crc_factory->create(crc_generator_poly::CRC24A);
    crcs.crc24B    = crc_factory->create(crc_generator_poly::CRC24B);
    cb_decoders[0] = std::make_unique<pusch_codeblock_decoder>(dem_factory->create(), dec_factory->create(), crcs);
  }
  auto pool = std::make_shared<pusch_decoder_impl::codeblock_decoder_pool>(std::move(cb_decoders));
  pusch_decoder_impl::sch_crc sw_crcs;
  sw_crcs.crc16  = crc_factory->create(crc_generator_poly::CRC16);
  sw_crcs.crc24A = crc_factory->create(crc_generator_poly::CRC24A);
  sw_crcs.crc24B = crc_factory->create(crc_generator_poly::CRC24B);
  pusch_decoder_impl sw_decoder(seg_factory->create(), pool, std::move(sw_crcs), nullptr, MAX_RB, 4);

  // ---- transmitter --------------------------------------------------------------------------------------------------------
  auto               seg_tx = create_ldpc_segmenter_tx_factory_sw(crc_factory);
  auto               enc_f  = create_ldpc_encoder_factory_sw("auto");
  auto               rm_f   = create_ldpc_rate_matcher_factory_sw();
  pdsch_encoder_impl encoder(seg_tx->create(), enc_f->create(), rm_f->create());

  // TBS / number of LLRs of SURVEY.md section 8 (156 RE per PRB): name, TBS, BG, Qm, layers, LLRs, Nref, mu, early stop, its
  const tb_case cases[] = {
      {"52prb_16qam_r658", 21000, 1, 4, 1, 32448, 25344, 1.15, true, 6},
      {"52prb_qpsk_r120_bg2", 1928, 2, 2, 1, 16224, 25344, 0.35, true, 6},
      {"25prb_qpsk_r120_bg2", 928, 2, 2, 1, 7800, 25344, 0.5, true, 6},
      {"52prb_16qam_r378_noearlystop", 12040, 1, 4, 1, 32448, 25344, 0.8, false, 3},
      {"273prb_256qam_r948_2layer", 638984, 1, 8, 2, 681408, 25223, 9.0, true, 6},
      {"273prb_256qam_r948_2layer_highsnr", 638984, 1, 8, 2, 681408, 25223, 18.0, true, 6},
  };
  std::mt19937 rgen(2026);
  int          failures = 0;
  unsigned     next_abs = 0, case_no = 0; // even / odd first code-block ids alternate between the two device contexts
  for (const tb_case& c : cases) {
    unsigned nof_cbs = ldpc::compute_nof_codeblocks(units::bits(c.tbs_bits),
                                                    (c.bg == 1) ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2);
    harness_rx_buffer sw_buffer(nof_cbs, 0), hw_buffer(nof_cbs, next_abs), cu_buffer(nof_cbs, 6 * next_abs + 2 + (case_no++ & 1), 3);
    next_abs += nof_cbs;
    std::vector<uint8_t> tb(c.tbs_bits / 8), tb_sw(tb.size()), tb_hw(tb.size()), tb_cu(tb.size()), cw(c.nof_llrs);
    for (uint8_t& b : tb) {
      b = static_cast<uint8_t>(rgen());
    }
    std::vector<int8_t>              llrs(c.nof_llrs);
    std::normal_distribution<double> noise(0.0, std::sqrt(2.0 * c.mu));
    const unsigned                   rvs[4] = {0, 2, 3, 1};
    for (unsigned i = 0; i != 4; ++i) {
      pdsch_encoder::configuration ecfg;
      ecfg.base_graph     = (c.bg == 1) ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
      ecfg.rv             = rvs[i];
      ecfg.mod            = static_cast<modulation_scheme>(c.Qm);
      ecfg.Nref           = c.Nref;
      ecfg.nof_layers     = c.nof_layers;
      ecfg.nof_ch_symbols = c.nof_llrs / c.Qm;
      encoder.encode(span<uint8_t>(cw), span<const uint8_t>(tb), ecfg);
      for (unsigned k = 0; k != c.nof_llrs; ++k) {
        double x = (cw[k] ? -c.mu : c.mu) + noise(rgen);
        llrs[k]  = static_cast<int8_t>(std::max(-120.0, std::min(120.0, std::round(4.0 * x))));
      }
      pusch_decoder_result r_sw = run(sw_decoder, sw_buffer, tb_sw, llrs, c, rvs[i], i == 0);
      pusch_decoder_result r_hw = run(hw_decoder, hw_buffer, tb_hw, llrs, c, rvs[i], i == 0);
      pusch_decoder_result r_cu = run(*cu_decoder, cu_buffer, tb_cu, llrs, c, rvs[i], i == 0, 1 + (i % 2));
      bool same_cu = (r_sw.tb_crc_ok == r_cu.tb_crc_ok) && (r_sw.nof_codeblocks_total == r_cu.nof_codeblocks_total) &&
                     (r_sw.ldpc_decoder_stats.get_nof_observations() == r_cu.ldpc_decoder_stats.get_nof_observations());
      if (same_cu && r_sw.ldpc_decoder_stats.get_nof_observations() != 0) {
        same_cu = (r_sw.ldpc_decoder_stats.get_min() == r_cu.ldpc_decoder_stats.get_min()) &&
                  (r_sw.ldpc_decoder_stats.get_max() == r_cu.ldpc_decoder_stats.get_max()) &&
                  (std::abs(r_sw.ldpc_decoder_stats.get_mean() - r_cu.ldpc_decoder_stats.get_mean()) < 1e-4);
      }
      if (same_cu && r_sw.tb_crc_ok) {
        same_cu = (tb_sw == tb_cu);
      }
      for (unsigned cb = 0; same_cu && cb != nof_cbs; ++cb) {
        same_cu = (sw_buffer.crcs[cb] == cu_buffer.crcs[cb]);
      }
      std::printf("%-36s rv%u: pusch_decoder_cuda_impl (%s) crc=%d obs=%u -> %s\n", c.name, rvs[i],
                  (i % 2 == 0) ? "streamed" : "blocks, total unknown", r_cu.tb_crc_ok,
                  (unsigned)r_cu.ldpc_decoder_stats.get_nof_observations(), same_cu ? "ok" : "MISMATCH");
      failures += same_cu ? 0 : 1;
      bool same = (r_sw.tb_crc_ok == r_hw.tb_crc_ok) && (r_sw.nof_codeblocks_total == r_hw.nof_codeblocks_total) &&
                  (r_sw.ldpc_decoder_stats.get_nof_observations() == r_hw.ldpc_decoder_stats.get_nof_observations());
      if (same && r_sw.ldpc_decoder_stats.get_nof_observations() != 0) {
        same = (r_sw.ldpc_decoder_stats.get_min() == r_hw.ldpc_decoder_stats.get_min()) &&
               (r_sw.ldpc_decoder_stats.get_max() == r_hw.ldpc_decoder_stats.get_max()) &&
               (std::abs(r_sw.ldpc_decoder_stats.get_mean() - r_hw.ldpc_decoder_stats.get_mean()) < 1e-4);
      }
      if (same && r_sw.tb_crc_ok) {
        same = (tb_sw == tb_hw) && (tb_hw == tb);
      }
      // Code-block CRC flags as the two rx buffers hold them.
      for (unsigned cb = 0; same && cb != nof_cbs; ++cb) {
        same = (sw_buffer.crcs[cb] == hw_buffer.crcs[cb]);
      }
      std::printf("%-36s rv%u: cbs=%u crc sw/hw=%d/%d obs=%u/%u it[min,max]=[%u,%u]/[%u,%u] -> %s\n", c.name, rvs[i], nof_cbs,
                  r_sw.tb_crc_ok, r_hw.tb_crc_ok, (unsigned)r_sw.ldpc_decoder_stats.get_nof_observations(),
                  (unsigned)r_hw.ldpc_decoder_stats.get_nof_observations(),
                  r_sw.ldpc_decoder_stats.get_nof_observations() ? r_sw.ldpc_decoder_stats.get_min() : 0,
                  r_sw.ldpc_decoder_stats.get_nof_observations() ? r_sw.ldpc_decoder_stats.get_max() : 0,
                  r_hw.ldpc_decoder_stats.get_nof_observations() ? r_hw.ldpc_decoder_stats.get_min() : 0,
                  r_hw.ldpc_decoder_stats.get_nof_observations() ? r_hw.ldpc_decoder_stats.get_max() : 0, same ? "ok" : "MISMATCH");
      failures += same ? 0 : 1;
      if (r_sw.tb_crc_ok) {
        break;
      }
    }
  }

  // ---- unit-level factories: "cuda" ldpc_decoder / rate dematcher / crc_calculator against the reference's "auto" ---------
  {
    auto device  = std::make_shared<hal::cuda_pusch_dec_device>(acc_cfg);
    auto crc_gpu = create_crc_calculator_factory_cuda(device);
    auto dem_gpu = create_ldpc_rate_dematcher_factory_cuda(device)->create();
    auto dec_gpu = create_ldpc_decoder_factory_cuda(device)->create();
    auto dem_cpu = dem_factory->create();
    auto dec_cpu = dec_factory->create();
    std::uniform_int_distribution<int> llr_dist(-120, 120);
    for (crc_generator_poly poly : {crc_generator_poly::CRC24A, crc_generator_poly::CRC24B, crc_generator_poly::CRC16}) {
      std::vector<uint8_t> msg(997);
      for (uint8_t& b : msg) {
        b = static_cast<uint8_t>(rgen());
      }
      bool same = crc_gpu->create(poly)->calculate_byte(msg) == crc_factory->create(poly)->calculate_byte(msg);
      std::printf("crc_calculator cuda vs auto, poly %d: %s\n", (int)poly, same ? "ok" : "MISMATCH");
      failures += same ? 0 : 1;
    }
    // One BG1 Z=384 code block: dematch (rv 0 then rv 2 combining) and decode with CRC24B early stop.
    codeblock_metadata meta;
    meta.tb_common.base_graph        = ldpc_base_graph_type::BG1;
    meta.tb_common.lifting_size      = ldpc::LS384;
    meta.tb_common.mod               = modulation_scheme::QAM256;
    meta.tb_common.Nref              = 12611;
    meta.cb_specific.nof_filler_bits = 16;
    meta.cb_specific.nof_crc_bits    = 24;
    std::vector<log_likelihood_ratio> soft_cpu(25344, 0), soft_gpu(25344, 0), in(8960);
    bool                              same = true;
    for (unsigned rv : {0U, 2U}) {
      meta.tb_common.rv = rv;
      for (auto& v : in) {
        v = llr_dist(rgen);
      }
      dem_cpu->rate_dematch(soft_cpu, in, rv == 0, meta);
      dem_gpu->rate_dematch(soft_gpu, in, rv == 0, meta);
      same = same && std::equal(soft_cpu.begin(), soft_cpu.end(), soft_gpu.begin());
    }
    std::printf("ldpc_rate_dematcher cuda vs auto: %s\n", same ? "ok" : "MISMATCH");
    failures += same ? 0 : 1;
    std::vector<uint8_t>    out_cpu(1056, 0), out_gpu(1056, 0);
    bit_buffer              bb_cpu = bit_buffer::from_bytes(out_cpu), bb_gpu = bit_buffer::from_bytes(out_gpu);
    ldpc_decoder::configuration dcfg;
    dcfg.block_conf = meta;
    auto crc24b     = crc_factory->create(crc_generator_poly::CRC24B);
    auto it_cpu     = dec_cpu->decode(bb_cpu, soft_cpu, crc24b.get(), dcfg);
    auto it_gpu     = dec_gpu->decode(bb_gpu, soft_gpu, crc24b.get(), dcfg);
    same            = (it_cpu == it_gpu) && (out_cpu == out_gpu);
    std::printf("ldpc_decoder cuda vs auto: %s\n", same ? "ok" : "MISMATCH");
    failures += same ? 0 : 1;
  }
  // ---- the same three "cuda" unit factories over a sweep of shapes: both base graphs, lifting sizes of every kernel class
  // (general, one code block per CTA, packed groups), every modulation order, limited buffers, fillers, rv sequences with
  // combining, valid code words (reference encoder + rate matcher, AWGN) so that early stop and CRC verdicts are exercised,
  // with and without a CRC calculator; CRC sizes of crc_calculator_test.cpp incl. bit lengths that are no byte multiple. -----
  {
    auto device  = std::make_shared<hal::cuda_pusch_dec_device>(acc_cfg);
    auto crc_gpu = create_crc_calculator_factory_cuda(device);
    auto dem_gpu = create_ldpc_rate_dematcher_factory_cuda(device)->create();
    auto dec_gpu = create_ldpc_decoder_factory_cuda(device)->create();
    auto dem_cpu = dem_factory->create();
    auto dec_cpu = dec_factory->create();
    auto enc     = create_ldpc_encoder_factory_sw("auto")->create();
    auto rm      = create_ldpc_rate_matcher_factory_sw()->create();
    unsigned crc_checked = 0, crc_bad = 0;
    for (crc_generator_poly poly : {crc_generator_poly::CRC24A, crc_generator_poly::CRC24B, crc_generator_poly::CRC16}) {
      auto g = crc_gpu->create(poly);
      auto c = crc_factory->create(poly);
      for (unsigned nbytes : {1U, 2U, 3U, 8U, 16U, 32U, 257U, 997U, 6012U}) {
        std::vector<uint8_t> msg(nbytes);
        for (uint8_t& b : msg) {
          b = static_cast<uint8_t>(rgen());
        }
        ++crc_checked;
        crc_bad += g->calculate_byte(msg) == c->calculate_byte(msg) ? 0 : 1;
        std::vector<uint8_t> bits(nbytes * 8 - 3); // not a byte multiple
        for (uint8_t& b : bits) {
          b = static_cast<uint8_t>(rgen() & 1U);
        }
        ++crc_checked;
        crc_bad += g->calculate_bit(bits) == c->calculate_bit(bits) ? 0 : 1;
      }
    }
    std::printf("crc_calculator cuda vs auto, %u messages: %u mismatches\n", crc_checked, crc_bad);
    failures += crc_bad;

    struct shape {
      unsigned bg, Z, Qm, nref_div, nfill, e_num, e_den; // Nref = N * nref_div / 6 (0: unlimited), E = N * e_num / e_den
      float    amp;
    };
    const shape shapes[] = {{1, 384, 8, 0, 16, 1, 2, 9.0F},  {1, 384, 8, 3, 16, 1, 2, 9.0F},  {1, 352, 6, 0, 40, 2, 3, 6.0F},
                            {1, 256, 4, 4, 0, 1, 1, 3.0F},   {1, 144, 2, 0, 8, 5, 4, 2.0F},   {1, 52, 2, 0, 0, 1, 1, 2.5F},
                            {1, 36, 4, 5, 4, 3, 4, 4.0F},    {1, 7, 2, 0, 0, 1, 1, 3.0F},     {2, 384, 2, 0, 136, 1, 1, 1.5F},
                            {2, 208, 2, 0, 136, 6, 5, 1.2F}, {2, 96, 4, 4, 16, 1, 2, 5.0F},   {2, 30, 6, 0, 2, 2, 3, 6.0F},
                            {2, 15, 1, 0, 0, 1, 1, 2.0F},    {2, 6, 2, 0, 0, 3, 2, 3.0F},     {1, 320, 8, 0, 8, 3, 8, 20.0F},
                            {2, 352, 8, 2, 24, 1, 3, 12.0F}};
    std::normal_distribution<float> noise(0.0F, 1.0F);
    unsigned dm_bad = 0, dec_bad = 0, dec_ok_crc = 0, nshapes = 0;
    for (const shape& sh : shapes) {
      ++nshapes;
      const unsigned Kb = sh.bg == 1 ? 22 : 10, K = Kb * sh.Z, N = (sh.bg == 1 ? 66 : 50) * sh.Z;
      codeblock_metadata meta;
      meta.tb_common.base_graph        = sh.bg == 1 ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
      meta.tb_common.lifting_size      = static_cast<ldpc::lifting_size_t>(sh.Z);
      meta.tb_common.mod               = static_cast<modulation_scheme>(sh.Qm);
      meta.tb_common.Nref              = sh.nref_div ? N * sh.nref_div / 6 : 0;
      meta.cb_specific.nof_filler_bits = sh.nfill;
      meta.cb_specific.nof_crc_bits    = 24;
      meta.cb_specific.full_length     = N;
      unsigned E                       = N * sh.e_num / sh.e_den;
      E -= E % sh.Qm;
      meta.cb_specific.rm_length = E;
      // Message with CRC24B over the first K - F - 24 bits, fillers zero.
      const unsigned       nb = K - sh.nfill;
      std::vector<uint8_t> msg_bits(K, 0);
      for (unsigned i = 0; i + 24 < nb; ++i) {
        msg_bits[i] = static_cast<uint8_t>(rgen() & 1U);
      }
      auto     crc24b = crc_factory->create(crc_generator_poly::CRC24B);
      unsigned crc    = crc24b->calculate_bit(span<const uint8_t>(msg_bits.data(), nb - 24));
      for (unsigned b = 0; b != 24; ++b) {
        msg_bits[nb - 24 + b] = (crc >> (23 - b)) & 1U;
      }
      dynamic_bit_buffer msg(K), cw(N), rmd(E);
      for (unsigned i = 0; i != K; ++i) {
        msg.insert(msg_bits[i], i, 1);
      }
      enc->encode(cw, msg, meta.tb_common);
      std::vector<log_likelihood_ratio> soft_cpu(N, 0), soft_gpu(N, 0), in(E);
      for (unsigned rv : {0U, 2U, 3U}) {
        meta.tb_common.rv = rv;
        rm->rate_match(rmd, cw, meta);
        for (unsigned i = 0; i != E; ++i) {
          float v = (rmd.extract(i, 1) ? -sh.amp : sh.amp) + sh.amp * 0.7F * noise(rgen);
          in[i]   = log_likelihood_ratio(static_cast<int>(std::max(-120.0F, std::min(120.0F, std::round(4.0F * v)))));
        }
        dem_cpu->rate_dematch(soft_cpu, in, rv == 0, meta);
        dem_gpu->rate_dematch(soft_gpu, in, rv == 0, meta);
        if (!std::equal(soft_cpu.begin(), soft_cpu.end(), soft_gpu.begin())) {
          ++dm_bad;
          std::printf("MISMATCH rate_dematch bg=%u Z=%u Qm=%u Nref=%u F=%u E=%u rv=%u\n", sh.bg, sh.Z, sh.Qm, meta.tb_common.Nref,
                      sh.nfill, E, rv);
        }
        for (bool with_crc : {true, false}) {
          std::vector<uint8_t> out_cpu((K + 7) / 8 + 8, 0x5A), out_gpu((K + 7) / 8 + 8, 0x5A);
          bit_buffer bb_cpu = bit_buffer::from_bytes(out_cpu).first(K), bb_gpu = bit_buffer::from_bytes(out_gpu).first(K);
          ldpc_decoder::configuration dcfg;
          dcfg.block_conf               = meta;
          dcfg.algorithm_conf.max_iterations = with_crc ? 6 : 3;
          auto it_cpu = dec_cpu->decode(bb_cpu, soft_cpu, with_crc ? crc24b.get() : nullptr, dcfg);
          auto it_gpu = dec_gpu->decode(bb_gpu, soft_gpu, with_crc ? crc24b.get() : nullptr, dcfg);
          dec_ok_crc += (with_crc && it_cpu.has_value()) ? 1 : 0;
          if (it_cpu != it_gpu || out_cpu != out_gpu) {
            ++dec_bad;
            std::printf("MISMATCH ldpc_decoder bg=%u Z=%u F=%u rv=%u crc=%d\n", sh.bg, sh.Z, sh.nfill, rv, (int)with_crc);
          }
        }
      }
    }
    std::printf("ldpc_rate_dematcher / ldpc_decoder cuda vs auto, %u shapes x rv 0,2,3: %u / %u mismatches (%u early stops)\n",
                nshapes, dm_bad, dec_bad, dec_ok_crc);
    failures += dm_bad + dec_bad;
    if (dec_ok_crc < nshapes) {
      std::printf("too few successful decodes: the early stop is barely exercised\n");
      ++failures;
    }
  }
  std::printf("hwacc_parity: %s (%d mismatches)\n", failures ? "FAILED" : "PASSED", failures);
  return failures ? 1 : 0;
}
