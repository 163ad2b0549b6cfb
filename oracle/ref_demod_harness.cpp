// TEST INFRASTRUCTURE (oracle): C interface to the UNMODIFIED reference's PUSCH soft-demodulation chain, compiled from
// the sources where they lie under /root/reference by oracle/Makefile into oracle/_ref/libref_oracle.so. It pins the
// device-side demodulation / descrambling / UL-SCH demultiplexing kernels:
//   ref_demodulate_soft       demodulation_mapper::demodulate_soft (lib/phy/upper/channel_modulation/demodulation_mapper_impl.cpp:76-106
//                             and demodulation_mapper_{qpsk,qam16,qam64,qam256}.cpp), one call = one block
//   ref_scrambling_sequence   pseudo_random_generator (TS 38.211 5.2.1; lib/phy/upper/sequence_generators/pseudo_random_generator_impl.cpp)
//   ref_pusch_demodulate      pusch_demodulator_impl::demodulate (pusch_demodulator_impl.cpp:129-301) feeding
//                             ulsch_demultiplex_impl (ulsch_demultiplex_impl.cpp:208-327) feeding a pusch_decoder_buffer that
//                             records the soft bits. The channel equalizer is the only stand-in: an injected stub that hands out
//                             the caller's equalized symbols and noise variances in order (the equalizer is upstream of this
//                             path), so the block partition, the demapper's SIMD / scalar-tail split, the descrambling and the
//                             demultiplexer's bypass are the reference's own code.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use this library.
#include "pusch_demodulator_impl.h"
#include "ulsch_demultiplex_impl.h"
#include "srsran/phy/support/resource_grid_reader_empty.h"
#include "srsran/phy/upper/channel_estimation.h"
#include "srsran/phy/upper/channel_modulation/channel_modulation_factories.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_codeword_buffer.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_buffer.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_demodulator_notifier.h"
#include "srsran/phy/upper/sequence_generators/sequence_generator_factories.h"
#include <cstring>
#include <vector>

using namespace srsran;

namespace {

modulation_scheme to_mod(int qm, int pi2)
{
  switch (qm) {
    case 1:
      return pi2 ? modulation_scheme::PI_2_BPSK : modulation_scheme::BPSK;
    case 2:
      return modulation_scheme::QPSK;
    case 4:
      return modulation_scheme::QAM16;
    case 6:
      return modulation_scheme::QAM64;
    default:
      return modulation_scheme::QAM256;
  }
}

/// Hands out the caller's equalized symbols / noise variances block after block.
class equalizer_feed : public channel_equalizer
{
public:
  const cf_t*  sym  = nullptr;
  const float* nv   = nullptr;
  size_t       pos  = 0;
  size_t       size = 0;

  void equalize(span<cf_t> eq_symbols, span<float> eq_noise_vars, const re_list&, const ch_est_list&, span<const float>, float) override
  {
    srsran_assert(pos + eq_symbols.size() <= size, "equalizer feed exhausted");
    std::memcpy(eq_symbols.data(), sym + pos, eq_symbols.size() * sizeof(cf_t));
    std::memcpy(eq_noise_vars.data(), nv + pos, eq_noise_vars.size() * sizeof(float));
    pos += eq_symbols.size();
  }
};

class notifier_null : public pusch_demodulator_notifier
{
public:
  void on_provisional_stats(const demodulation_stats&) override {}
  void on_end_stats(const demodulation_stats&) override {}
};

class softbit_recorder : public pusch_decoder_buffer
{
public:
  std::vector<log_likelihood_ratio> data;
  std::vector<log_likelihood_ratio> scratch;
  bool                              ended = false;

  span<log_likelihood_ratio> get_next_block_view(unsigned block_size) override
  {
    scratch.resize(block_size);
    return scratch;
  }
  void on_new_softbits(span<const log_likelihood_ratio> softbits) override
  {
    data.insert(data.end(), softbits.begin(), softbits.end());
  }
  void on_end_softbits() override { ended = true; }
};

} // namespace

extern "C" {

/// demodulation_mapper::demodulate_soft on one block of n symbols (symbols: n x (re, im) floats). Returns 0.
int ref_demodulate_soft(int qm, int pi2, int8_t* llrs, const float* symbols, const float* noise_vars, uint32_t n)
{
  static std::unique_ptr<demodulation_mapper> demapper = create_channel_modulation_sw_factory()->create_demodulation_mapper();
  demapper->demodulate_soft(span<log_likelihood_ratio>(reinterpret_cast<log_likelihood_ratio*>(llrs), static_cast<size_t>(n) * qm),
                            span<const cf_t>(reinterpret_cast<const cf_t*>(symbols), n),
                            span<const float>(noise_vars, n),
                            to_mod(qm, pi2));
  return 0;
}

/// The first nbits bits of the scrambling sequence of TS 38.211 5.2.1 initialised with c_init, one bit per byte.
int ref_scrambling_sequence(uint32_t c_init, uint8_t* bits, uint32_t nbits)
{
  std::unique_ptr<pseudo_random_generator> prg = create_pseudo_random_generator_sw_factory()->create();
  prg->init(c_init);
  std::vector<uint8_t> zeros(nbits, 0);
  prg->apply_xor(span<uint8_t>(bits, nbits), span<const uint8_t>(zeros.data(), nbits));
  return 0;
}

/// pusch_demodulator_impl::demodulate + ulsch_demultiplex_impl (no UCI) for one PUSCH allocation of `nof_prb` PRB starting
/// at PRB 0, OFDM symbols [start_symbol, start_symbol + nof_symbols), DM-RS type 1 in the symbols of dmrs_mask with
/// nof_cdm_groups_without_data CDM groups. `symbols` / `noise_vars`: the equalizer's output in the order the demodulator
/// consumes it ([re][layer], OFDM symbol after OFDM symbol). Writes the SCH soft bits to llr_out. Returns their number.
int ref_pusch_demodulate(int          qm,
                         int          pi2,
                         uint32_t     rnti,
                         uint32_t     n_id,
                         uint32_t     nof_layers,
                         uint32_t     nof_prb,
                         uint32_t     start_symbol,
                         uint32_t     nof_symbols,
                         uint32_t     dmrs_mask,
                         uint32_t     nof_cdm_groups_without_data,
                         const float* symbols,
                         const float* noise_vars,
                         uint32_t     nof_eq_symbols,
                         int8_t*      llr_out,
                         uint32_t     llr_capacity)
{
  auto feed   = std::make_unique<equalizer_feed>();
  auto* feedp = feed.get();
  feedp->sym  = reinterpret_cast<const cf_t*>(symbols);
  feedp->nv   = noise_vars;
  feedp->size = nof_eq_symbols;
  pusch_demodulator_impl demod(std::move(feed),
                               create_channel_modulation_sw_factory()->create_demodulation_mapper(),
                               nullptr,
                               create_pseudo_random_generator_sw_factory()->create(),
                               false);
  ulsch_demultiplex_impl demux;

  pusch_demodulator::configuration cfg;
  cfg.rnti    = static_cast<uint16_t>(rnti);
  cfg.rb_mask = bounded_bitset<MAX_RB>(nof_prb);
  cfg.rb_mask.fill(0, nof_prb);
  cfg.modulation         = to_mod(qm, pi2);
  cfg.start_symbol_index = start_symbol;
  cfg.nof_symbols        = nof_symbols;
  cfg.dmrs_symb_pos      = symbol_slot_mask(MAX_NSYMB_PER_SLOT);
  for (unsigned i = 0; i != MAX_NSYMB_PER_SLOT; ++i) {
    if ((dmrs_mask >> i) & 1U) {
      cfg.dmrs_symb_pos.set(i);
    }
  }
  cfg.dmrs_config_type            = dmrs_type::TYPE1;
  cfg.nof_cdm_groups_without_data = nof_cdm_groups_without_data;
  cfg.n_id                        = n_id;
  cfg.nof_tx_layers               = nof_layers;
  cfg.rx_ports                    = {0};

  ulsch_demultiplex::configuration dcfg;
  dcfg.modulation                  = cfg.modulation;
  dcfg.nof_layers                  = nof_layers;
  dcfg.nof_prb                     = nof_prb;
  dcfg.start_symbol_index          = start_symbol;
  dcfg.nof_symbols                 = nof_symbols;
  dcfg.nof_harq_ack_rvd            = 0;
  dcfg.dmrs                        = dmrs_type::TYPE1;
  dcfg.dmrs_symbol_mask            = cfg.dmrs_symb_pos;
  dcfg.nof_cdm_groups_without_data = nof_cdm_groups_without_data;
  dcfg.nof_harq_ack_bits           = 0;
  dcfg.nof_enc_harq_ack_bits       = 0;
  dcfg.nof_csi_part1_bits          = 0;
  dcfg.nof_enc_csi_part1_bits      = 0;

  softbit_recorder sch, ack, csi1;
  pusch_codeword_buffer& cw = demux.demultiplex(sch, ack, csi1, dcfg);

  resource_grid_reader_empty grid(1, MAX_NSYMB_PER_SLOT, nof_prb);
  channel_estimate::channel_estimate_dimensions dims;
  dims.nof_prb       = nof_prb;
  dims.nof_symbols   = MAX_NSYMB_PER_SLOT;
  dims.nof_rx_ports  = 1;
  dims.nof_tx_layers = nof_layers;
  channel_estimate est(dims);
  notifier_null    notifier;
  demod.demodulate(cw, notifier, grid, est, cfg);

  if (!sch.ended || sch.data.size() > llr_capacity || feedp->pos != nof_eq_symbols) {
    return -1;
  }
  std::memcpy(llr_out, sch.data.data(), sch.data.size());
  return static_cast<int>(sch.data.size());
}

} // extern "C"
