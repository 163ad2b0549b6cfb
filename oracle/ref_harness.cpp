// TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT PATH.
//
// Thin extern "C" wrapper around the UNMODIFIED reference classes (srsRAN Project), compiled by oracle/Makefile from
// the sources where they lie under /root/reference into oracle/_ref/libref_oracle.so. Nothing of the reference is
// copied into this repository: this file only instantiates the reference's factories and forwards plain pointers.
//
// What it exposes (all names ref_*):
//   - the reference's ldpc_rate_dematcher / ldpc_decoder / crc_calculator through their public factories
//     (lib/phy/upper/channel_coding/channel_coding_factories.cpp:74-175), selectable by the reference's own type string
//     ("auto", "generic", "avx2", "avx512", "lut", "clmul");
//   - the reference's Tx chain (segmenter_tx + ldpc_encoder + ldpc_rate_matcher via pdsch_encoder_impl,
//     lib/phy/upper/channel_processors/pdsch_encoder_impl.cpp:28-78) to synthesise valid codewords;
//   - the reference's TB-level software decoder pusch_decoder_impl
//     (lib/phy/upper/channel_processors/pusch/pusch_decoder_impl.cpp:89-497) over a harness-owned rx_buffer whose
//     code-block storage is never cleared, like rx_buffer_pool (include/srsran/phy/upper/rx_buffer_pool.h:62-63);
//   - timing loops used ONLY by bench.py's cpu_baseline / --impl reference legs.

#include "pdsch_encoder_impl.h"
#include "pusch_codeblock_decoder.h"
#include "pusch_decoder_impl.h"
#include "srsran/phy/upper/channel_coding/channel_coding_factories.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_buffer.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_notifier.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_result.h"
#include "srsran/phy/upper/unique_rx_buffer.h"
#include "srsran/support/cpu_features.h"
#include "srsran/support/executors/task_worker_pool.h"
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <thread>
#include <vector>

using namespace srsran;

namespace {

crc_generator_poly to_poly(int poly)
{
  switch (poly) {
    case 1:
      return crc_generator_poly::CRC24A;
    case 2:
      return crc_generator_poly::CRC24B;
    default:
      return crc_generator_poly::CRC16;
  }
}

struct cache_t {
  std::map<std::string, std::unique_ptr<ldpc_decoder>>        decoders;
  std::map<std::string, std::unique_ptr<ldpc_rate_dematcher>> dematchers;
  std::map<std::string, std::unique_ptr<crc_calculator>>      crcs;
};

cache_t& cache()
{
  static thread_local cache_t c;
  return c;
}

ldpc_decoder* get_decoder(const char* type)
{
  auto& m = cache().decoders;
  auto  i = m.find(type);
  if (i == m.end()) {
    auto f = create_ldpc_decoder_factory_sw(type);
    if (!f) {
      return nullptr;
    }
    i = m.emplace(type, f->create()).first;
  }
  return i->second.get();
}

ldpc_rate_dematcher* get_dematcher(const char* type)
{
  auto& m = cache().dematchers;
  auto  i = m.find(type);
  if (i == m.end()) {
    auto f = create_ldpc_rate_dematcher_factory_sw(type);
    if (!f) {
      return nullptr;
    }
    i = m.emplace(type, f->create()).first;
  }
  return i->second.get();
}

crc_calculator* get_crc(const char* type, int poly)
{
  std::string key = std::string(type) + "/" + std::to_string(poly);
  auto&       m   = cache().crcs;
  auto        i   = m.find(key);
  if (i == m.end()) {
    auto f = create_crc_calculator_factory_sw(type);
    if (!f) {
      return nullptr;
    }
    i = m.emplace(key, f->create(to_poly(poly))).first;
  }
  return i->second.get();
}

codeblock_metadata make_meta(int bg, int Z, int rv, int Qm, unsigned Nref, unsigned F, unsigned crc_bits)
{
  codeblock_metadata meta        = {};
  meta.tb_common.base_graph      = (bg == 1) ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
  meta.tb_common.lifting_size    = static_cast<ldpc::lifting_size_t>(Z);
  meta.tb_common.rv              = rv;
  meta.tb_common.mod             = static_cast<modulation_scheme>(Qm);
  meta.tb_common.Nref            = Nref;
  meta.cb_specific.nof_filler_bits = F;
  meta.cb_specific.nof_crc_bits  = crc_bits;
  return meta;
}

// Harness-owned HARQ buffer: code-block storage that is zero-initialised once and never cleared afterwards.
class harness_rx_buffer : public unique_rx_buffer::callback
{
public:
  explicit harness_rx_buffer(unsigned nof_cbs_) :
    nof_cbs(nof_cbs_), crcs(new bool[nof_cbs_]()), soft(nof_cbs_), data(nof_cbs_)
  {
    for (unsigned i = 0; i != nof_cbs; ++i) {
      soft[i].assign(ldpc::MAX_CODEBLOCK_SIZE, log_likelihood_ratio(0));
      data[i].assign(ldpc::MAX_CODEBLOCK_SIZE / 8 + 8, 0);
    }
  }
  unsigned   get_nof_codeblocks() const override { return nof_cbs; }
  void       reset_codeblocks_crc() override { std::fill(crcs.get(), crcs.get() + nof_cbs, false); }
  span<bool> get_codeblocks_crc() override { return span<bool>(crcs.get(), nof_cbs); }
  unsigned   get_absolute_codeblock_id(unsigned codeblock_id) const override { return codeblock_id; }
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
  void release() override { ++nof_releases; }

  unsigned                                       nof_cbs;
  std::unique_ptr<bool[]>                        crcs;
  std::vector<std::vector<log_likelihood_ratio>> soft;
  std::vector<std::vector<uint8_t>>              data;
  unsigned                                       nof_releases = 0;
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
  std::atomic<bool>    done{false};
};

struct pusch_handle {
  std::shared_ptr<crc_calculator_factory>                     crc_factory;
  std::shared_ptr<ldpc_segmenter_rx_factory>                  seg_factory;
  std::shared_ptr<pusch_decoder_impl::codeblock_decoder_pool> pool;
  std::unique_ptr<task_worker_pool<concurrent_queue_policy::lockfree_mpmc>> workers;
  std::unique_ptr<task_executor>                              executor;
  std::vector<std::unique_ptr<pusch_decoder_impl>>            decoders;
  std::map<uint64_t, std::unique_ptr<harness_rx_buffer>>      harq;
  unsigned                                                    nof_threads;
};

std::unique_ptr<pusch_decoder_impl> make_decoder(pusch_handle& h)
{
  pusch_decoder_impl::sch_crc crcs;
  crcs.crc16  = h.crc_factory->create(crc_generator_poly::CRC16);
  crcs.crc24A = h.crc_factory->create(crc_generator_poly::CRC24A);
  crcs.crc24B = h.crc_factory->create(crc_generator_poly::CRC24B);
  return std::make_unique<pusch_decoder_impl>(
      h.seg_factory->create(), h.pool, std::move(crcs), h.executor.get(), MAX_RB, 4);
}

} // namespace

extern "C" {

struct ref_cb_meta {
  uint32_t bg, Z, full_length, rm_length, nof_filler_bits, cw_offset, nof_crc_bits;
};

struct ref_tb_result {
  int32_t  tb_crc_ok;
  uint32_t nof_codeblocks;
  uint32_t nof_observations;
  uint32_t iter_min;
  uint32_t iter_max;
  float    iter_mean;
};

/// Reports which SIMD flavours the reference's "auto" factories can pick on this host.
const char* ref_info()
{
  static std::string s;
  s = std::string("avx2=") + (cpu_supports_feature(cpu_feature::avx2) ? "1" : "0") + " avx512f=" +
      (cpu_supports_feature(cpu_feature::avx512f) ? "1" : "0") + " avx512bw=" +
      (cpu_supports_feature(cpu_feature::avx512bw) ? "1" : "0") + " avx512vbmi=" +
      (cpu_supports_feature(cpu_feature::avx512vbmi) ? "1" : "0") + " pclmul=" +
      (cpu_supports_feature(cpu_feature::pclmul) ? "1" : "0");
  return s.c_str();
}

/// ldpc_rate_dematcher::rate_dematch on a caller-owned N-byte soft buffer. Returns 0, or -1 if the type is unsupported.
int ref_dematch(const char*   type,
                int8_t*       softbuf,
                uint32_t      N,
                const int8_t* llrs,
                uint32_t      E,
                int           new_data,
                int           rv,
                int           Qm,
                uint32_t      Nref,
                uint32_t      F)
{
  ldpc_rate_dematcher* dm = get_dematcher(type);
  if (dm == nullptr) {
    return -1;
  }
  codeblock_metadata meta = make_meta(1, 2, rv, Qm, Nref, F, 24);
  dm->rate_dematch(span<log_likelihood_ratio>(reinterpret_cast<log_likelihood_ratio*>(softbuf), N),
                   span<const log_likelihood_ratio>(reinterpret_cast<const log_likelihood_ratio*>(llrs), E),
                   new_data != 0,
                   meta);
  return 0;
}

/// ldpc_decoder::decode. crc_poly: 0 = no CRC (nullptr), 1 = CRC24A, 2 = CRC24B, 3 = CRC16.
/// out_packed holds K_bg * Z bits MSB-first and is only written where the reference writes it.
/// Returns the iteration count, -1 for std::nullopt, -2 if the type is unsupported on this host.
int ref_ldpc_decode(const char*   type,
                    uint8_t*      out_packed,
                    const int8_t* in,
                    uint32_t      n_in,
                    int           bg,
                    int           Z,
                    uint32_t      F,
                    int           crc_poly,
                    int           max_it,
                    float         scaling)
{
  ldpc_decoder* dec = get_decoder(type);
  if (dec == nullptr) {
    return -2;
  }
  crc_calculator* crc = (crc_poly == 0) ? nullptr : get_crc("auto", crc_poly);
  unsigned        K   = ((bg == 1) ? 22 : 10) * Z;

  ldpc_decoder::configuration cfg;
  cfg.block_conf                    = make_meta(bg, Z, 0, 1, 0, F, (crc_poly == 3) ? 16 : 24);
  cfg.algorithm_conf.max_iterations = max_it;
  cfg.algorithm_conf.scaling_factor = scaling;

  bit_buffer out = bit_buffer::from_bytes(span<uint8_t>(out_packed, (K + 7) / 8)).first(K);
  std::optional<unsigned> r =
      dec->decode(out, span<const log_likelihood_ratio>(reinterpret_cast<const log_likelihood_ratio*>(in), n_in), crc, cfg);
  return r.has_value() ? static_cast<int>(r.value()) : -1;
}

/// crc_calculator::calculate over the first nbits bits of an MSB-first packed buffer.
uint32_t ref_crc(const char* type, int poly, const uint8_t* packed, uint32_t nbits)
{
  crc_calculator* crc = get_crc(type, poly);
  if (crc == nullptr) {
    return 0xffffffffU;
  }
  // bit_buffer needs a mutable span; it is only read.
  bit_buffer buf =
      bit_buffer::from_bytes(span<uint8_t>(const_cast<uint8_t*>(packed), (nbits + 7) / 8)).first(nbits);
  return crc->calculate(buf);
}

/// crc_calculator::calculate_byte.
uint32_t ref_crc_byte(const char* type, int poly, const uint8_t* bytes, uint32_t nbytes)
{
  crc_calculator* crc = get_crc(type, poly);
  if (crc == nullptr) {
    return 0xffffffffU;
  }
  return crc->calculate_byte(span<const uint8_t>(bytes, nbytes));
}

/// crc_calculator::calculate_bit (one bit per byte).
uint32_t ref_crc_bit(const char* type, int poly, const uint8_t* bits, uint32_t nbits)
{
  crc_calculator* crc = get_crc(type, poly);
  if (crc == nullptr) {
    return 0xffffffffU;
  }
  return crc->calculate_bit(span<const uint8_t>(bits, nbits));
}

/// hard_decision (lib/phy/upper/log_likelihood_ratio.cpp:226-252). Returns the "no zero LLR" flag.
int ref_hard_decision(uint8_t* out_packed, const int8_t* llrs, uint32_t n)
{
  bit_buffer out = bit_buffer::from_bytes(span<uint8_t>(out_packed, (n + 7) / 8)).first(n);
  return hard_decision(out, span<const log_likelihood_ratio>(reinterpret_cast<const log_likelihood_ratio*>(llrs), n)) ? 1 : 0;
}

/// ldpc_segmenter_rx::segment metadata for a TB. Returns the number of code blocks.
int ref_segment_rx(uint32_t     tbs_bits,
                   int          bg,
                   int          rv,
                   int          Qm,
                   uint32_t     Nref,
                   int          nof_layers,
                   uint32_t     nof_llrs,
                   ref_cb_meta* out)
{
  static thread_local std::unique_ptr<ldpc_segmenter_rx> seg = create_ldpc_segmenter_rx_factory_sw()->create();
  std::vector<log_likelihood_ratio>                       dummy(nof_llrs);
  static_vector<described_rx_codeblock, MAX_NOF_SEGMENTS> cbs;
  segmenter_config                                        cfg;
  cfg.base_graph     = (bg == 1) ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
  cfg.rv             = rv;
  cfg.mod            = static_cast<modulation_scheme>(Qm);
  cfg.Nref           = Nref;
  cfg.nof_layers     = nof_layers;
  cfg.nof_ch_symbols = nof_llrs / Qm;
  seg->segment(cbs, dummy, tbs_bits, cfg);
  for (unsigned i = 0; i != cbs.size(); ++i) {
    const codeblock_metadata& m = cbs[i].second;
    out[i] = {static_cast<uint32_t>(m.tb_common.base_graph),
              static_cast<uint32_t>(m.tb_common.lifting_size),
              m.cb_specific.full_length,
              m.cb_specific.rm_length,
              m.cb_specific.nof_filler_bits,
              m.cb_specific.cw_offset,
              m.cb_specific.nof_crc_bits};
  }
  return cbs.size();
}

/// Reference Tx chain: TB bytes -> codeword bits (one bit per byte, nof_ch_symbols * Qm of them).
int ref_encode_tb(const uint8_t* tb,
                  uint32_t       tb_bytes,
                  int            bg,
                  int            rv,
                  int            Qm,
                  uint32_t       Nref,
                  int            nof_layers,
                  uint32_t       nof_ch_symbols,
                  uint8_t*       codeword_bits)
{
  static thread_local std::unique_ptr<pdsch_encoder_impl> enc;
  if (!enc) {
    auto crc_f = create_crc_calculator_factory_sw("auto");
    auto seg_f = create_ldpc_segmenter_tx_factory_sw(crc_f);
    auto enc_f = create_ldpc_encoder_factory_sw("auto");
    auto rm_f  = create_ldpc_rate_matcher_factory_sw();
    enc        = std::make_unique<pdsch_encoder_impl>(seg_f->create(), enc_f->create(), rm_f->create());
  }
  pdsch_encoder::configuration cfg;
  cfg.base_graph     = (bg == 1) ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
  cfg.rv             = rv;
  cfg.mod            = static_cast<modulation_scheme>(Qm);
  cfg.Nref           = Nref;
  cfg.nof_layers     = nof_layers;
  cfg.nof_ch_symbols = nof_ch_symbols;
  enc->encode(span<uint8_t>(codeword_bits, static_cast<size_t>(nof_ch_symbols) * Qm), span<const uint8_t>(tb, tb_bytes), cfg);
  return 0;
}

/// Reference LDPC encoder on one message of K_bg*Z bits (one bit per byte, filler bits zero) -> N = 66Z/50Z bits.
int ref_ldpc_encode(const uint8_t* msg_bits, int bg, int Z, uint8_t* cw_bits)
{
  static thread_local std::unique_ptr<ldpc_encoder> enc = create_ldpc_encoder_factory_sw("auto")->create();
  unsigned K = ((bg == 1) ? 22 : 10) * Z;
  unsigned N = ((bg == 1) ? 66 : 50) * Z;
  dynamic_bit_buffer in(K), out(N);
  for (unsigned i = 0; i != K; ++i) {
    in.insert(msg_bits[i] & 1U, i, 1);
  }
  codeblock_metadata meta = make_meta(bg, Z, 0, 1, 0, 0, 24);
  enc->encode(out, in, meta.tb_common);
  for (unsigned i = 0; i != N; ++i) {
    cw_bits[i] = out.extract(i, 1);
  }
  return 0;
}

/// TB-level software decoder. nof_threads <= 1: synchronous; otherwise the reference's own CB-level fan-out on a
/// task_worker_pool (pusch_decoder_impl.cpp:372-381).
void* ref_pusch_create(const char* dec_type, const char* dem_type, const char* crc_type, int nof_threads)
{
  auto h         = std::make_unique<pusch_handle>();
  h->nof_threads = (nof_threads < 1) ? 1 : nof_threads;
  h->crc_factory = create_crc_calculator_factory_sw(crc_type);
  h->seg_factory = create_ldpc_segmenter_rx_factory_sw();
  auto dec_f     = create_ldpc_decoder_factory_sw(dec_type);
  auto dem_f     = create_ldpc_rate_dematcher_factory_sw(dem_type);
  if (!h->crc_factory || !dec_f || !dem_f) {
    return nullptr;
  }
  unsigned nof_cb_decoders = (h->nof_threads > 1) ? h->nof_threads + 1 : 1;
  std::vector<std::unique_ptr<pusch_codeblock_decoder>> cb_decoders(nof_cb_decoders);
  for (auto& d : cb_decoders) {
    pusch_codeblock_decoder::sch_crc crcs;
    crcs.crc16  = h->crc_factory->create(crc_generator_poly::CRC16);
    crcs.crc24A = h->crc_factory->create(crc_generator_poly::CRC24A);
    crcs.crc24B = h->crc_factory->create(crc_generator_poly::CRC24B);
    d           = std::make_unique<pusch_codeblock_decoder>(dem_f->create(), dec_f->create(), crcs);
  }
  h->pool = std::make_shared<pusch_decoder_impl::codeblock_decoder_pool>(std::move(cb_decoders));
  if (h->nof_threads > 1) {
    h->workers = std::make_unique<task_worker_pool<concurrent_queue_policy::lockfree_mpmc>>(
        "ref_pusch_dec", h->nof_threads, 4096);
    h->executor = std::make_unique<task_worker_pool_executor<concurrent_queue_policy::lockfree_mpmc>>(*h->workers);
  }
  h->decoders.push_back(make_decoder(*h));
  return h.release();
}

void ref_pusch_destroy(void* handle)
{
  auto* h = static_cast<pusch_handle*>(handle);
  if (h == nullptr) {
    return;
  }
  h->decoders.clear();
  if (h->workers) {
    h->workers->stop();
  }
  delete h;
}

/// Decodes one TB transmission through pusch_decoder_impl (new_data -> on_new_softbits -> on_end_softbits -> notifier).
/// The HARQ buffer keyed by harq_key persists in the handle across calls. If softbuf_out is not null it receives, for
/// each code block, the first `full_length` soft bits of the buffer after the call (concatenated).
int ref_pusch_decode(void*          handle,
                     uint64_t       harq_key,
                     uint8_t*       tb_out,
                     uint32_t       tb_bytes,
                     const int8_t*  llrs,
                     uint32_t       nof_llrs,
                     int            bg,
                     int            rv,
                     int            Qm,
                     uint32_t       Nref,
                     int            nof_layers,
                     int            max_it,
                     int            early_stop,
                     int            new_data,
                     ref_tb_result* result,
                     uint8_t*       cb_crc_out,
                     int8_t*        softbuf_out)
{
  auto*    h       = static_cast<pusch_handle*>(handle);
  unsigned nof_cbs = ldpc::compute_nof_codeblocks(units::bits(tb_bytes * 8),
                                                  (bg == 1) ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2);
  auto& slot = h->harq[harq_key];
  if (!slot || slot->nof_cbs != nof_cbs) {
    slot = std::make_unique<harness_rx_buffer>(nof_cbs);
  }
  if (new_data) {
    // rx_buffer_pool_impl::reserve resets the CRC flags of a buffer reserved for new data.
    slot->reset_codeblocks_crc();
  }

  pusch_decoder::configuration cfg;
  cfg.base_graph          = (bg == 1) ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
  cfg.rv                  = rv;
  cfg.mod                 = static_cast<modulation_scheme>(Qm);
  cfg.Nref                = Nref;
  cfg.nof_layers          = nof_layers;
  cfg.nof_ldpc_iterations = max_it;
  cfg.use_early_stop      = early_stop != 0;
  cfg.new_data            = new_data != 0;

  notifier_t            notifier;
  pusch_decoder_impl&   dec = *h->decoders[0];
  pusch_decoder_buffer& buf = dec.new_data(span<uint8_t>(tb_out, tb_bytes), unique_rx_buffer(*slot), notifier, cfg);
  buf.on_new_softbits(span<const log_likelihood_ratio>(reinterpret_cast<const log_likelihood_ratio*>(llrs), nof_llrs));
  buf.on_end_softbits();
  while (!notifier.done.load()) {
    std::this_thread::yield();
  }

  result->tb_crc_ok        = notifier.res.tb_crc_ok ? 1 : 0;
  result->nof_codeblocks   = notifier.res.nof_codeblocks_total;
  result->nof_observations = notifier.res.ldpc_decoder_stats.get_nof_observations();
  result->iter_min         = result->nof_observations ? notifier.res.ldpc_decoder_stats.get_min() : 0;
  result->iter_max         = result->nof_observations ? notifier.res.ldpc_decoder_stats.get_max() : 0;
  result->iter_mean        = result->nof_observations ? notifier.res.ldpc_decoder_stats.get_mean() : 0;
  if (cb_crc_out != nullptr) {
    for (unsigned i = 0; i != nof_cbs; ++i) {
      cb_crc_out[i] = slot->crcs[i] ? 1 : 0;
    }
  }
  if (softbuf_out != nullptr) {
    ref_cb_meta metas[MAX_NOF_SEGMENTS];
    ref_segment_rx(tb_bytes * 8, bg, rv, Qm, Nref, nof_layers, nof_llrs, metas);
    size_t off = 0;
    for (unsigned i = 0; i != nof_cbs; ++i) {
      std::memcpy(softbuf_out + off, slot->soft[i].data(), metas[i].full_length);
      off += metas[i].full_length;
    }
  }
  return 0;
}

/// Timing loop for the CPU baseline: decodes the same TB `reps` times as new data (fresh decode every time) on
/// nof_threads workers created with ref_pusch_create. Returns seconds of wall time.
double ref_pusch_bench(void*         handle,
                       uint32_t      tb_bytes,
                       const int8_t* llrs,
                       uint32_t      nof_llrs,
                       int           bg,
                       int           Qm,
                       uint32_t      Nref,
                       int           nof_layers,
                       int           max_it,
                       int           early_stop,
                       int           reps,
                       int*          nof_crc_ok)
{
  std::vector<uint8_t> tb(tb_bytes);
  ref_tb_result        res;
  int                  ok = 0;
  auto                 t0 = std::chrono::steady_clock::now();
  for (int r = 0; r != reps; ++r) {
    ref_pusch_decode(
        handle, 0, tb.data(), tb_bytes, llrs, nof_llrs, bg, 0, Qm, Nref, nof_layers, max_it, early_stop, 1, &res, nullptr, nullptr);
    ok += res.tb_crc_ok;
  }
  auto t1 = std::chrono::steady_clock::now();
  if (nof_crc_ok != nullptr) {
    *nof_crc_ok = ok;
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

/// Timing loop for the CPU baseline on all host threads: nof_threads independent synchronous decoders (one per
/// std::thread, like pusch_processor_benchmark's worker threads, pusch_processor_benchmark.cpp:757-781), each decoding
/// the same TB `reps` times as new data. Returns the wall time in seconds for nof_threads * reps TBs.
double ref_pusch_bench_mt(const char*   dec_type,
                          int           nof_threads,
                          uint32_t      tb_bytes,
                          const int8_t* llrs,
                          uint32_t      nof_llrs,
                          int           bg,
                          int           Qm,
                          uint32_t      Nref,
                          int           nof_layers,
                          int           max_it,
                          int           early_stop,
                          int           reps,
                          int*          nof_crc_ok)
{
  std::vector<void*> handles(nof_threads);
  for (auto& h : handles) {
    h = ref_pusch_create(dec_type, "auto", "auto", 1);
    if (h == nullptr) {
      return -1;
    }
  }
  std::vector<int>         oks(nof_threads, 0);
  std::vector<std::thread> threads;
  std::atomic<int>         ready{0};
  std::atomic<bool>        go{false};
  for (int t = 0; t != nof_threads; ++t) {
    threads.emplace_back([&, t]() {
      // One warm-up TB outside the timed region (first-touch of the HARQ buffer).
      ref_pusch_bench(handles[t], tb_bytes, llrs, nof_llrs, bg, Qm, Nref, nof_layers, max_it, early_stop, 1, nullptr);
      ++ready;
      while (!go.load()) {
        std::this_thread::yield();
      }
      ref_pusch_bench(handles[t], tb_bytes, llrs, nof_llrs, bg, Qm, Nref, nof_layers, max_it, early_stop, reps, &oks[t]);
    });
  }
  while (ready.load() != nof_threads) {
    std::this_thread::yield();
  }
  auto t0 = std::chrono::steady_clock::now();
  go      = true;
  for (auto& th : threads) {
    th.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  int  ok = 0;
  for (int t = 0; t != nof_threads; ++t) {
    ok += oks[t];
    ref_pusch_destroy(handles[t]);
  }
  if (nof_crc_ok != nullptr) {
    *nof_crc_ok = ok;
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

/// Timing loop for BASELINE config 1: `reps` single-code-block decodes on one thread. Returns seconds.
double ref_ldpc_decode_bench(const char*   type,
                             const int8_t* in,
                             uint32_t      n_in,
                             int           bg,
                             int           Z,
                             int           crc_poly,
                             int           max_it,
                             int           reps)
{
  unsigned             K = ((bg == 1) ? 22 : 10) * Z;
  std::vector<uint8_t> out((K + 7) / 8);
  auto                 t0 = std::chrono::steady_clock::now();
  for (int r = 0; r != reps; ++r) {
    ref_ldpc_decode(type, out.data(), in, n_in, bg, Z, 0, crc_poly, max_it, 0.8F);
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
