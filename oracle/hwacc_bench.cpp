// MEASUREMENT HARNESS (test infrastructure): throughput and latency of the CUDA path THROUGH THE REFERENCE'S OWN PLUGIN
// INTERFACE (srsran::pusch_decoder, include/srsran/phy/upper/channel_processors/pusch/pusch_decoder.h:54-99), in the shape
// of the reference's tests/benchmarks/phy/upper/channel_processors/pusch/pusch_decoder_hwacc_benchmark.cpp:340-503:
// many pusch_decoder_cuda_impl instances (one transport block in flight each, like the pusch_processor pool) fed by a few
// host threads, completion through pusch_decoder_notifier on an executor; beside it the reference's software
// pusch_decoder_impl on all host cores with the same transport blocks. Built by oracle/Makefile (target hwacc) into
// oracle/_ref/hwacc_bench from the unmodified reference sources + srsran_projectvtlmo_b200/host/.
//
//   hwacc_bench --llrs FILE [--decoders 64] [--sets 3] [--slots 1000] [--threads 8] [--devices 1]
//               [--agg-tbs 64] [--agg-us 50] [--ref-seconds 5] [--check]
//
// FILE (written by tools/make_tb_file.py): header {magic, nof_tbs, tbs_bits, bg, qm, layers, nref, nof_llrs} as uint32, then
// per TB the payload bytes and the int8 soft bits. One "slot" = one transport block per decoder of a set; `sets` sets of
// decoders are in flight at a time (a decoder instance holds one transport block, pusch_decoder_impl.h:113-137). Prints one
// JSON line: Gbit/s of decoded info bits through the interface, per-TB latency (on_end_softbits -> on_sch_data) and slot
// latency (first on_end_softbits of a slot -> its last notification) percentiles, mean batch size of the slot aggregator,
// and the reference's number on the host cores. Exit code 0 ok, 1 a transport block failed or differed, 2 no CUDA device.
#include "pusch_decoder_cuda_impl.h"
#include "pusch_decoder_impl.h"
#include "srsran/phy/upper/channel_coding/channel_coding_factories.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_notifier.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_result.h"
#include "srsran/phy/upper/unique_rx_buffer.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace srsran;
using clk = std::chrono::steady_clock;

namespace {

/// rx buffer whose soft bits live on the accelerator (external_soft_bits): only CRC flags and ids on the host.
class external_rx_buffer : public unique_rx_buffer::callback
{
public:
  external_rx_buffer(unsigned nof_cbs_, unsigned first_id_) : nof_cbs(nof_cbs_), first_id(first_id_), crcs(new bool[nof_cbs_]()) {}
  unsigned   get_nof_codeblocks() const override { return nof_cbs; }
  void       reset_codeblocks_crc() override { std::fill(crcs.get(), crcs.get() + nof_cbs, false); }
  span<bool> get_codeblocks_crc() override { return span<bool>(crcs.get(), nof_cbs); }
  unsigned   get_absolute_codeblock_id(unsigned i) const override { return first_id + i; }
  span<log_likelihood_ratio> get_codeblock_soft_bits(unsigned, unsigned) override { return {}; }
  bit_buffer                 get_codeblock_data_bits(unsigned, unsigned n) override
  {
    return bit_buffer::from_bytes(span<uint8_t>(scratch)).first(std::min<unsigned>(n, scratch.size() * 8));
  }
  void lock() override {}
  void unlock() override {}
  void release() override {}
  unsigned                nof_cbs, first_id;
  std::unique_ptr<bool[]> crcs;
  std::vector<uint8_t>    scratch = std::vector<uint8_t>(ldpc::MAX_CODEBLOCK_SIZE / 8 + 8, 0); // never used by the CUDA decoder
};

/// Host-memory rx buffer for the reference's software decoder.
class host_rx_buffer : public unique_rx_buffer::callback
{
public:
  explicit host_rx_buffer(unsigned nof_cbs_) : nof_cbs(nof_cbs_), crcs(new bool[nof_cbs_]()), soft(nof_cbs_), data(nof_cbs_)
  {
    for (unsigned i = 0; i != nof_cbs; ++i) {
      soft[i].assign(ldpc::MAX_CODEBLOCK_SIZE, log_likelihood_ratio(0));
      data[i].assign(ldpc::MAX_CODEBLOCK_SIZE / 8 + 8, 0);
    }
  }
  unsigned   get_nof_codeblocks() const override { return nof_cbs; }
  void       reset_codeblocks_crc() override { std::fill(crcs.get(), crcs.get() + nof_cbs, false); }
  span<bool> get_codeblocks_crc() override { return span<bool>(crcs.get(), nof_cbs); }
  unsigned   get_absolute_codeblock_id(unsigned i) const override { return i; }
  span<log_likelihood_ratio> get_codeblock_soft_bits(unsigned i, unsigned n) override
  {
    return span<log_likelihood_ratio>(soft[i]).first(n);
  }
  bit_buffer get_codeblock_data_bits(unsigned i, unsigned n) override
  {
    return bit_buffer::from_bytes(span<uint8_t>(data[i])).first(n);
  }
  void lock() override {}
  void unlock() override {}
  void release() override {}
  unsigned                                       nof_cbs;
  std::unique_ptr<bool[]>                        crcs;
  std::vector<std::vector<log_likelihood_ratio>> soft;
  std::vector<std::vector<uint8_t>>              data;
};

/// Small worker pool for the completions (the reference gives its decoders a task_executor backed by the upper PHY's
/// pusch-decoder worker pool, nof_pusch_decoder_threads).
class queue_executor : public task_executor
{
public:
  explicit queue_executor(unsigned nof_workers)
  {
    for (unsigned i = 0; i != nof_workers; ++i) {
      workers.emplace_back([this]() { loop(); });
    }
  }
  ~queue_executor() override
  {
    {
      std::lock_guard<std::mutex> lock(m);
      stop = true;
    }
    cv.notify_all();
    for (auto& w : workers) {
      w.join();
    }
  }
  bool execute(unique_task task) override
  {
    {
      std::lock_guard<std::mutex> lock(m);
      q.push_back(std::move(task));
    }
    cv.notify_one();
    return true;
  }
  bool defer(unique_task task) override { return execute(std::move(task)); }

private:
  void loop()
  {
    for (;;) {
      unique_task t;
      {
        std::unique_lock<std::mutex> lock(m);
        cv.wait(lock, [this]() { return stop || !q.empty(); });
        if (q.empty()) {
          return;
        }
        t = std::move(q.front());
        q.pop_front();
      }
      t();
    }
  }
  std::mutex              m;
  std::condition_variable cv;
  std::deque<unique_task> q;
  bool                    stop = false;
  std::vector<std::thread> workers;
};

struct slot_state;

/// One decoder instance with its rx buffer, output and notifier.
struct lane : public pusch_decoder_notifier {
  std::unique_ptr<pusch_decoder>      decoder;
  std::unique_ptr<external_rx_buffer> rxbuf;
  std::vector<uint8_t>                tb;
  unsigned                            tb_index = 0; // which transport block of the file it decodes
  std::atomic<bool>                   busy{false};
  bool                                staged = false; // the soft bits already sit in the decoder's page-locked staging
  clk::time_point                     t_end;
  slot_state*                         slot = nullptr;
  std::vector<double>*                lat  = nullptr;
  std::mutex*                         lat_mutex = nullptr;
  std::atomic<unsigned>*              failures  = nullptr;
  const std::vector<uint8_t>*         payload   = nullptr;
  bool                                check     = false;
  void on_sch_data(const pusch_decoder_result& result) override;
};

struct slot_state {
  std::atomic<unsigned> pending{0};
  clk::time_point       t_first;
  std::atomic<bool>     first_set{false};
  double                latency_us = 0;
};

void lane::on_sch_data(const pusch_decoder_result& result)
{
  const auto now = clk::now();
  double     us  = std::chrono::duration<double, std::micro>(now - t_end).count();
  if (!result.tb_crc_ok || (check && std::memcmp(tb.data(), payload->data(), tb.size()) != 0)) {
    failures->fetch_add(1);
  }
  {
    std::lock_guard<std::mutex> lock(*lat_mutex);
    lat->push_back(us);
  }
  slot_state* s = slot;
  busy.store(false, std::memory_order_release);
  if (s->pending.fetch_sub(1) == 1) {
    s->latency_us = std::chrono::duration<double, std::micro>(now - s->t_first).count();
  }
}

double pct(std::vector<double>& v, double p)
{
  if (v.empty()) {
    return 0;
  }
  std::sort(v.begin(), v.end());
  return v[std::min(v.size() - 1, static_cast<size_t>(p * v.size()))];
}

} // namespace

int main(int argc, char** argv)
{
  std::string file;
  unsigned    nof_decoders = 64, nof_sets = 3, nof_slots = 1000, nof_threads = 8, nof_devices = 1, agg_tbs = 64, agg_us = 50, nof_workers = 4;
  double      ref_seconds = 5;
  bool        check       = false;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto        next = [&]() { return (i + 1 < argc) ? std::string(argv[++i]) : std::string(); };
    if (a == "--llrs") file = next();
    else if (a == "--decoders") nof_decoders = std::stoul(next());
    else if (a == "--sets") nof_sets = std::stoul(next());
    else if (a == "--slots") nof_slots = std::stoul(next());
    else if (a == "--threads") nof_threads = std::stoul(next());
    else if (a == "--devices") nof_devices = std::stoul(next());
    else if (a == "--agg-tbs") agg_tbs = std::stoul(next());
    else if (a == "--agg-us") agg_us = std::stoul(next());
    else if (a == "--workers") nof_workers = std::stoul(next());
    else if (a == "--ref-seconds") ref_seconds = std::stod(next());
    else if (a == "--check") check = true;
  }
  FILE* f = file.empty() ? nullptr : std::fopen(file.c_str(), "rb");
  if (f == nullptr) {
    std::fprintf(stderr, "hwacc_bench: --llrs FILE is required (tools/make_tb_file.py writes it)\n");
    return 1;
  }
  uint32_t hdr[8];
  if (std::fread(hdr, 4, 8, f) != 8 || hdr[0] != 0x50425443U) {
    std::fprintf(stderr, "hwacc_bench: bad file header\n");
    return 1;
  }
  const unsigned nof_tbs = hdr[1], tbs_bits = hdr[2], bg = hdr[3], qm = hdr[4], layers = hdr[5], nref = hdr[6], nof_llrs = hdr[7];
  std::vector<std::vector<uint8_t>> payloads(nof_tbs, std::vector<uint8_t>(tbs_bits / 8));
  std::vector<std::vector<int8_t>>  llrs(nof_tbs, std::vector<int8_t>(nof_llrs));
  for (unsigned i = 0; i != nof_tbs; ++i) {
    if (std::fread(payloads[i].data(), 1, payloads[i].size(), f) != payloads[i].size() ||
        std::fread(llrs[i].data(), 1, nof_llrs, f) != nof_llrs) {
      std::fprintf(stderr, "hwacc_bench: short file\n");
      return 1;
    }
  }
  std::fclose(f);
  const ldpc_base_graph_type bgt     = (bg == 1) ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
  const unsigned             nof_cbs = ldpc::compute_nof_codeblocks(units::bits(tbs_bits), bgt);

  // ---- devices, decoders ---------------------------------------------------------------------------------------------------
  const unsigned nof_lanes = nof_decoders * nof_sets;
  std::vector<std::shared_ptr<hal::cuda_pusch_dec_device>> devices;
  try {
    for (unsigned d = 0; d != nof_devices; ++d) {
      hal::cuda_hwacc_pusch_dec_configuration cfg;
      cfg.device            = static_cast<int>(d);
      cfg.max_cbs_in_flight = agg_tbs * nof_cbs;
      cfg.nof_harq_cb_slots = (nof_lanes / nof_devices + 1) * (nof_cbs + nof_devices) + nof_devices;
      devices.push_back(std::make_shared<hal::cuda_pusch_dec_device>(cfg));
      devices.back()->set_aggregation(agg_tbs, std::chrono::microseconds(agg_us));
    }
  } catch (const std::exception& e) {
    std::printf("{\"hwacc_bench\": \"no usable CUDA device\", \"error\": \"%s\"}\n", e.what());
    return 2;
  }
  queue_executor notify_executor(nof_workers);
  auto           factory = create_pusch_decoder_factory_cuda(devices, &notify_executor, MAX_RB, 4);
  std::vector<std::unique_ptr<lane>> lanes(nof_lanes);
  std::vector<double>                latencies;
  std::mutex                         lat_mutex;
  std::atomic<unsigned>              failures{0};
  latencies.reserve(static_cast<size_t>(nof_slots) * nof_decoders);
  std::vector<unsigned> next_id(nof_devices, 0);
  for (unsigned i = 0; i != nof_lanes; ++i) {
    auto l      = std::make_unique<lane>();
    l->decoder  = factory->create();
    // Sticky sharding: the first code-block id decides the device (id % nof_devices); ids of a device are packed.
    unsigned dev = i % nof_devices;
    unsigned id0 = next_id[dev] + (nof_devices + dev - next_id[dev] % nof_devices) % nof_devices; // id0 % nof_devices == dev
    next_id[dev] = id0 + nof_cbs; // the code blocks of the TB take the consecutive HARQ slots id0 .. id0 + nof_cbs - 1 there
    l->rxbuf     = std::make_unique<external_rx_buffer>(nof_cbs, id0);
    l->tb.resize(tbs_bits / 8);
    l->tb_index  = i % nof_tbs;
    l->lat       = &latencies;
    l->lat_mutex = &lat_mutex;
    l->failures  = &failures;
    l->payload   = &payloads[l->tb_index];
    l->check     = check;
    lanes[i]     = std::move(l);
  }

  pusch_decoder::configuration cfg;
  cfg.base_graph          = bgt;
  cfg.rv                  = 0;
  cfg.mod                 = static_cast<modulation_scheme>(qm);
  cfg.Nref                = nref;
  cfg.nof_layers          = layers;
  cfg.nof_ldpc_iterations = 6;
  cfg.use_early_stop      = true;
  cfg.new_data            = true;

  // ---- slots ---------------------------------------------------------------------------------------------------------------
  std::vector<slot_state> slots(nof_slots);
  auto                    feed = [&](unsigned thread_id, unsigned slot_begin, unsigned slot_end) {
    for (unsigned s = slot_begin; s != slot_end; ++s) {
      const unsigned set = s % nof_sets;
      for (unsigned k = thread_id; k < nof_decoders; k += nof_threads) {
        lane& l = *lanes[set * nof_decoders + k];
        while (l.busy.load(std::memory_order_acquire)) {
          std::this_thread::yield(); // the decoder still holds the transport block of slot s - nof_sets
        }
        l.busy.store(true);
        l.slot = &slots[s];
        l.rxbuf->reset_codeblocks_crc(); // rx_buffer_pool_impl::reserve does this for new data
        pusch_decoder_buffer& buf = l.decoder->new_data(span<uint8_t>(l.tb), unique_rx_buffer(*l.rxbuf), l, cfg);
        // The demodulator writes the soft bits through the decoder's view (ulsch_demultiplex_impl.cpp:253-262); here they
        // are produced once and stay in the page-locked staging buffer behind that view.
        span<log_likelihood_ratio> view = buf.get_next_block_view(nof_llrs);
        if (!l.staged) {
          std::memcpy(view.data(), llrs[l.tb_index].data(), nof_llrs);
          l.staged = true;
        }
        buf.on_new_softbits(view);
        l.t_end = clk::now();
        bool expected = false;
        if (slots[s].first_set.compare_exchange_strong(expected, true)) {
          slots[s].t_first = l.t_end;
        }
        buf.on_end_softbits();
      }
    }
  };
  for (slot_state& s : slots) {
    s.pending.store(nof_decoders);
  }
  // Warm-up: a few slots outside the timed region (buffers, first-use allocations).
  const unsigned warm = std::min(nof_slots / 10 + nof_sets, nof_slots / 2);
  auto           run_range = [&](unsigned b, unsigned e) {
    std::vector<std::thread> th;
    for (unsigned t = 0; t != nof_threads; ++t) {
      th.emplace_back(feed, t, b, e);
    }
    for (auto& x : th) {
      x.join();
    }
    for (auto& l : lanes) {
      while (l->busy.load(std::memory_order_acquire)) {
        std::this_thread::yield();
      }
    }
  };
  run_range(0, warm);
  {
    std::lock_guard<std::mutex> lock(lat_mutex);
    latencies.clear();
  }
  for (auto& d : devices) {
    d->times.submit_us = d->times.busy_wait_us = d->times.device_wait_us = d->times.collect_us = d->times.callbacks_us =
        d->times.consume_us                                                                   = 0;
  }
  auto t0 = clk::now();
  run_range(warm, nof_slots);
  double secs = std::chrono::duration<double>(clk::now() - t0).count();
  const unsigned timed_slots = nof_slots - warm;
  double gbps = static_cast<double>(timed_slots) * nof_decoders * tbs_bits / secs / 1e9;
  std::vector<double> slot_lat;
  for (unsigned s = warm; s != nof_slots; ++s) {
    slot_lat.push_back(slots[s].latency_us);
  }
  uint64_t batches = 0, batched = 0;
  for (auto& d : devices) {
    auto st = d->aggregation_stats();
    batches += st.first;
    batched += st.second;
    std::fprintf(stderr,
                 "device threads (ms over the whole run): flusher submit %.1f, busy wait %.1f | completer device wait %.1f, "
                 "collect %.1f, callbacks %.1f, consume %.1f | timed region %.1f ms\n",
                 d->times.submit_us / 1e3, d->times.busy_wait_us / 1e3, d->times.device_wait_us / 1e3, d->times.collect_us / 1e3,
                 d->times.callbacks_us / 1e3, d->times.consume_us / 1e3, secs * 1e3);
  }

  // ---- reference software decoder on all host cores ------------------------------------------------------------------------
  double   ref_gbps    = 0;
  unsigned ref_threads = std::max(1U, std::thread::hardware_concurrency());
  if (ref_seconds > 0) {
    auto crc_factory = create_crc_calculator_factory_sw("auto");
    auto seg_factory = create_ldpc_segmenter_rx_factory_sw();
    auto dec_factory = create_ldpc_decoder_factory_sw("auto");
    auto dem_factory = create_ldpc_rate_dematcher_factory_sw("auto");
    std::atomic<uint64_t> done_tbs{0};
    std::atomic<bool>     stop{false};
    std::vector<std::thread> th;
    for (unsigned t = 0; t != ref_threads; ++t) {
      th.emplace_back([&, t]() {
        std::vector<std::unique_ptr<pusch_codeblock_decoder>> cbd(1);
        pusch_codeblock_decoder::sch_crc crcs;
        crcs.crc16  = crc_factory->create(crc_generator_poly::CRC16);
        crcs.crc24A = crc_factory->create(crc_generator_poly::CRC24A);
        crcs.crc24B = crc_factory->create(crc_generator_poly::CRC24B);
        cbd[0]      = std::make_unique<pusch_codeblock_decoder>(dem_factory->create(), dec_factory->create(), crcs);
        auto pool   = std::make_shared<pusch_decoder_impl::codeblock_decoder_pool>(std::move(cbd));
        pusch_decoder_impl::sch_crc sw_crcs;
        sw_crcs.crc16  = crc_factory->create(crc_generator_poly::CRC16);
        sw_crcs.crc24A = crc_factory->create(crc_generator_poly::CRC24A);
        sw_crcs.crc24B = crc_factory->create(crc_generator_poly::CRC24B);
        pusch_decoder_impl   dec(seg_factory->create(), pool, std::move(sw_crcs), nullptr, MAX_RB, 4);
        host_rx_buffer       buf(nof_cbs);
        std::vector<uint8_t> out(tbs_bits / 8);
        struct nt : pusch_decoder_notifier {
          void on_sch_data(const pusch_decoder_result&) override {}
        } n;
        const std::vector<int8_t>& in = llrs[t % nof_tbs];
        while (!stop.load()) {
          buf.reset_codeblocks_crc();
          pusch_decoder_buffer& b = dec.new_data(span<uint8_t>(out), unique_rx_buffer(buf), n, cfg);
          b.on_new_softbits(span<const log_likelihood_ratio>(reinterpret_cast<const log_likelihood_ratio*>(in.data()), in.size()));
          b.on_end_softbits();
          done_tbs.fetch_add(1);
        }
      });
    }
    std::this_thread::sleep_for(std::chrono::duration<double>(0.5)); // first TB of every thread: page faults
    uint64_t n0 = done_tbs.load();
    auto     r0 = clk::now();
    std::this_thread::sleep_for(std::chrono::duration<double>(ref_seconds));
    uint64_t n1 = done_tbs.load();
    double   rs = std::chrono::duration<double>(clk::now() - r0).count();
    stop.store(true);
    for (auto& x : th) {
      x.join();
    }
    ref_gbps = static_cast<double>(n1 - n0) * tbs_bits / rs / 1e9;
  }

  std::printf("{\"bench\": \"pusch_decoder plugin interface (C++)\", \"value\": %.3f, \"unit\": \"Gbit/s\", \"n_gpus\": %u, "
              "\"decoders\": %u, \"sets_in_flight\": %u, \"feeder_threads\": %u, \"completion_workers\": %u, \"slots\": %u, \"tbs_per_slot\": %u, "
              "\"tbs_bits\": %u, \"codeblocks_per_tb\": %u, \"tb_latency_us\": {\"p50\": %.1f, \"p99\": %.1f, \"max\": %.1f, \"n\": %zu}, "
              "\"slot_latency_us\": {\"p50\": %.1f, \"p99\": %.1f, \"max\": %.1f, \"n\": %zu}, "
              "\"aggregator\": {\"max_tbs\": %u, \"deadline_us\": %u, \"batches\": %llu, \"mean_tbs_per_batch\": %.2f}, "
              "\"failed_or_wrong_tbs\": %u, \"payload_checked\": %s, "
              "\"reference\": {\"value\": %.3f, \"unit\": \"Gbit/s\", \"cores\": %u, \"what\": \"reference pusch_decoder_impl, one per host thread, same transport blocks\"}}\n",
              gbps, nof_devices, nof_decoders, nof_sets, nof_threads, nof_workers, timed_slots, nof_decoders, tbs_bits, nof_cbs,
              pct(latencies, 0.5), pct(latencies, 0.99), latencies.empty() ? 0.0 : *std::max_element(latencies.begin(), latencies.end()),
              latencies.size(), pct(slot_lat, 0.5), pct(slot_lat, 0.99),
              slot_lat.empty() ? 0.0 : *std::max_element(slot_lat.begin(), slot_lat.end()), slot_lat.size(), agg_tbs, agg_us,
              static_cast<unsigned long long>(batches), batches ? static_cast<double>(batched) / batches : 0.0, failures.load(),
              check ? "true" : "false", ref_gbps, ref_threads);
  return failures.load() ? 1 : 0;
}
