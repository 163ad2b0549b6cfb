#include "cuda_pusch_dec_device.h"
#include <cstdio>
#include <cstdlib>

using namespace srsran;
using namespace hal;

namespace {
using sclk = std::chrono::steady_clock;
uint64_t us_since(sclk::time_point t0)
{
  return static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::microseconds>(sclk::now() - t0).count());
}
} // namespace

cuda_pusch_dec_device::cuda_pusch_dec_device(const cuda_hwacc_pusch_dec_configuration& cfg)
{
  int st = srsran_cuda_pusch_dec_create(cfg.device, cfg.max_cbs_in_flight, cfg.nof_harq_cb_slots, &handle);
  if (st != SRSRAN_CUDA_OK) {
    // No CPU fallback: the caller gets a null factory (like the ACC100 factory without DPDK).
    throw std::runtime_error(std::string("srsran_cuda_pusch_dec_create failed: ") + srsran_cuda_pusch_dec_last_error(nullptr));
  }
}

cuda_pusch_dec_device::~cuda_pusch_dec_device()
{
  {
    std::lock_guard<std::mutex> lock(agg_mtx);
    stop = true;
  }
  cv_pending.notify_all();
  cv_flying.notify_all();
  cv_progress.notify_all();
  if (flusher.joinable()) {
    flusher.join();
  }
  if (completer.joinable()) {
    completer.join();
  }
  srsran_cuda_pusch_dec_destroy(handle);
}

void cuda_pusch_dec_device::set_aggregation(unsigned max_tbs_per_batch, std::chrono::microseconds flush_deadline)
{
  std::lock_guard<std::mutex> lock(agg_mtx);
  max_batch = std::max(1U, std::min(max_tbs_per_batch, 1024U));
  deadline  = flush_deadline;
}

void cuda_pusch_dec_device::start_threads()
{
  if (started) {
    return;
  }
  started   = true;
  flusher   = std::thread([this]() { flusher_loop(); });
  completer = std::thread([this]() { completer_loop(); });
}

void cuda_pusch_dec_device::submit(const cuda_tb_request& request, completion_fn on_done, task_executor* executor)
{
  {
    std::lock_guard<std::mutex> lock(agg_mtx);
    start_threads();
    pending.push_back({request, std::move(on_done), executor, std::chrono::steady_clock::now()});
  }
  cv_pending.notify_one();
}

void cuda_pusch_dec_device::flusher_loop()
{
  std::vector<pending_tb>                      batch;
  std::vector<srsran_cuda_pusch_dec_tb_config> cfgs;
  std::vector<const int8_t*>                   llrs;
  std::vector<uint32_t>                        nof_llrs, cb_ids, nof_ids;
  std::vector<int>                             ingest, tickets;
  for (;;) {
    {
      std::unique_lock<std::mutex> lock(agg_mtx);
      cv_pending.wait(lock, [this]() { return stop || !pending.empty(); });
      if (stop && pending.empty()) {
        return;
      }
      // Collect until the batch is full or its oldest transport block has waited long enough.
      const auto due = pending.front().arrival + deadline;
      cv_pending.wait_until(lock, due, [this]() { return stop || pending.size() >= max_batch; });
      batch.clear();
      while (!pending.empty() && batch.size() < max_batch) {
        batch.push_back(std::move(pending.front()));
        pending.pop_front();
      }
    }
    const uint32_t n = static_cast<uint32_t>(batch.size());
    cfgs.resize(n);
    llrs.resize(n);
    nof_llrs.resize(n);
    ingest.resize(n);
    nof_ids.resize(n);
    tickets.assign(n, -1);
    cb_ids.clear();
    for (uint32_t i = 0; i != n; ++i) {
      const cuda_tb_request& r = batch[i].req;
      cfgs[i]                  = r.cfg;
      llrs[i]                  = r.llrs;
      nof_llrs[i]              = r.nof_llrs;
      ingest[i]                = r.ingest_stream;
      nof_ids[i]               = r.nof_cbs;
      cb_ids.insert(cb_ids.end(), r.cb_ids, r.cb_ids + r.nof_cbs);
    }
    for (;;) {
      int st;
      {
        const auto                            t0 = sclk::now();
        std::lock_guard<std::recursive_mutex> lock(mtx);
        st = srsran_cuda_pusch_dec_submit_tbs_cb_ids(
            handle, n, cfgs.data(), llrs.data(), nof_llrs.data(), ingest.data(), cb_ids.data(), nof_ids.data(), tickets.data());
        times.submit_us += us_since(t0);
      }
      if (st == SRSRAN_CUDA_OK) {
        break;
      }
      if (st == SRSRAN_CUDA_ERR_BUSY) {
        // Every batch context is in flight and unpolled: wait for the completion thread to consume one.
        const auto                   t0 = sclk::now();
        std::unique_lock<std::mutex> lock(agg_mtx);
        if (stop) {
          return;
        }
        cv_progress.wait_for(lock, std::chrono::microseconds(200));
        times.busy_wait_us += us_since(t0);
        continue;
      }
      std::fprintf(stderr, "CUDA PUSCH decoder: batch submission failed (%d): %s\n", st, srsran_cuda_pusch_dec_last_error(handle));
      std::abort();
    }
    nof_batches.fetch_add(1);
    nof_batched_tbs.fetch_add(n);
    auto fb     = std::make_shared<flying_batch>();
    fb->tickets = tickets;
    for (uint32_t i = 0; i != n; ++i) {
      fb->callbacks.push_back(std::move(batch[i].on_done));
      fb->executors.push_back(batch[i].executor);
      fb->nof_cbs.push_back(batch[i].req.nof_cbs);
    }
    {
      std::lock_guard<std::mutex> lock(agg_mtx);
      flying.push_back(std::move(fb));
    }
    cv_flying.notify_one();
  }
}

void cuda_pusch_dec_device::consume(flying_batch& fb)
{
  // Every callback of the batch has copied its transport block out: the tickets are consumed (one call), which lets a later
  // submission reuse the batch context and its result buffer.
  const auto t0 = sclk::now();
  {
    std::lock_guard<std::recursive_mutex> lock(mtx);
    srsran_cuda_pusch_dec_poll_tbs(handle, static_cast<uint32_t>(fb.tickets.size()), fb.tickets.data(), 0, nullptr, nullptr);
  }
  times.consume_us += us_since(t0);
  cv_progress.notify_all();
}

void cuda_pusch_dec_device::completer_loop()
{
  for (;;) {
    std::shared_ptr<flying_batch> fbp;
    {
      std::unique_lock<std::mutex> lock(agg_mtx);
      cv_flying.wait(lock, [this]() { return stop || !flying.empty(); });
      if (flying.empty()) {
        return;
      }
      fbp = std::move(flying.front());
      flying.pop_front();
    }
    flying_batch& fb = *fbp;
    // Wait on the device without holding the handle: submissions of the next batches go on meanwhile.
    auto t0 = sclk::now();
    int  st = srsran_cuda_pusch_dec_wait_ticket(handle, fb.tickets.front());
    times.device_wait_us += us_since(t0);
    t0 = sclk::now();
    if (st != SRSRAN_CUDA_OK) {
      std::fprintf(stderr, "CUDA PUSCH decoder: waiting for a batch failed (%d)\n", st);
      std::abort();
    }
    fb.done.resize(fb.tickets.size());
    {
      std::lock_guard<std::recursive_mutex> lock(mtx);
      for (size_t i = 0; i != fb.tickets.size(); ++i) {
        cuda_tb_completion& c = fb.done[i];
        st = srsran_cuda_pusch_dec_tb_cb_outputs(handle, fb.tickets[i], c.cb_crc, c.cb_iterations, fb.nof_cbs[i]);
        if (st >= 0) {
          st = srsran_cuda_pusch_dec_tb_data(handle, fb.tickets[i], &c.tb_data);
        }
        if (st >= 0) {
          st = srsran_cuda_pusch_dec_peek_tb(handle, fb.tickets[i], &c.result);
        }
        if (st < 0) {
          std::fprintf(stderr, "CUDA PUSCH decoder: collecting a result failed (%d): %s\n", st, srsran_cuda_pusch_dec_last_error(handle));
          std::abort();
        }
      }
    }
    times.collect_us += us_since(t0);
    t0 = sclk::now();
    // The callbacks copy the transport blocks out of the batch's page-locked result buffer - on the executors' threads where
    // the decoders have one (in parallel), else here. The last one to finish consumes the tickets.
    fb.remaining.store(static_cast<unsigned>(fb.tickets.size()));
    for (size_t i = 0; i != fb.tickets.size(); ++i) {
      auto run = [this, fbp, i]() {
        fbp->callbacks[i](fbp->done[i]);
        if (fbp->remaining.fetch_sub(1) == 1) {
          consume(*fbp);
        }
      };
      if (fb.executors[i] == nullptr || !fb.executors[i]->execute(run)) {
        run();
      }
    }
    times.callbacks_us += us_since(t0);
  }
}
