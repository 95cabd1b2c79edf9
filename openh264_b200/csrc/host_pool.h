// host_pool.h — minimal fork-join thread pool for the per-stream host jobs (CAVLC writing in the encoder, Annex-B / CAVLC
// parsing in the decoder) and the number of cores this process may actually use.
#pragma once
#include <sched.h>
#include <stdio.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace b2h264 {

// cores usable by this process: the affinity mask, capped by the cgroup CPU quota (containers expose all host cores through
// hardware_concurrency but schedule only a few of them)
inline int usable_cores() {
  int n = (int)std::thread::hardware_concurrency();
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) { const int c = CPU_COUNT(&set); if (c > 0) n = c; }
  if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
    long long quota = 0, period = 0;
    char q[32] = {0};
    if (fscanf(f, "%31s %lld", q, &period) == 2 && q[0] != 'm' && period > 0 && sscanf(q, "%lld", &quota) == 1 && quota > 0) {
      const int c = (int)((quota + period - 1) / period);
      if (c > 0 && c < n) n = c;
    }
    fclose(f);
  }
  return n < 1 ? 1 : n;
}

class Pool {
 public:
  explicit Pool(int n) {
    for (int i = 0; i < n; i++) th_.emplace_back([this] { loop(); });
  }
  ~Pool() {
    { std::lock_guard<std::mutex> l(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  void run(int n, const std::function<void(int)>& fn) {
    { std::lock_guard<std::mutex> l(m_); fn_ = &fn; next_ = 0; total_ = n; done_ = 0; }
    cv_.notify_all();
    std::unique_lock<std::mutex> l(m_);
    done_cv_.wait(l, [&] { return done_ == total_; });
    fn_ = nullptr;
  }
 private:
  void loop() {
    for (;;) {
      int job;
      const std::function<void(int)>* fn;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return stop_ || (fn_ && next_ < total_); });
        if (stop_) return;
        job = next_++; fn = fn_;
      }
      (*fn)(job);
      { std::lock_guard<std::mutex> l(m_); if (++done_ == total_) done_cv_.notify_all(); }
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int)>* fn_ = nullptr;
  int next_ = 0, total_ = 0, done_ = 0;
  bool stop_ = false;
};

}  // namespace b2h264
