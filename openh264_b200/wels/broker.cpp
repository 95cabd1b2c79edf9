// broker.cpp — see broker.h
#include "broker.h"

#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace b2wels {

static long env_long(const char* name, long dflt) {
  const char* v = getenv(name);
  return v && *v ? atol(v) : dflt;
}

Pool::Pool(const PoolKey& key, int capacity, int device) : key_(key), cap_(capacity), device_(device) {
  frame_bytes_ = (size_t)key.width * key.height * 3 / 2;
  wait_us_ = env_long("B2H264_BROKER_WAIT_US", 2000);
  b2h264_enc_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.width = key.width; cfg.height = key.height; cfg.qp = key.qp; cfg.fps = key.fps;
  cfg.target_bitrate = key.bitrate;
  cfg.n_streams = capacity;
  cfg.entropy_threads = 0;
  cfg.device = device;
  cfg.sps_pps_id_strategy = key.strategy;
  cfg.complexity_low = key.complexity_low;
  cfg.entropy_cabac = key.entropy_cabac;
  cfg.profile_idc = key.profile_idc;
  cfg.intra_period = key.intra_period;
  cfg.loop_filter_idc = key.dbk_idc; cfg.loop_filter_alpha_c0_offset = key.dbk_alpha; cfg.loop_filter_beta_offset = key.dbk_beta;
  if (b2h264_enc_create(&cfg, &enc_) != 0) { enc_ = nullptr; return; }
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaHostAlloc((void**)&pinned_, frame_bytes_ * capacity, cudaHostAllocPortable) != cudaSuccess) {
    b2h264_enc_destroy(enc_);
    enc_ = nullptr; pinned_ = nullptr;
    return;
  }
  cudaStream_t us = nullptr;
  if (cudaMalloc((void**)&d_stage_, frame_bytes_ * capacity) != cudaSuccess ||
      cudaStreamCreateWithFlags(&us, cudaStreamNonBlocking) != cudaSuccess) {
    b2h264_enc_destroy(enc_); cudaFreeHost(pinned_); if (d_stage_) cudaFree(d_stage_);
    enc_ = nullptr; pinned_ = nullptr; d_stage_ = nullptr;
    return;
  }
  up_stream_ = us;
  state_.assign(capacity, FREE);
  au_.resize(capacity);
  idr_.assign(capacity, 0);
}

Pool::~Pool() {
  if (enc_) b2h264_enc_destroy(enc_);
  if (pinned_) cudaFreeHost(pinned_);
  if (d_stage_) { cudaSetDevice(device_); cudaFree(d_stage_); }
  if (up_stream_) cudaStreamDestroy((cudaStream_t)up_stream_);
}

int Pool::upload(int slot) {
  if (cudaSetDevice(device_) != cudaSuccess) return -1;
  return cudaMemcpyAsync(d_stage_ + (size_t)slot * frame_bytes_, pinned_ + (size_t)slot * frame_bytes_, frame_bytes_, cudaMemcpyHostToDevice,
                         (cudaStream_t)up_stream_) == cudaSuccess ? 0 : -1;
}

int Pool::acquire() {
  std::unique_lock<std::mutex> lk(m_);
  for (int s = 0; s < cap_; s++)
    if (state_[s] == FREE) { state_[s] = IDLE; n_registered_++; return s; }
  return -1;
}

void Pool::release(int slot) {
  std::unique_lock<std::mutex> lk(m_);
  cv_.wait(lk, [&] { return !flushing_; });            // nothing of this pool is in flight now
  b2h264_enc_reset_stream(enc_, slot);
  state_[slot] = FREE;
  n_registered_--;
  cv_.notify_all();                                    // "everybody is waiting" may have become true
}

int Pool::registered() {
  std::unique_lock<std::mutex> lk(m_);
  return n_registered_;
}

int Pool::force_idr(int slot) {
  std::unique_lock<std::mutex> lk(m_);
  cv_.wait(lk, [&] { return !flushing_; });
  return b2h264_enc_force_idr(enc_, slot);
}

// codes every PENDING picture as one batch; called with the lock held, returns with it held
void Pool::flush_locked(std::unique_lock<std::mutex>& lk) {
  flushing_ = true;
  std::vector<const uint8_t*> src(cap_, nullptr);
  for (int s = 0; s < cap_; s++)
    if (state_[s] == PENDING) { state_[s] = INFLIGHT; src[s] = d_stage_ + (size_t)s * frame_bytes_; }
  n_pending_ = 0;
  lk.unlock();
  std::vector<const uint8_t*> bs(cap_, nullptr);
  std::vector<int32_t> nb(cap_, 0), ft(cap_, 0);
  // every pending stream enqueued its upload before it registered as pending: one wait covers them all
  int rc = cudaSetDevice(device_) == cudaSuccess && cudaStreamSynchronize((cudaStream_t)up_stream_) == cudaSuccess ? 0 : -1;
  if (rc == 0) rc = b2h264_enc_submit(enc_, src.data(), 1);
  if (rc == 0) rc = b2h264_enc_collect(enc_, bs.data(), nb.data(), ft.data());
  lk.lock();
  for (int s = 0; s < cap_; s++) {
    if (state_[s] != INFLIGHT) continue;
    if (rc == 0 && bs[s]) { au_[s].assign(bs[s], bs[s] + nb[s]); idr_[s] = ft[s] == 1; state_[s] = DONE; }
    else state_[s] = FAILED;
  }
  last_rc_ = rc;
  flushing_ = false;
  cv_.notify_all();
}

int Pool::encode(int slot, std::vector<uint8_t>* au, bool* idr) {
  std::unique_lock<std::mutex> lk(m_);
  if (slot < 0 || slot >= cap_ || state_[slot] != IDLE) return -1;
  state_[slot] = PENDING;
  n_pending_++;
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(wait_us_);
  while (state_[slot] != DONE && state_[slot] != FAILED) {
    if (state_[slot] == PENDING && !flushing_) {
      if (n_pending_ >= n_registered_ || std::chrono::steady_clock::now() >= deadline) { flush_locked(lk); continue; }
      cv_.wait_until(lk, deadline);
    } else {
      cv_.wait(lk);
    }
  }
  const bool ok = state_[slot] == DONE;
  if (ok) { au->swap(au_[slot]); *idr = idr_[slot] != 0; }
  state_[slot] = IDLE;
  return ok ? 0 : (last_rc_ ? last_rc_ : -1);
}

// ---------------------------------------------------------------------------------------------------------------------
DecPool::DecPool(int width, int height, int capacity, int device) : w_(width), h_(height), cap_(capacity), device_(device) {
  frame_bytes_ = (size_t)width * height * 3 / 2;
  wait_us_ = env_long("B2H264_BROKER_WAIT_US", 2000);
  b2h264_dec_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.width = width; cfg.height = height; cfg.n_streams = capacity; cfg.device = device;
  if (b2h264_dec_create(&cfg, &dec_) != 0) { dec_ = nullptr; return; }
  pinned_ = static_cast<uint8_t*>(b2h264_host_alloc(frame_bytes_ * capacity));
  if (!pinned_) { b2h264_dec_destroy(dec_); dec_ = nullptr; return; }
  state_.assign(capacity, FREE);
  au_.assign(capacity, nullptr);
  bytes_.assign(capacity, 0);
  status_.assign(capacity, 0);
}

DecPool::~DecPool() {
  if (dec_) b2h264_dec_destroy(dec_);
  if (pinned_) b2h264_host_free(pinned_);
}

int DecPool::acquire() {
  std::unique_lock<std::mutex> lk(m_);
  for (int s = 0; s < cap_; s++)
    if (state_[s] == FREE) { state_[s] = IDLE; n_registered_++; return s; }
  return -1;
}

void DecPool::release(int slot) {
  std::unique_lock<std::mutex> lk(m_);
  cv_.wait(lk, [&] { return !flushing_; });
  b2h264_dec_reset_stream(dec_, slot);
  state_[slot] = FREE;
  n_registered_--;
  cv_.notify_all();
}

int DecPool::registered() {
  std::unique_lock<std::mutex> lk(m_);
  return n_registered_;
}

void DecPool::flush_locked(std::unique_lock<std::mutex>& lk) {
  flushing_ = true;
  std::vector<const uint8_t*> au(cap_, nullptr);
  std::vector<int32_t> nb(cap_, 0), st(cap_, 0);
  std::vector<uint8_t*> out(cap_, nullptr);
  for (int s = 0; s < cap_; s++) {
    out[s] = pinned_ + (size_t)s * frame_bytes_;
    if (state_[s] == PENDING) { state_[s] = INFLIGHT; au[s] = au_[s]; nb[s] = bytes_[s]; }
  }
  n_pending_ = 0;
  lk.unlock();
  const int rc = b2h264_dec_decode3(dec_, au.data(), nb.data(), out.data(), st.data());
  lk.lock();
  for (int s = 0; s < cap_; s++) {
    if (state_[s] != INFLIGHT) continue;
    status_[s] = rc ? (rc > 0 ? -1000 - rc : rc) : st[s];        // a failure of the call as a whole fails every stream of the batch
    state_[s] = DONE;
  }
  flushing_ = false;
  cv_.notify_all();
}

int DecPool::decode(int slot, const uint8_t* au, int32_t bytes) {
  std::unique_lock<std::mutex> lk(m_);
  if (slot < 0 || slot >= cap_ || state_[slot] != IDLE) return -1;
  au_[slot] = au; bytes_[slot] = bytes;
  state_[slot] = PENDING;
  n_pending_++;
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(wait_us_);
  while (state_[slot] != DONE) {
    if (state_[slot] == PENDING && !flushing_) {
      if (n_pending_ >= n_registered_ || std::chrono::steady_clock::now() >= deadline) { flush_locked(lk); continue; }
      cv_.wait_until(lk, deadline);
    } else {
      cv_.wait(lk);
    }
  }
  const int st = status_[slot];
  state_[slot] = IDLE;
  return st;
}

// ---------------------------------------------------------------------------------------------------------------------
Broker& Broker::get() {
  static Broker* b = new Broker();       // never destroyed: encoder objects may outlive static destruction order
  return *b;
}

std::shared_ptr<Pool> Broker::attach(const PoolKey& key, int* slot) {
  std::lock_guard<std::mutex> g(m_);
  int same_class = 0;
  for (auto& p : pools_) {
    if (!(p->key() == key)) continue;
    same_class++;
    const int s = p->acquire();
    if (s >= 0) { *slot = s; return p; }
  }
  // capacity: B2H264_BROKER_SLOTS, else 4, 16, 64, 128, 128, ... for successive pools of a class (a lone encoder stays
  // small: a slot costs ~30 MB of HBM and ~25 MB of pinned host memory at 1080p)
  long cap = env_long("B2H264_BROKER_SLOTS", 0);
  if (cap <= 0) { cap = 4L << (2 * (same_class < 3 ? same_class : 3)); if (cap > 128) cap = 128; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return nullptr;
  const long dev_env = env_long("B2H264_DEVICE", -1);
  const int device = dev_env >= 0 ? (int)(dev_env % ndev) : (next_device_++ % ndev);
  auto p = std::make_shared<Pool>(key, (int)cap, device);
  if (!p->ok()) return nullptr;
  pools_.push_back(p);
  *slot = p->acquire();
  return p;
}

std::shared_ptr<DecPool> Broker::attach_decoder(int width, int height, int* slot) {
  std::lock_guard<std::mutex> g(m_);
  int same_class = 0;
  for (auto& p : dec_pools_) {
    if (p->width() != width || p->height() != height) continue;
    same_class++;
    const int s = p->acquire();
    if (s >= 0) { *slot = s; return p; }
  }
  long cap = env_long("B2H264_BROKER_SLOTS", 0);
  if (cap <= 0) { cap = 4L << (2 * (same_class < 3 ? same_class : 3)); if (cap > 128) cap = 128; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return nullptr;
  const long dev_env = env_long("B2H264_DEVICE", -1);
  const int device = dev_env >= 0 ? (int)(dev_env % ndev) : (next_device_++ % ndev);
  auto p = std::make_shared<DecPool>(width, height, (int)cap, device);
  if (!p->ok()) return nullptr;
  dec_pools_.push_back(p);
  *slot = p->acquire();
  return p;
}

void Broker::detach_decoder(const std::shared_ptr<DecPool>& pool, int slot) {
  pool->release(slot);
  std::lock_guard<std::mutex> g(m_);
  if (pool->registered() == 0)
    for (size_t i = 0; i < dec_pools_.size(); i++)
      if (dec_pools_[i] == pool) { dec_pools_.erase(dec_pools_.begin() + i); break; }
}

void Broker::detach(const std::shared_ptr<Pool>& pool, int slot) {
  pool->release(slot);
  std::lock_guard<std::mutex> g(m_);
  if (pool->registered() == 0)
    for (size_t i = 0; i < pools_.size(); i++)
      if (pools_[i] == pool) { pools_.erase(pools_.begin() + i); break; }     // the last reference frees the encoder
}

}  // namespace b2wels
