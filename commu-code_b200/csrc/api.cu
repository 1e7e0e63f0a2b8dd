// C-ABI housekeeping: error buffer, device info, launch counter, per-kernel-class device timers.
#include <mutex>
#include <vector>
#include "api_common.h"

namespace cb_host {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

namespace {
struct ProfState {
  unsigned mask = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> used[PROF_NUM];
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
  long long launches = 0;
  std::mutex mu;
};
ProfState& prof() {
  static ProfState s;
  return s;
}
}  // namespace

void count_launch(int n) { prof().launches += n; }

ProfScope::ProfScope(int cls, cudaStream_t s) : cls_(cls), s_(s), slot_(-1) {
  ProfState& st = prof();
  if (!(st.mask & (1u << cls))) return;
  std::lock_guard<std::mutex> g(st.mu);
  std::pair<cudaEvent_t, cudaEvent_t> ev;
  if (!st.pool.empty()) {
    ev = st.pool.back();
    st.pool.pop_back();
  } else {
    cudaEventCreate(&ev.first);
    cudaEventCreate(&ev.second);
  }
  cudaEventRecord(ev.first, s);
  st.used[cls].push_back(ev);
  slot_ = static_cast<int>(st.used[cls].size()) - 1;
}
ProfScope::~ProfScope() {
  if (slot_ < 0) return;
  ProfState& st = prof();
  std::lock_guard<std::mutex> g(st.mu);
  cudaEventRecord(st.used[cls_][slot_].second, s_);
}

}  // namespace cb_host

extern "C" {

const char* commu_last_error(void) { return cb_host::error_buffer(); }

int commu_abi_version(void) { return 3; }   // 3: commu_clip_adam weight_decay, commu_relattn_bwd workspace (materialised dS), stored probabilities

int commu_device_info(int* sm_major, int* sm_minor, int* num_sms) {
  int dev = 0;
  CB_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CB_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_major) *sm_major = prop.major;
  if (sm_minor) *sm_minor = prop.minor;
  if (num_sms) *num_sms = prop.multiProcessorCount;
  return 0;
}

int commu_prof_arm(unsigned class_mask) {
  cb_host::prof().mask = class_mask;
  return 0;
}

// Sums (and clears) the recorded device time of one kernel class. Synchronises the device.
int commu_prof_read(int cls, float* total_ms, int* launches) {
  using namespace cb_host;
  CB_REQUIRE(cls >= 0 && cls < PROF_NUM, "prof: bad class %d", cls);
  CB_CHECK_CUDA(cudaDeviceSynchronize());
  auto& st = prof();
  std::lock_guard<std::mutex> g(st.mu);
  float tot = 0.f;
  for (auto& ev : st.used[cls]) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev.first, ev.second);
    tot += ms;
    st.pool.push_back(ev);
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = static_cast<int>(st.used[cls].size());
  st.used[cls].clear();
  return 0;
}

long long commu_launch_count(int reset) {
  long long v = cb_host::prof().launches;
  if (reset) cb_host::prof().launches = 0;
  return v;
}

}  // extern "C"
