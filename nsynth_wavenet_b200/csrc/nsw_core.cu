// Library-wide host utilities: error string, launch counter, named-tensor lookup.
#include "nsw_common.cuh"

namespace nsw {

static thread_local char g_err[1024] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::string canonical_name(const char* raw) {
  std::string s(raw ? raw : "");
  const std::string ema = "/ExponentialMovingAverage";  // fastgen.py:12-14
  if (s.size() >= 2 && s.compare(s.size() - 2, 2, ":0") == 0) s.resize(s.size() - 2);
  if (s.size() >= ema.size() && s.compare(s.size() - ema.size(), ema.size(), ema) == 0)
    s.resize(s.size() - ema.size());
  return s;
}

TensorMap::TensorMap(const nsw_tensor* t, int n) {
  for (int i = 0; i < n; ++i) map_[canonical_name(t[i].name)] = &t[i];
}

const float* TensorMap::get(const std::string& name, int64_t expect_numel) const {
  auto it = map_.find(name);
  if (it == map_.end()) {
    set_error("missing weight tensor '%s'", name.c_str());
    return nullptr;
  }
  int64_t n = 1;
  for (int d = 0; d < it->second->ndim; ++d) n *= it->second->shape[d];
  if (n != expect_numel || it->second->data == nullptr) {
    set_error("weight tensor '%s' has %lld elements, expected %lld", name.c_str(), (long long)n,
              (long long)expect_numel);
    return nullptr;
  }
  return it->second->data;
}

int upload(DevBuf& buf, const void* host, size_t bytes) {
  NSW_TRY(buf.ensure(bytes));
  NSW_CUDA(cudaMemcpy(buf.p, host, bytes, cudaMemcpyHostToDevice));
  return NSW_OK;
}

}  // namespace nsw

extern "C" {
int nsw_version(void) { return 100; }
const char* nsw_last_error(void) { return nsw::g_err; }
uint64_t nsw_kernel_launch_count(void) { return nsw::g_launch_count.load(); }
}
