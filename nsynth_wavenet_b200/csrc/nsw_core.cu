// Library-wide host utilities: error string, launch counter, named-tensor lookup.
#include "nsw_common.cuh"

#include <mutex>

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

static std::vector<range_setter_fn>& range_setters() {
  static std::vector<range_setter_fn> v;
  return v;
}
void register_range_setter(range_setter_fn fn) { range_setters().push_back(fn); }

static std::mutex g_range_mu;
static volatile unsigned int* g_range_flag = nullptr;  // mapped pinned host word, shared by all devices
static std::vector<int> g_range_devices;

int range_guard_init(int device) {
  std::lock_guard<std::mutex> lock(g_range_mu);
  if (!g_range_flag) {
    void* p = nullptr;
    NSW_CUDA(cudaHostAlloc(&p, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(p, 0, 64);
    g_range_flag = static_cast<volatile unsigned int*>(p);
  }
  for (int d : g_range_devices)
    if (d == device) return NSW_OK;
  unsigned int* dptr = nullptr;
  NSW_CUDA(cudaHostGetDevicePointer((void**)&dptr, (void*)g_range_flag, 0));
  for (range_setter_fn fn : range_setters()) NSW_TRY(fn(dptr));
  g_range_devices.push_back(device);
  return NSW_OK;
}

int range_check(const char* who) {
  if (!g_range_flag || *g_range_flag == 0) return NSW_OK;
  *g_range_flag = 0;
  set_error("%s: an activation outside the fp16 range (|v| > 65504, or NaN) reached a split-fp16 tensor-core "
            "operand; the result is invalid.  Use engine 'ffma' (fp32 CUDA cores) for this model.", who);
  return NSW_ERANGE;
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
int nsw_range_status(int32_t device) {
  if (cudaSetDevice(device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    nsw::set_error("nsw_range_status: device %d is not usable", device);
    return NSW_ECUDA;
  }
  return nsw::range_check("nsw_range_status");
}
}
