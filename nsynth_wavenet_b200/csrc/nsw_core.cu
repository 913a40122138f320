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

int ensure_dynamic_smem(const void* kernel, int bytes, std::atomic<uint64_t>& done) {
  int dev = 0;
  NSW_CUDA(cudaGetDevice(&dev));
  const uint64_t bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return NSW_OK;
  NSW_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));  // idempotent
  done.fetch_or(bit, std::memory_order_release);
  return NSW_OK;
}

int upload(DevBuf& buf, const void* host, size_t bytes) {
  NSW_TRY(buf.ensure(bytes));
  NSW_CUDA(cudaMemcpy(buf.p, host, bytes, cudaMemcpyHostToDevice));
  return NSW_OK;
}

}  // namespace nsw

// CRC-32C (Castagnoli, reflected polynomial 0x82f63b78), slice-by-8: the checksum TensorFlow's V2 checkpoint bundles
// carry for every index block and every tensor (tensorflow/core/lib/hash/crc32c.h); host only.
static uint32_t g_crc_tab[8][256];
static void crc_init() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82f63b78u : c >> 1;
    g_crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xff];
}

extern "C" {
uint32_t nsw_crc32c(const void* data, size_t n, uint32_t crc) {
  static std::once_flag once;
  std::call_once(once, crc_init);
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint32_t c = crc ^ 0xffffffffu;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = g_crc_tab[7][lo & 0xff] ^ g_crc_tab[6][(lo >> 8) & 0xff] ^ g_crc_tab[5][(lo >> 16) & 0xff] ^ g_crc_tab[4][lo >> 24] ^
        g_crc_tab[3][hi & 0xff] ^ g_crc_tab[2][(hi >> 8) & 0xff] ^ g_crc_tab[1][(hi >> 16) & 0xff] ^ g_crc_tab[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = g_crc_tab[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return c ^ 0xffffffffu;
}
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
