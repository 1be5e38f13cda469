// Thread-local error string behind cidnet_last_error().
#include "common.cuh"

#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

namespace cidnet {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) { g_last_error = msg; return code; }

namespace {
std::mutex g_attr_mutex;
std::map<std::pair<int, const void*>, int> g_attr_bytes;     // (device, kernel) -> largest size configured
int g_sms[64];
}  // namespace

int ensure_dynamic_smem(const void* kernel, int bytes) {
    int dev = 0;
    CIDNET_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    int& have = g_attr_bytes[std::make_pair(dev, kernel)];
    if (have >= bytes) return CIDNET_OK;
    CIDNET_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    have = bytes;
    return CIDNET_OK;
}

bool pdl_enabled() {
    static const bool on = [] { const char* v = getenv("CIDNET_PDL"); return !(v && v[0] == '0'); }();
    return on;
}

int device_sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    if (g_sms[dev] <= 0) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        g_sms[dev] = sms > 0 ? sms : 148;
    }
    return g_sms[dev];
}
}  // namespace cidnet

extern "C" const char* cidnet_last_error(void) { return cidnet::g_last_error.c_str(); }
extern "C" int cidnet_abi_version(void) { return 1; }
extern "C" int cidnet_act_dtype(void) {
#ifdef CIDNET_ACT_BF16
    return 1;
#else
    return 0;
#endif
}
