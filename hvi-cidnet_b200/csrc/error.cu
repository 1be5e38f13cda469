// Thread-local error string behind cidnet_last_error().
#include "common.cuh"

namespace cidnet {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
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
