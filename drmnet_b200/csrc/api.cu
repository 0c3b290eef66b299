// ABI glue for libdrmrender.so: version, thread-local error string, driver entry point lookup.
#include <string.h>

#include "common.cuh"

namespace drm {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
long long launches() { return (long long)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

}  // namespace drm

namespace drm { long long launches(); }
extern "C" int64_t drm_launch_count(void) { return drm::launches(); }
extern "C" int drm_version(void) { return DRM_VERSION; }
extern "C" const char* drm_last_error(void) { return drm::g_error; }
