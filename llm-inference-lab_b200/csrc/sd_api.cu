// Library-level entry points: ABI version, error strings, device query.
#include "sd_common.cuh"

#include <stddef.h>
#include <stdlib.h>

namespace sd {
static int g_pdl = -1;
int pdl_enabled() {
    if (g_pdl < 0) {
        const char* e = getenv("SD_B200_PDL");
        g_pdl = (e && e[0] == '0') ? 0 : 1;
    }
    return g_pdl;
}
}  // namespace sd

extern "C" {

int sd_set_pdl(int enabled) {
    sd::g_pdl = enabled ? 1 : 0;
    return SD_OK;
}

int sd_abi_version(void) { return SD_ABI_VERSION; }

// Layout self-description for foreign-function bindings: sizeof / selected offsetof of the public structs.
size_t sd_struct_layout(int which) {
    switch (which) {
        case 0: return sizeof(sd_kv_layer);
        case 1: return sizeof(sd_exchange);
        case 2: return offsetof(sd_exchange, peer);
        case 3: return sizeof(sd_step_args);
        case 4: return offsetof(sd_step_args, layers);
        case 5: return offsetof(sd_step_args, row_bytes);
        case 6: return offsetof(sd_step_args, seq_lens);
        case 7: return offsetof(sd_step_args, step_state);
        default: return 0;
    }
}

const char* sd_error_string(int code) {
    switch (code) {
        case SD_OK: return "ok";
        case SD_ERR_INVALID_ARG: return "invalid argument";
        case SD_ERR_SHAPE: return "shape mismatch";
        case SD_ERR_WORKSPACE: return "workspace missing or too small";
        case SD_ERR_ALIGNMENT: return "pointer or stride misaligned";
        case SD_ERR_UNSUPPORTED: return "unsupported configuration";
        case SD_ERR_CAPACITY: return "KV arena capacity exceeded";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown sd_status";
}

int sd_device_error(int clear, unsigned* sites_out) {
    unsigned a = 0u, b = 0u, c = 0u;
    int rc = sd::verify_tu_spin_error(clear, &a);
    if (rc == 0) rc = sd::kv_tu_spin_error(clear, &b);
    if (rc == 0) rc = sd::exchange_tu_spin_error(clear, &c);
    if (sites_out) *sites_out = a | b | c;
    return rc;
}

int sd_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int v = 0;
    if (sm_count) {
        e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        *sm_count = v;
    }
    if (cc_major) {
        e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev);
        if (e != cudaSuccess) return (int)e;
        *cc_major = v;
    }
    if (cc_minor) {
        e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev);
        if (e != cudaSuccess) return (int)e;
        *cc_minor = v;
    }
    return 0;
}

}  // extern "C"
