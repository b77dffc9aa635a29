// Kernel 3 placeholder (replaced by the real kernel in this round; keeps the ABI complete).
#include "sd_common.cuh"
extern "C" {
size_t sd_verify_attention_workspace_bytes(int, int, int, int, int) { return 256; }
int sd_verify_attention(const void*, const void*, const void*, const int32_t*, void*, void*, size_t, int, int, int,
                        int, int, int, int64_t, int64_t, int64_t, float, int, void*) {
    return SD_ERR_UNSUPPORTED;
}
}
