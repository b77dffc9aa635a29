// Peer-memory result exchange for the batch-sharded verify step (include/sd_verify.h, "Multi-GPU
// result exchange").  The producing side lives in the verify kernel's sequence finalizer
// (sd_verify.cu); this file holds the buffer management (cudaMalloc + CUDA IPC) and the wait kernel.
#include "sd_common.cuh"

namespace sd {

// One CTA: gather the records of the current epoch (polls the tagged words), then advance the epoch.
__global__ void __launch_bounds__(256) exchange_wait_kernel(const sd_exchange* x, int32_t* out) {
    pdl_wait();
    pdl_trigger();
    exchange_gather_cta(x, out);
}

int exchange_tu_spin_error(int clear, unsigned* out) { return tu_spin_error(clear, out); }

}  // namespace sd

extern "C" {

size_t sd_exchange_bytes(int world, int rows_per_rank, int width) {
    if (world <= 0 || rows_per_rank <= 0 || width <= 0) return 0;
    return (size_t)sd::kExchangeHeaderBytes + (size_t)sd::kExchangeSlots * world * rows_per_rank * width * sizeof(unsigned long long);
}

int sd_exchange_alloc(size_t bytes, void** ptr_out, void* handle_out) {
    if (!ptr_out || bytes == 0) return SD_ERR_INVALID_ARG;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return sd::to_status(e);
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess && handle_out) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, p);
        if (e == cudaSuccess) memcpy(handle_out, &h, sizeof(h));
    }
    if (e != cudaSuccess) {
        cudaFree(p);
        return sd::to_status(e);
    }
    *ptr_out = p;
    return SD_OK;
}

int sd_exchange_open(const void* handle, void** peer_ptr_out) {
    if (!handle || !peer_ptr_out) return SD_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    return sd::to_status(cudaIpcOpenMemHandle(peer_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
}

int sd_exchange_close(void* peer_ptr) {
    if (!peer_ptr) return SD_OK;
    return sd::to_status(cudaIpcCloseMemHandle(peer_ptr));
}

int sd_exchange_free(void* ptr) {
    if (!ptr) return SD_OK;
    return sd::to_status(cudaFree(ptr));
}

int sd_exchange_wait(const sd_exchange* exchange, int32_t* out, void* stream) {
    if (!exchange) return SD_ERR_INVALID_ARG;
    return sd::to_status(sd::launch_kernel(sd::exchange_wait_kernel, dim3(1), dim3(256), 0,
                                           reinterpret_cast<cudaStream_t>(stream), exchange, out));
}

}  // extern "C"
