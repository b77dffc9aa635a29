// Peer-memory result exchange for the batch-sharded verify step (include/sd_verify.h, "Multi-GPU
// result exchange").  The producing side lives in the verify kernel's sequence finalizer
// (sd_verify.cu); this file holds the buffer management (cudaMalloc + CUDA IPC) and the wait kernel.
#include "sd_common.cuh"

namespace sd {

constexpr int kExchangeHeaderBytes = 256;  // {u32 epoch, u32 arrived[2]}; must match sd_verify.cu

SD_DEVINL unsigned ld_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
SD_DEVINL int ld_sys_s32(const int32_t* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// One CTA.  Spins (with back-off) until all total_rows records of the current epoch have been counted
// in this rank's buffer, copies them out, then advances the epoch for the next verify launch.
__global__ void __launch_bounds__(256) exchange_wait_kernel(const sd_exchange* x, int32_t* out) {
    pdl_wait();
    pdl_trigger();
    char* own = reinterpret_cast<char*>(x->peer[x->rank]);
    unsigned* hdr = reinterpret_cast<unsigned*>(own);
    const unsigned epoch = ld_sys_u32(hdr);
    const unsigned parity = epoch & 1u;
    const unsigned target = (epoch / 2u + 1u) * (unsigned)x->total_rows;
    if (threadIdx.x == 0) {
        while ((int)(ld_sys_u32(hdr + 1 + parity) - target) < 0) __nanosleep(200);
        __threadfence_system();
    }
    __syncthreads();
    const int n = x->world * x->rows_per_rank * x->width;
    if (out) {
        const int32_t* recs = reinterpret_cast<const int32_t*>(own + kExchangeHeaderBytes) + (size_t)parity * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = ld_sys_s32(recs + i);
    }
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned*>(hdr) = epoch + 1u;
}

}  // namespace sd

extern "C" {

size_t sd_exchange_bytes(int world, int rows_per_rank, int width) {
    if (world <= 0 || rows_per_rank <= 0 || width <= 0) return 0;
    return (size_t)sd::kExchangeHeaderBytes + (size_t)2 * world * rows_per_rank * width * sizeof(int32_t);
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
