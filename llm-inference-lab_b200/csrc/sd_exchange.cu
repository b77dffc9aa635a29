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

// Bounded wait of the start gate.  Unlike the waits inside the step kernels (microseconds; they almost never reach a
// second look at the clock) a gate legitimately waits for as long as the slowest host needs, so its guard has to survive a
// step of %globaltimer: on the first gate launch after the GPU had idled for seconds (the CPU-side oracle check of the
// B = 256 test before it) two reads 40 us apart differed by more than the whole 2 s budget, the plain SpinGuard gave up
// and the gate opened 78 us after the host's release with no peer present (gpurun_out -> profiles/r2j_gate_probe.md).
// This guard restarts its window when the clock runs backwards and needs the budget to be exceeded TWICE, each time
// measured from a fresh reading, before it records the error -- one step of the clock costs nothing, a dead peer is
// still reported after 2 x kSpinBudgetNs.
struct GateClockWindow {
    unsigned long long t0 = 0ull;
    unsigned strikes = 0u;
};
// One look at the clock: true = keep waiting, false = give up.  Host-callable so that the rule is tested on the CPU
// (sd_selftest_gate_clock, tests/test_capi_symbols.py).
__host__ __device__ __forceinline__ bool gate_clock_step(GateClockWindow& w, const unsigned long long now,
                                                         const unsigned long long budget_ns) {
    if (w.t0 == 0ull || now < w.t0) {
        w.t0 = now;
        return true;
    }
    if (now - w.t0 < budget_ns) return true;
    if (++w.strikes < 2u) {
        w.t0 = now;
        return true;
    }
    return false;
}
struct GateSpinGuard {
    GateClockWindow w;
    unsigned polls = 0u;
    SD_DEVINL bool keep_waiting(unsigned site) {
        if ((++polls & 127u) != 0u) return true;
        if (gate_clock_step(w, global_timer_ns(), kSpinBudgetNs)) return true;
        atomicOr(&g_spin_error, site);
        return false;
    }
};

// Start gate (sd_start_gate): one warp.  (1) wait until the HOST has raised *host_go to go_value -- the caller sets it
// after everything behind the gate has been enqueued, so no rank's GPU outruns its own launch queue; (2) tell every
// peer "rank r has arrived at gate g" with one 8-byte store into its header, wait for every peer's word; (3) advance
// the gate counter.  After the gate all ranks proceed within an NVLink store + poll of each other (~2 us), whatever
// the skew of their host threads was.  Exchange header (first 256 bytes of the buffer): [0] u32 exchange epoch,
// [16] u32 gate counter, [64 + 8 r] u64 arrival word of rank r (monotonic: gate + 1).
__global__ void __launch_bounds__(32) start_gate_kernel(const sd_exchange* x, const unsigned* host_go, const unsigned go_value) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x;
    if (host_go && lane == 0) {
        GateSpinGuard guard;
        for (;;) {
            unsigned v;
            asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(host_go) : "memory");
            if (v == go_value) break;
            if (!guard.keep_waiting(kSpinStartGate)) break;
            __nanosleep(200);
        }
    }
    __syncwarp();
    if (!x) return;
    char* own = reinterpret_cast<char*>(x->peer[x->rank]);
    unsigned gate;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(gate) : "l"(own + 16) : "memory");
    const unsigned long long want = (unsigned long long)gate + 1ull;
    for (int r = lane; r < x->world; r += 32) {
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(x->peer[r]) + 64 + 8 * x->rank);
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(want) : "memory");
    }
    for (int r = lane; r < x->world; r += 32) {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(own + 64 + 8 * r);
        GateSpinGuard guard;
        for (;;) {
            unsigned long long v;
            asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
            if (v >= want) break;
            if (!guard.keep_waiting(kSpinStartGate)) break;
            __nanosleep(100);
        }
    }
    __syncwarp();
    if (lane == 0) *reinterpret_cast<volatile unsigned*>(own + 16) = gate + 1u;
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

int sd_start_gate(const sd_exchange* exchange, const void* host_go, unsigned go_value, void* stream) {
    if (!exchange && !host_go) return SD_OK;
    return sd::to_status(sd::launch_kernel(sd::start_gate_kernel, dim3(1), dim3(32), 0, reinterpret_cast<cudaStream_t>(stream),
                                           exchange, reinterpret_cast<const unsigned*>(host_go), go_value));
}

int sd_selftest_gate_clock(const unsigned long long* readings, int n, unsigned long long budget_ns) {
    if (!readings || n < 0) return -2;
    sd::GateClockWindow w;
    for (int i = 0; i < n; ++i)
        if (!sd::gate_clock_step(w, readings[i], budget_ns)) return i;
    return -1;
}

int sd_exchange_wait(const sd_exchange* exchange, int32_t* out, void* stream) {
    if (!exchange) return SD_ERR_INVALID_ARG;
    return sd::to_status(sd::launch_kernel(sd::exchange_wait_kernel, dim3(1), dim3(256), 0,
                                           reinterpret_cast<cudaStream_t>(stream), exchange, out));
}

}  // extern "C"
