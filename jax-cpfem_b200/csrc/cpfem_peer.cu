// cpfem_peer.cu - peer-memory mailbox for the interface exchange of element-partitioned runs.
//
// One process per GPU on one node.  After the local assembly a rank ships the residual entries and CSR rows of the
// interface nodes it does not own to their owner (cpfem_b200/partition.py, DESIGN.md "Multi-GPU").  The reference is
// single-device and has no such step; round 1 did it with NCCL send/recv (0.40 ms per assembly at 4 GPUs, 53 MB of CSR
// rows each way at ~200 GB/s plus a rendezvous).  Here the SENDER's copy kernel stores the rows straight into a mailbox
// in the owner's memory over NVLink (CUDA IPC mapping, plain 16-byte stores, no staging, no rendezvous) and releases a
// flag the owner's stream waits on:
//     cpfem_peer_alloc / cpfem_peer_open   mailbox = cudaMalloc'ed by this library (not the caching allocator of the
//                                          host framework, so the IPC handle is that of exactly this buffer)
//     cpfem_peer_put                       dst_peer[i] = src[map ? map[i] : i], then flag_peer = epoch (system scope)
//     cpfem_peer_wait                      stream-ordered wait until flag_local >= epoch (bounded spin, status on timeout)
//     cpfem_peer_signal                    flag_peer = epoch (acknowledgement: "mailbox consumed")
// All calls are enqueue-only.  The flags are 64-bit epochs that only grow, so a late reader never sees an old state as new.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "cpfem_internal.h"

static_assert(sizeof(cudaIpcMemHandle_t) == CPFEM_PEER_HANDLE_BYTES, "cpfem.h: CPFEM_PEER_HANDLE_BYTES");

extern "C" int cpfem_peer_alloc(int64_t bytes, void** ptr, uint8_t* handle) {
    if (bytes <= 0 || !ptr || !handle) return set_err(-1, "cpfem_peer_alloc: bad argument");
    void* p = nullptr;
    CU_TRY(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);          // flags start at epoch 0
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return set_err(-2, "cpfem_peer_alloc", e);
    }
    memcpy(handle, &h, sizeof h);
    *ptr = p;
    return 0;
}

extern "C" int cpfem_peer_open(const uint8_t* handle, void** ptr) {
    if (!handle || !ptr) return set_err(-1, "cpfem_peer_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    CU_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr = p;
    return 0;
}

extern "C" int cpfem_peer_close(void* ptr) {
    if (ptr) CU_TRY(cudaIpcCloseMemHandle(ptr));
    return 0;
}

extern "C" int cpfem_peer_free(void* ptr) {
    if (ptr) CU_TRY(cudaFree(ptr));
    return 0;
}

// -----------------------------------------------------------------------------------------------
// kernels
// -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Copy (optionally gathered) into peer memory; the last block to finish releases the flag.  `done` is a local counter
// (one per in-flight put, zeroed by the kernel that consumes it).  Grid-stride, two doubles per store where aligned.
__global__ void __launch_bounds__(256)
k_peer_put(double* __restrict__ dst, const double* __restrict__ src, const int64_t* __restrict__ map, int64_t n,
           unsigned long long* flag_peer, unsigned long long epoch, unsigned int* done) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    if (map) {
        for (int64_t i = tid; i < n; i += nth) dst[i] = src[map[i]];
    } else if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
        const int64_t n2 = n >> 1;
        const double2* s2 = reinterpret_cast<const double2*>(src);
        double2* d2 = reinterpret_cast<double2*>(dst);
        for (int64_t i = tid; i < n2; i += nth) d2[i] = s2[i];
        if (tid == 0 && (n & 1)) dst[n - 1] = src[n - 1];
    } else {
        for (int64_t i = tid; i < n; i += nth) dst[i] = src[i];
    }
    if (!flag_peer) return;
    __threadfence_system();                    // this thread's stores are visible system-wide before the count below
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int k = atomicAdd(done, 1u);
        if (k == gridDim.x - 1) {
            *done = 0u;                        // ready for the next put on this stream
            __threadfence_system();
            st_release_sys(flag_peer, epoch);
        }
    }
}

__global__ void k_peer_signal(unsigned long long* flag_peer, unsigned long long epoch) {
    __threadfence_system();
    st_release_sys(flag_peer, epoch);
}

// One thread spins until the flag reaches `epoch`.  Bounded: after `timeout_cycles` it records the miss in status[1]
// (the "error" word of the library's 4-word status) and returns, so that a lost peer cannot hang the device.
__global__ void k_peer_wait(const unsigned long long* flag_local, unsigned long long epoch, long long timeout_cycles,
                            long long* status) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag_local) < epoch) {
        if (clock64() - t0 > timeout_cycles) {
            if (status) atomicAdd((unsigned long long*)&status[1], 1ULL);
            return;
        }
        __nanosleep(200);
    }
}

// -----------------------------------------------------------------------------------------------
// entry points
// -----------------------------------------------------------------------------------------------
static unsigned int* peer_counter(int device) {
    // one completion counter per device, zero between puts (puts of one process are stream-ordered by the caller)
    static unsigned int* ctr[64] = {nullptr};
    if (device < 0 || device >= 64) return nullptr;
    if (!ctr[device]) {
        if (cudaMalloc((void**)&ctr[device], 256) != cudaSuccess) return nullptr;
        cudaMemset(ctr[device], 0, 256);
        cudaDeviceSynchronize();
    }
    return ctr[device];
}

extern "C" int cpfem_peer_put(double* dst_peer, const double* src, const int64_t* map, int64_t n, uint64_t* flag_peer,
                              uint64_t epoch, void* stream_) {
    if (n < 0 || (n > 0 && (!dst_peer || !src))) return set_err(-1, "cpfem_peer_put: bad argument");
    if (n == 0 && !flag_peer) return 0;
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    unsigned int* done = peer_counter(dev);
    if (!done) return set_err(-2, "cpfem_peer_put: counter allocation failed");
    int64_t blocks = (n / 2 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 4) blocks = 148 * 4;      // enough stores in flight to fill the links, few enough to leave SMs alone
    k_peer_put<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(dst_peer, src, map, n, (unsigned long long*)flag_peer,
                                                                   (unsigned long long)epoch, done);
    cpfem_count_launches(1);
    CU_TRY(cudaGetLastError());
    return 0;
}

extern "C" int cpfem_peer_signal(uint64_t* flag_peer, uint64_t epoch, void* stream_) {
    if (!flag_peer) return set_err(-1, "cpfem_peer_signal: null argument");
    k_peer_signal<<<1, 1, 0, (cudaStream_t)stream_>>>((unsigned long long*)flag_peer, (unsigned long long)epoch);
    cpfem_count_launches(1);
    CU_TRY(cudaGetLastError());
    return 0;
}

extern "C" int cpfem_peer_wait(const uint64_t* flag_local, uint64_t epoch, double timeout_s, int64_t* status, void* stream_) {
    if (!flag_local) return set_err(-1, "cpfem_peer_wait: null argument");
    if (!(timeout_s > 0.0)) timeout_s = 20.0;
    const long long cycles = (long long)(timeout_s * 1.9e9);
    k_peer_wait<<<1, 1, 0, (cudaStream_t)stream_>>>((const unsigned long long*)flag_local, (unsigned long long)epoch, cycles,
                                                    (long long*)status);
    cpfem_count_launches(1);
    CU_TRY(cudaGetLastError());
    return 0;
}
