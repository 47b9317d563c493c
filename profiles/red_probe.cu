// red_probe.cu - fp64 atomicAdd (RED.E.ADD.F64) throughput of one GPU for the access shapes of the CSR scatter.
// Stand-alone probe (not part of the library):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/red_probe profiles/red_probe.cu
//   pattern 0: streaming, every lane its own consecutive double (fully coalesced, no reuse)
//   pattern 1: 24-byte runs (3 doubles) at pseudo-random 24-byte-aligned places (one CSR (row, neighbour) block)
//   pattern 2: like the element kernel: 32 lanes = 32 consecutive doubles of "rows" of 81 doubles, 8 passes over the
//              same 2^k rows (every slot receives 8 contributions, as a node's 8 cells give)
// Output: G lane-atomics / s and the equivalent GB/s of 8-byte operands, for a footprint inside and outside the L2.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__global__ void k_red(double* dst, uint64_t n, int pattern, int reps) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t T = (uint64_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        uint64_t i;
        if (pattern == 0) {
            i = (t + (uint64_t)r * T) % n;
        } else if (pattern == 1) {
            const uint64_t run = (t + (uint64_t)r * T) / 3, k = (t + (uint64_t)r * T) % 3;
            const uint64_t h = (run * 0x9E3779B97F4A7C15ull) >> 20;
            i = ((h % (n / 3)) * 3 + k) % n;
        } else {
            const uint64_t w = (t + (uint64_t)r * T) >> 5, lane = t & 31;
            const uint64_t row = (w * 2654435761ull) % (n / 96);
            i = row * 96 + ((w >> 3) % 3) * 32 + lane;
        }
        atomicAdd(dst + i, 1.0);
    }
}

int main() {
    const uint64_t sizes[2] = {4ull << 20, 256ull << 20};      // doubles: 32 MB (inside L2), 2 GB (outside)
    double* d = nullptr;
    cudaMalloc(&d, sizes[1] * 8);
    cudaMemset(d, 0, sizes[1] * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int s = 0; s < 2; ++s)
        for (int pat = 0; pat < 3; ++pat) {
            const int reps = 64, blocks = 148 * 16, threads = 256;
            k_red<<<blocks, threads>>>(d, sizes[s], pat, 4);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k_red<<<blocks, threads>>>(d, sizes[s], pat, reps);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double n = (double)blocks * threads * reps;
            printf("footprint %5llu MB pattern %d: %.1f G atomics/s (%.0f GB/s of operands), %.3f ms\n",
                   (unsigned long long)(sizes[s] * 8 >> 20), pat, n / ms / 1e6, n * 8 / ms / 1e6, ms);
        }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
