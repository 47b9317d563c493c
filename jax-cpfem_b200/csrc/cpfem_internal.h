// cpfem_internal.h - shared by the translation units of libcpfem_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/cpfem.h"
#include "cp_point.cuh"

// -----------------------------------------------------------------------------------------------
// error plumbing
// -----------------------------------------------------------------------------------------------
int cpfem_set_err(int code, const char* what, cudaError_t e = cudaSuccess);
void cpfem_count_launches(int n);     // feeds cpfem_launch_count()
static inline int set_err(int code, const char* what, cudaError_t e = cudaSuccess) { return cpfem_set_err(code, what, e); }
#define CU_TRY(x)                                                  \
    do {                                                           \
        cudaError_t _e = (x);                                      \
        if (_e != cudaSuccess) return set_err(-2, #x, _e);         \
    } while (0)

// -----------------------------------------------------------------------------------------------
// plan
// -----------------------------------------------------------------------------------------------
struct cpfem_plan {
    int64_t nc = 0, nn = 0, nnz = 0;
    int64_t nc_active = 0;        // kernels loop over the first nc_active cells (owned cells of a partition)
    int32_t ns = 0;
    int32_t max_valence = 0;
    int32_t* cells = nullptr;     // (nc,8)
    double* points = nullptr;     // (nn,3)
    int64_t* indptr = nullptr;    // (3 nn + 1)
    int32_t* indices = nullptr;   // (nnz)
    uint8_t* rank = nullptr;      // (nc,8,8): rank of node b in the sorted neighbour list of node a
    // node-level adjacency behind the pattern (kept for the node-block SpMV of the device solver): the rows 3n..3n+2
    // own the 9 m(n) entries starting at 9 nbr_ptr[n], laid out [i][j][k] over the sorted neighbours nbr[nbr_ptr[n] + j]
    int64_t* nbr_ptr = nullptr;   // (nn + 1)
    int32_t* nbr = nullptr;       // (nnz / 9)
    // assembly pipeline: the point kernel of chunk i+1 (FP64-bound, caller's stream) overlaps the element kernel of
    // chunk i (load/store- and atomics-bound, plan-owned high-priority stream); two scratch buffers alternate
    double* scratch[2] = {nullptr, nullptr};   // each (quads, 90, 32): P JxW and dP/dH JxW, quad-major (SCR_QUAD)
    int64_t chunk_cells = 0;                   // cells per assembly chunk
    std::vector<int64_t> zero_end;             // per chunk: first CSR slot behind the rows that chunks 0..k can touch
    cudaStream_t elem_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_point[2] = {nullptr, nullptr}, ev_elem[2] = {nullptr, nullptr};
    // multi-GPU overlap: cpfem_newton_update records `progress_event` (caller-owned) on its stream once the contributions
    // of cells [0, progress_cells) are complete, so that the interface exchange can start while later chunks compute
    cudaEvent_t progress_event = nullptr;
    int64_t progress_cells = 0;
    CpSlip slip;
    int device = 0;
    int sm_count = 148;
    void* solver_ws = nullptr;    // device linear solver workspace (cpfem_solver.cu), allocated on first use
};
void cpfem_solver_ws_free(void* ws);
