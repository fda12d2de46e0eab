// Roofline denominators for the fold kernels: the DP is bound by the SM integer pipe
// (one VIADDMNMX per relaxation) and the shared-memory load pipe, not by HBM, and MEASURED_PEAKS.json
// only carries HBM / bf16 numbers.  These two kernels measure the machine's add-min issue rate and its
// conflict-free 32-bit shared-memory load rate with CUDA events so that bench.py can state
// roofline.peak from a live measurement on the same GPU and clocks as the timed run.
#include <cuda_runtime.h>

#include "../../include/scanfold_b200.h"

namespace {

constexpr int MB_THREADS = 1024;
constexpr int MB_CHAINS = 8;

// 8 rotating accumulators: acc[k] = min(acc[k] + x, acc[k+1]) -- one VIADDMNMX each, no other ALU work
__global__ void __launch_bounds__(MB_THREADS) addmin_kernel(int *out, int iters, int x) {
    int acc[MB_CHAINS];
#pragma unroll
    for (int k = 0; k < MB_CHAINS; k++) acc[k] = threadIdx.x * (k + 1) + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int k = 0; k < MB_CHAINS; k++) acc[k] = __viaddmin_s32(acc[k], x, acc[(k + 1) % MB_CHAINS]);
        }
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < MB_CHAINS; k++) s ^= acc[k];
    if (s == 0x7fffffff) out[0] = s;
}

// conflict-free LDS.32 stream: lane l of every warp reads bank l
__global__ void __launch_bounds__(MB_THREADS) smem_kernel(int *out, int iters) {
    __shared__ int buf[8192];
    for (int k = threadIdx.x; k < 8192; k += MB_THREADS) buf[k] = k ^ blockIdx.x;
    __syncthreads();
    int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++) {
            a0 += buf[(idx + 0 * 1024) & 8191];
            a1 += buf[(idx + 1 * 1024) & 8191];
            a2 += buf[(idx + 2 * 1024) & 8191];
            a3 += buf[(idx + 3 * 1024) & 8191];
            idx += 4096 + 32;
        }
    }
    int s = a0 ^ a1 ^ a2 ^ a3;
    if (s == 0x7fffffff) out[0] = s;
}

// 8 independent DFMA chains per thread: the fp64 FMA issue rate the partition-function kernels are measured against
__global__ void __launch_bounds__(MB_THREADS) dfma_kernel(int *out, int iters, double x) {
    double acc[MB_CHAINS];
#pragma unroll
    for (int k = 0; k < MB_CHAINS; k++) acc[k] = 1e-3 * (threadIdx.x + k) + blockIdx.x;
    const double y = 1.0 - 1e-9 * x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int k = 0; k < MB_CHAINS; k++) acc[k] = fma(acc[k], y, x);
        }
    }
    double s = 0.;
#pragma unroll
    for (int k = 0; k < MB_CHAINS; k++) s += acc[k];
    if (s == 12345.678) out[0] = 1;
}

}  // namespace

extern "C" int sfb_microbench(int which, double *ops_per_s) {
    if (!ops_per_s || (which != SFB_MICROBENCH_ADDMIN && which != SFB_MICROBENCH_SMEM_LD32 && which != SFB_MICROBENCH_DFMA)) return SFB_E_ARG;
    int dev = 0, n_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return SFB_E_CUDA;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    int *d_out = nullptr;
    if (cudaMalloc(&d_out, 64) != cudaSuccess) return SFB_E_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = n_sm * 2;
    const int iters = which == SFB_MICROBENCH_ADDMIN ? 8192 : (which == SFB_MICROBENCH_DFMA ? 1024 : 2048);
    double best = 0.;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        if (which == SFB_MICROBENCH_ADDMIN)
            addmin_kernel<<<grid, MB_THREADS>>>(d_out, iters, rep + 1);
        else if (which == SFB_MICROBENCH_DFMA)
            dfma_kernel<<<grid, MB_THREADS>>>(d_out, iters, 1e-3 * (rep + 1));
        else
            smem_kernel<<<grid, MB_THREADS>>>(d_out, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double per_thread = which == SFB_MICROBENCH_SMEM_LD32 ? (double)iters * 8 * 4 : (double)iters * 4 * MB_CHAINS;
        const double ops = per_thread * MB_THREADS * grid;
        if (rep > 0 && ms > 0.f) best = ops / (ms * 1e-3) > best ? ops / (ms * 1e-3) : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    if (cudaGetLastError() != cudaSuccess || best <= 0.) return SFB_E_CUDA;
    *ops_per_s = best;
    return 0;
}
