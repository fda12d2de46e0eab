// single-warp latency probes (cycles per dependent op)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, int n, double x) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 33 + 1) % 1024;   // pointer chase values
    __syncthreads();
    long long t0, t1;
    double a = x, b = x + 1;
    // dependent DFMA
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
    for (int i = 0; i < n; i++) { a = fma(a, x, b); a = fma(a, x, b); a = fma(a, x, b); a = fma(a, x, b); }
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
    if (threadIdx.x == 0) printf("warps %d  DFMA dependent: %.1f cycles\n", blockDim.x / 32, (double)(t1 - t0) / (4 * n));
    // dependent DADD
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
    for (int i = 0; i < n; i++) { a = a + b; a = a + b; a = a + b; a = a + b; }
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
    if (threadIdx.x == 0) printf("warps %d  DADD dependent: %.1f cycles\n", blockDim.x / 32, (double)(t1 - t0) / (4 * n));
    // 4 independent DFMA chains
    double c0 = x, c1 = x + 2, c2 = x + 3, c3 = x + 4;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
    for (int i = 0; i < n; i++) { c0 = fma(c0, x, b); c1 = fma(c1, x, b); c2 = fma(c2, x, b); c3 = fma(c3, x, b); }
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
    if (threadIdx.x == 0) printf("warps %d  DFMA 4 chains: %.1f cycles per DFMA\n", blockDim.x / 32, (double)(t1 - t0) / (4 * n));
    a += c0 + c1 + c2 + c3;
    // LDS.64 pointer chase
    int idx = threadIdx.x & 31;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
    for (int i = 0; i < n; i++) { idx = (int)sm[idx]; idx = (int)sm[idx]; }
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
    if (threadIdx.x == 0) printf("warps %d  LDS.64 + F2I chase: %.1f cycles\n", blockDim.x / 32, (double)(t1 - t0) / (2 * n));
    a += idx;
    // LDS -> DFMA accumulate, independent loads (address independent), one chain
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
    for (int i = 0; i < n; i++) {
        const double *p = sm + ((i * 8 + threadIdx.x) & 511);
        a = fma(p[0], x, a); a = fma(p[33], x, a); a = fma(p[66], x, a); a = fma(p[99], x, a);
    }
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
    if (threadIdx.x == 0) printf("warps %d  LDS+DFMA one chain: %.1f cycles per pair\n", blockDim.x / 32, (double)(t1 - t0) / (4 * n));
    // shuffle reduction of a double (xor 16..1)
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
    for (int i = 0; i < n; i++) { for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o); }
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
    if (threadIdx.x == 0) printf("warps %d  warp_sum(double): %.1f cycles\n", blockDim.x / 32, (double)(t1 - t0) / n);
    // int dependent IMAD
    int q = threadIdx.x;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
    for (int i = 0; i < n; i++) { q = q * 3 + i; q = q * 5 + i; q = q * 7 + i; q = q * 9 + i; }
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
    if (threadIdx.x == 0) printf("warps %d  IMAD dependent: %.1f cycles\n", blockDim.x / 32, (double)(t1 - t0) / (4 * n));
    // global load chase (L2 / L1)
    if (a == 1234.5 || q == 77) out[0] = a;
}
__global__ void g(const int *__restrict__ chain, int *out, int n) {
    long long t0, t1;
    int idx = threadIdx.x & 31;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
    for (int i = 0; i < n; i++) idx = __ldcg(chain + idx);
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
    if (threadIdx.x == 0) printf("ldcg (L2) chase: %.1f cycles\n", (double)(t1 - t0) / n);
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
    for (int i = 0; i < n; i++) idx = __ldg(chain + idx);
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
    if (threadIdx.x == 0) printf("ldg (L1) chase: %.1f cycles\n", (double)(t1 - t0) / n);
    if (idx == -5) out[0] = idx;
}
int main() {
    double *d; cudaMalloc(&d, 64);
    for (int w : {1, 4, 16}) { k<<<1, 32 * w>>>(d, 256, 1.0000001); cudaDeviceSynchronize(); }
    int h[4096]; for (int i = 0; i < 4096; i++) h[i] = (i * 97 + 31) % 4096;
    int *c; cudaMalloc(&c, sizeof(h)); cudaMemcpy(c, h, sizeof(h), cudaMemcpyHostToDevice);
    g<<<1, 32>>>(c, (int *)d, 512); cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
