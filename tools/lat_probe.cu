// Micro-probe: dependent-chain latency of FP64 ops, smem loads and FP64 division on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    double a = out[0] + 1.0000001, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = fma(a, b, c);
    long long t1 = clock64();
    double d = a;
    for (int i = 0; i < n; ++i) d = 1.0 / (d + 1.5);
    long long t2 = clock64();
    int idx = ((int)d) & 1023;
    double e = 0;
    for (int i = 0; i < n; ++i) { e += sm[idx]; idx = ((int)(e * 1e-30) + idx + 1) & 1023; }
    long long t3 = clock64();
    double f = d;
    for (int i = 0; i < n; ++i) f = f * b;
    long long t4 = clock64();
    double g = d;
    for (int i = 0; i < n; ++i) g = g + c;
    long long t5 = clock64();
    double s = e, cs;
    for (int i = 0; i < n / 8; ++i) { sincos(s, &s, &cs); s += cs; }
    long long t6 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; }
    out[threadIdx.x + blockIdx.x * blockDim.x + 1] = a + d + e + f + g + s;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMemset(out, 0, 1 << 24); cudaMalloc(&cyc, 64);
    const int n = 4096;
    for (int warps_per_sm : {1, 4, 8, 16, 32}) {
        int threads = 32, blocks = 148 * warps_per_sm;
        k<<<blocks, threads>>>(out, cyc, n); cudaDeviceSynchronize();
        k<<<blocks, threads>>>(out, cyc, n); cudaDeviceSynchronize();
        long long h[6]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("warps/SM %2d: dfma %.1f  ddiv %.1f  lds-chain %.1f  dmul %.1f  dadd %.1f  sincos %.1f cycles per op\n", warps_per_sm,
               h[0] / (double)n, h[1] / (double)n, h[2] / (double)n, h[3] / (double)n, h[4] / (double)n, h[5] / (double)(n / 8));
    }
    return 0;
}
