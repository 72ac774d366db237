// dependent-chain latency of DADD / FADD on this GPU (one warp), and of an LDS-fed DADD chain
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x) {
    __shared__ double buf[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) buf[i] = x * i;
    __syncwarp();
    double a = x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) a = __dadd_rn(a, x);
    long long t1 = clock64();
    float f = (float)x;
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) f = __fadd_rn(f, (float)x);
    long long t2 = clock64();
    double b = 0;
#pragma unroll 4
    for (int i = 0; i < 1024; ++i) b = __dadd_rn(b, buf[i]);
    long long t3 = clock64();
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll 4
    for (int i = 0; i < 1024; i += 4) { c0 = __dadd_rn(c0, buf[i]); c1 = __dadd_rn(c1, buf[i + 1]); c2 = __dadd_rn(c2, buf[i + 2]); c3 = __dadd_rn(c3, buf[i + 3]); }
    long long t4 = clock64();
    out[threadIdx.x] = a + f + b + c0 + c1 + c2 + c3;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 256); cudaMalloc(&c, 64);
    for (int lanes : {32, 1}) {
        k<<<1, 32>>>(o, c, 1.5);
        long long h[4]; cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
        printf("DADD chain %.1f cyc/op, FADD chain %.1f, LDS-fed DADD chain %.1f, 4 independent LDS-fed chains %.1f cyc/op\n", h[0] / 1024.0, h[1] / 1024.0, h[2] / 1024.0, h[3] / 1024.0);
    }
}
