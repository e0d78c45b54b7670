// dependent-issue latencies on one warp: DADD, DMUL, SHFL(64-bit), LDS -- to bound the diffusion wavefront's step time
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b) {
    __shared__ double sm[256];
    sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
    __syncwarp();
    double x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 4096; i++) x = x + b;
    long long t1 = clock64();
#pragma unroll 16
    for (int i = 0; i < 4096; i++) x = x * b;
    long long t2 = clock64();
#pragma unroll 16
    for (int i = 0; i < 4096; i++) x = __shfl_up_sync(0xffffffffu, x, 1);
    long long t3 = clock64();
    int idx = threadIdx.x;
#pragma unroll 16
    for (int i = 0; i < 4096; i++) { idx = (int)sm[idx & 63] + (idx & 31); }
    long long t4 = clock64();
    // the diffusion step: ((n + s) + e) + w) * 0.25 with s = previous result and w = shfl_up(previous result)
    double res = a;
#pragma unroll 8
    for (int i = 0; i < 4096; i++) {
        const double nv = sm[(threadIdx.x + i) & 63];
        const double ev = __shfl_down_sync(0xffffffffu, nv, 1);
        const double wv = __shfl_up_sync(0xffffffffu, res, 1);
        res = (((nv + res) + ev) + wv) * 0.25;
        sm[64 + ((threadIdx.x + i) & 63)] = res;
    }
    long long t5 = clock64();
    out[threadIdx.x] = x + idx + res;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 5 * 8);
    k<<<1, 32>>>(out, cyc, 1.0, 1.0000001);
    k<<<1, 32>>>(out, cyc, 1.0, 1.0000001);
    long long h[5];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("cycles per op: DADD %.1f DMUL %.1f SHFL64 %.1f LDS(f64->idx) %.1f diffusion-step %.1f\n", h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0, h[3] / 4096.0, h[4] / 4096.0);
    return 0;
}
