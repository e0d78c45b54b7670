// FP64 issue throughput of one warp (independent chains) and the cost of the column-chain diffusion step in isolation
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 130 * 160; i += 32) sm[i] = a + i * 1e-3;
    __syncwarp();
    double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 2048; i++) { x0 += b; x1 += b; x2 += b; x3 += b; x4 += b; x5 += b; x6 += b; x7 += b; }
    long long t1 = clock64();
    // column-chain step: 4 rows per lane, ring of 160 columns, skew 2
    const int lane = threadIdx.x;
    double s_prev[4] = {a, a, a, a}, old_cur[4] = {a, a, a, a}, hand0 = a, hand1 = a;
    int ys = (160 - (2 * lane) % 160) % 160;
    const int row0 = 1 + 4 * lane;
    for (int k = 0; k < 2048; k++) {
        const int yn = ys + 1 == 160 ? 0 : ys + 1;
        double w_in = __shfl_up_sync(0xffffffffu, hand0, 1);
        if (lane == 0) w_in = sm[ys];
        double nv[4], t[4];
#pragma unroll
        for (int r = 0; r < 4; r++) nv[r] = sm[(row0 + r) * 160 + yn];
        const double e_last = sm[(row0 + 4) * 160 + ys];
#pragma unroll
        for (int r = 0; r < 4; r++) t[r] = (nv[r] + s_prev[r]) + (r < 3 ? old_cur[r + 1] : e_last);
        double up = w_in;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const double res = (t[r] + up) * 0.25;
            sm[(row0 + r) * 160 + ys] = res;
            s_prev[r] = res; old_cur[r] = nv[r]; up = res;
        }
        hand0 = hand1; hand1 = up; ys = yn;
    }
    long long t2 = clock64();
    out[threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + hand1;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 2 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 160 * 8);
    for (int rep = 0; rep < 2; rep++) k<<<1, 32, 130 * 160 * 8>>>(out, cyc, 1.0, 1.0000001);
    long long h[2];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("8 independent DADD chains: %.1f cycles per 8 DADDs (%.2f per DADD); column-chain step: %.1f cycles\n", h[0] / 2048.0, h[0] / 2048.0 / 8, h[1] / 2048.0);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
