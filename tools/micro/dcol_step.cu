// the engine's dcol_fast_block in isolation (one warp, no helpers): cycles per step for the three variants
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
#define DCOL_R 4
#define DCOL_ROWS (32 * DCOL_R)
// (the IEEE fallback is a call the compiler cannot hoist or if-convert: speculated, its zero/denormal slow path
// would run for every cell of an unpolluted region)
__device__ static __noinline__ double div3_ieee(double x) { return x / 3.0; }
__device__ __forceinline__ double dcol_lds(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void dcol_sts(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
// x / 3.0 for the fast path: Markstein step in the normal range (also exact for zeros), IEEE division otherwise
__device__ __forceinline__ double dcol_div3(double x) {
    const double r3 = 1.0 / 3.0;
    const double q = x * r3;
    double d = __fma_rn(__fma_rn(-3.0, q, x), r3, q);
    const double ax = fabs(x);
    if (!(ax > 1e-280 && ax < 1e280) && x != 0.0) d = div3_ieee(x);
    return d;
}
struct DcolState {
    double s_prev[DCOL_R], old_cur[DCOL_R], hand[2];
    uint32_t a_row;         // shared address of (my first row, slot 0)
    uint32_t a_above;       // shared address of (row above the band, slot 0)
    uint32_t off_s;         // byte offset of my current column's ring slot
    double* out;            // lane 31: where the band below picks up my last row (column of this step)
};
// 32 steps of the column-chain wavefront for a full band on interior columns.  TOP / BOT: the band holds the
// lattice's first / last row (lane 0's first row has no w, lane 31's last row no e: three neighbours, x/3).
template <int S, int RING, bool TOP, bool BOT>
__device__ __forceinline__ void dcol_fast_block(DcolState& st, int lane) {
    constexpr uint32_t ROWB = RING * 8u;
#pragma unroll 2
    for (int k = 0; k < 32; k++) {
        const uint32_t off_n = st.off_s + 8u == ROWB ? 0u : st.off_s + 8u;
        double w_in = __shfl_up_sync(FULL, st.hand[0], 1);
        double nv[DCOL_R];
#pragma unroll
        for (int r = 0; r < DCOL_R; r++) nv[r] = dcol_lds(st.a_row + r * ROWB + off_n);
        double e_last = 0.0;
        if (!BOT || lane != 31) e_last = dcol_lds(st.a_row + DCOL_R * ROWB + st.off_s);
        if (!TOP) { const double ab = dcol_lds(st.a_above + st.off_s); if (lane == 0) w_in = ab; }
        double t[DCOL_R];
#pragma unroll
        for (int r = 0; r < DCOL_R; r++) {
            t[r] = nv[r] + st.s_prev[r];
            if (r + 1 < DCOL_R) t[r] = t[r] + st.old_cur[r + 1];
        }
        const double t_last_e = t[DCOL_R - 1] + e_last;     // the usual last row: n, s, e (, w)
        double up = w_in;
#pragma unroll
        for (int r = 0; r < DCOL_R; r++) {
            double res;
            if (r + 1 < DCOL_R) {
                res = (t[r] + up) * 0.25;
                if (TOP && r == 0) { const double d = dcol_div3(t[0]); if (lane == 0) res = d; }          // n, s, e
            } else {
                res = (t_last_e + up) * 0.25;
                if (BOT) { const double d = dcol_div3(t[r] + up); if (lane == 31) res = d; }               // n, s, w
            }
            dcol_sts(st.a_row + r * ROWB + st.off_s, res);
            st.s_prev[r] = res;
            st.old_cur[r] = nv[r];
            up = res;
        }
#pragma unroll
        for (int i = 0; i + 1 < S; i++) st.hand[i] = st.hand[i + 1];
        st.hand[S - 1] = up;
        if (!BOT && lane == 31) *st.out = up;
        st.out++;
        st.off_s = off_n;
    }
}

template <bool TOP, bool BOT, bool STORE>
__global__ void k(double* out, double* hand, long long* cyc, double a) {
    extern __shared__ __align__(16) double sm[];
    constexpr int RING = 160;
    for (int i = threadIdx.x; i < 130 * RING; i += 32) sm[i] = a + i * 1e-3;
    __syncwarp();
    const int lane = threadIdx.x;
    DcolState st;
    for (int r = 0; r < 4; r++) { st.s_prev[r] = a; st.old_cur[r] = a; }
    st.hand[0] = st.hand[1] = a;
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(sm);
    st.a_row = tile_s + (uint32_t)((1 + 4 * lane) * RING) * 8u;
    st.a_above = tile_s;
    st.off_s = (uint32_t)((RING - (2 * lane) % RING) % RING) * 8u;
    st.out = hand;
    long long t0 = clock64();
    for (int blk = 0; blk < 64; blk++) {
        if (!STORE) st.out = hand;
        dcol_fast_block<2, RING, TOP, BOT || !STORE>(st, lane);
    }
    long long t1 = clock64();
    out[threadIdx.x] = st.hand[1] + st.s_prev[0];
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <bool TOP, bool BOT, bool STORE> void run(const char* name, double* out, double* hand, long long* cyc) {
    cudaFuncSetAttribute(k<TOP, BOT, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 160 * 8);
    for (int rep = 0; rep < 2; rep++) k<TOP, BOT, STORE><<<1, 32, 130 * 160 * 8>>>(out, hand, cyc, 1.0);
    long long h;
    cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("%s: %.1f cycles per step (%s)\n", name, h / 2048.0, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    double *out, *hand; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&hand, 8 * 4096); cudaMalloc(&cyc, 8);
    run<false, false, true>("middle band, hand-off store", out, hand, cyc);
    run<false, true, true>("last band (no store)", out, hand, cyc);
    run<true, false, true>("first band, hand-off store", out, hand, cyc);
    return 0;
}
