// Dependent-chain latencies of the instructions on the PUCT critical path (one warp, clock64 around N dependent ops).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 2048
__device__ double g_sink;
__device__ unsigned g_isink;
template <int OP>
__global__ void k(long long *out, const double *tab, double x0, unsigned u0) {
    __shared__ uint4 sm[256];
    for (int i = threadIdx.x; i < 256; i += 32) sm[i] = make_uint4((i * 7 + 1) & 255, 0, 0, 0);
    __syncwarp();
    double x = x0, y = 1.0000001;
    unsigned u = u0 + threadIdx.x;
    float f = (float)x0;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = __fma_rn(x, y, 0.5);
        if (OP == 1) x = __dmul_rn(x, y);
        if (OP == 2) x = __dadd_rn(x, y);
        if (OP == 3) { double o = __longlong_as_double(__double_as_longlong(x) ^ 1); x = (o > x) ? o : __dadd_rn(x, 0.0) * 1.0 == x ? x : o; }
        if (OP == 4) { double o = __shfl_xor_sync(0xffffffffu, x, 4); x = o; }
        if (OP == 5) u = __shfl_sync(0xffffffffu, u, (u + 1) & 31);
        if (OP == 6) u = sm[u & 255].x;
        if (OP == 7) u = (unsigned)__double_as_longlong(__ldg(tab + (u & 127)));
        if (OP == 8) u = __ballot_sync(0xffffffffu, u & 1) + u;
        if (OP == 9) u = __ffs(u | 0x80000000u) + u;
        if (OP == 10) u = __popc(u) + u;
        if (OP == 11) { x = (double)f; f = (float)u + __double2float_rn(x); u++; }
        if (OP == 12) x = (double)u + x * 0.0, u = (unsigned)__double2int_rn(x) + 1;
        if (OP == 13) u = u * 3 + 1;
        if (OP == 14) { bool p = x > y; x = p ? y : x; y = p ? x + 1.0 : y; }
        if (OP == 15) f = fmaf(f, 1.0001f, 0.5f);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[OP] = t1 - t0;
    g_sink = x + y + f;
    g_isink = u;
}
int main() {
    long long *out; double *tab;
    cudaMallocManaged(&out, 16 * sizeof(long long));
    cudaMalloc(&tab, 128 * 8);
    cudaMemset(tab, 0, 128 * 8);
    const char *names[16] = {"DFMA", "DMUL", "DADD", "(mixed dsetp)", "SHFL.BFLY f64 (2 shfl)", "SHFL.IDX + iadd/and", "LDS.128 + and", "LDG.CONSTANT L1 + and + imad.wide",
                             "VOTE + lop + iadd", "FLO/ffs + 2 alu", "POPC + iadd", "F2F.F64.F32 + F2F.F32.F64 + I2F + FADD", "I2F.F64 + DFMA + F2I + iadd", "IMAD", "DSETP + FSEL x2 + DADD", "FFMA"};
#define RUN(i) k<i><<<1, 32>>>(out, tab, 1.5, 3u); k<i><<<1, 32>>>(out, tab, 1.5, 3u);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15)
    cudaDeviceSynchronize();
    for (int i = 0; i < 16; ++i) printf("%-45s %7.1f cycles per iteration\n", names[i], (double)out[i] / N);
    return 0;
}
