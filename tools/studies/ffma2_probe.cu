// Does FFMA2 (fma.rn.f32x2, sm_100) halve the ISSUE cost of paired FP32 work?  Four variants, same arithmetic:
//   s0: 8 FFMA / iter            p0: 4 FFMA2 / iter
//   s1: 8 FFMA + 8 LOP3 / iter   p1: 4 FFMA2 + 8 LOP3 / iter
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, unsigned* iout, int iters) {
    float a[8]; unsigned u[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { a[j] = threadIdx.x * 1e-3f + j; u[j] = threadIdx.x * 2654435761u + j; }
    const float b = 0.9999f, c = 1e-4f;
    float2 p[4];
#pragma unroll
    for (int j = 0; j < 4; j++) p[j] = make_float2(a[2 * j], a[2 * j + 1]);
    const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (MODE == 0 || MODE == 2) {
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = __fmaf_rn(a[j], b, c);
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) p[j] = __ffma2_rn(p[j], b2, c2);
            }
            if (MODE >= 2) {
#pragma unroll
                for (int j = 0; j < 8; j++) u[j] = ((u[j] ^ (u[(j + 1) & 7] >> 3)) & 0x7fffffffu) ^ 0x9E3779B9u;
            }
        }
    }
    float r = 0; unsigned v = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { r += a[j]; v ^= u[j]; }
#pragma unroll
    for (int j = 0; j < 4; j++) r += p[j].x + p[j].y;
    if (r == 12345.678f) out[0] = r;
    if (v == 0x12345678u) iout[0] = v;
}
template <int MODE> float run(int iters) {
    float* o; unsigned* io; cudaMalloc(&o, 16); cudaMalloc(&io, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(o, io, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    return best;
}
int main() {
    const int iters = 4096;
    const double fl = 2.0 * 8 * 8 * iters * 256.0 * 148 * 8;
    float t0 = run<0>(iters), t1 = run<1>(iters), t2 = run<2>(iters), t3 = run<3>(iters);
    printf("s0 8 FFMA        : %.3f ms  %.1f TFLOP/s\n", t0, fl / t0 * 1e-9);
    printf("p0 4 FFMA2       : %.3f ms  %.1f TFLOP/s\n", t1, fl / t1 * 1e-9);
    printf("s1 8 FFMA +8 int : %.3f ms  %.1f TFLOP/s\n", t2, fl / t2 * 1e-9);
    printf("p1 4 FFMA2+8 int : %.3f ms  %.1f TFLOP/s\n", t3, fl / t3 * 1e-9);
    return 0;
}
