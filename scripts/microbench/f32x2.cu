// Throughput of scalar vs packed fp32 instructions on sm_100a (one number per variant: warp-instructions per
// cycle per SM at 32 resident warps).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int V> __global__ void k(float* out, int iters, float s) {
    float2 a0 = make_float2(threadIdx.x, 1.f), a1 = make_float2(2.f, threadIdx.x), a2 = make_float2(3.f, 4.f), a3 = make_float2(5.f, 6.f);
    float2 a4 = make_float2(threadIdx.x, 7.f), a5 = make_float2(8.f, threadIdx.x), a6 = make_float2(9.f, 4.f), a7 = make_float2(5.f, 1.f);
    const float2 b = make_float2(s, s * 0.5f), c = make_float2(s * 0.25f, s);
    for (int i = 0; i < iters; ++i) {
        if (V == 0) {   // 16 scalar FFMA (3 register operands)
            a0.x = fmaf(a0.x, b.x, c.x); a0.y = fmaf(a0.y, b.y, c.y); a1.x = fmaf(a1.x, b.x, c.x); a1.y = fmaf(a1.y, b.y, c.y);
            a2.x = fmaf(a2.x, b.x, c.x); a2.y = fmaf(a2.y, b.y, c.y); a3.x = fmaf(a3.x, b.x, c.x); a3.y = fmaf(a3.y, b.y, c.y);
            a4.x = fmaf(a4.x, b.x, c.x); a4.y = fmaf(a4.y, b.y, c.y); a5.x = fmaf(a5.x, b.x, c.x); a5.y = fmaf(a5.y, b.y, c.y);
            a6.x = fmaf(a6.x, b.x, c.x); a6.y = fmaf(a6.y, b.y, c.y); a7.x = fmaf(a7.x, b.x, c.x); a7.y = fmaf(a7.y, b.y, c.y);
        } else if (V == 1) {   // 8 FFMA2
            a0 = __ffma2_rn(a0, b, c); a1 = __ffma2_rn(a1, b, c); a2 = __ffma2_rn(a2, b, c); a3 = __ffma2_rn(a3, b, c);
            a4 = __ffma2_rn(a4, b, c); a5 = __ffma2_rn(a5, b, c); a6 = __ffma2_rn(a6, b, c); a7 = __ffma2_rn(a7, b, c);
        } else if (V == 2) {   // 8 FADD2
            a0 = __fadd2_rn(a0, b); a1 = __fadd2_rn(a1, b); a2 = __fadd2_rn(a2, b); a3 = __fadd2_rn(a3, b);
            a4 = __fadd2_rn(a4, b); a5 = __fadd2_rn(a5, b); a6 = __fadd2_rn(a6, b); a7 = __fadd2_rn(a7, b);
        } else if (V == 3) {   // 16 scalar FADD
            a0.x += b.x; a0.y += b.y; a1.x += b.x; a1.y += b.y; a2.x += b.x; a2.y += b.y; a3.x += b.x; a3.y += b.y;
            a4.x += b.x; a4.y += b.y; a5.x += b.x; a5.y += b.y; a6.x += b.x; a6.y += b.y; a7.x += b.x; a7.y += b.y;
        } else if (V == 4) {   // 8 FFMA2 of the form d = a*a + d (the distance chain)
            a0 = __ffma2_rn(a1, a1, a0); a2 = __ffma2_rn(a3, a3, a2); a4 = __ffma2_rn(a5, a5, a4); a6 = __ffma2_rn(a7, a7, a6);
            a1 = __ffma2_rn(a0, a0, a1); a3 = __ffma2_rn(a2, a2, a3); a5 = __ffma2_rn(a4, a4, a5); a7 = __ffma2_rn(a6, a6, a7);
        } else if (V == 5) {   // one fence.proxy.async per 16 FFMA
            a0.x = fmaf(a0.x, b.x, c.x); a0.y = fmaf(a0.y, b.y, c.y); a1.x = fmaf(a1.x, b.x, c.x); a1.y = fmaf(a1.y, b.y, c.y);
            a2.x = fmaf(a2.x, b.x, c.x); a2.y = fmaf(a2.y, b.y, c.y); a3.x = fmaf(a3.x, b.x, c.x); a3.y = fmaf(a3.y, b.y, c.y);
            a4.x = fmaf(a4.x, b.x, c.x); a4.y = fmaf(a4.y, b.y, c.y); a5.x = fmaf(a5.x, b.x, c.x); a5.y = fmaf(a5.y, b.y, c.y);
            a6.x = fmaf(a6.x, b.x, c.x); a6.y = fmaf(a6.y, b.y, c.y); a7.x = fmaf(a7.x, b.x, c.x); a7.y = fmaf(a7.y, b.y, c.y);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0.x + a0.y + a1.x + a1.y + a2.x + a2.y + a3.x + a3.y + a4.x + a4.y + a5.x + a5.y + a6.x + a6.y + a7.x + a7.y;
}
template <int V> void run(const char* name, int instr_per_iter) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V><<<148 * 8, 128>>>(out, 100, 1.0001f);
    cudaEventRecord(e0);
    k<V><<<148 * 8, 128>>>(out, iters, 1.0001f);   // 8 blocks x 4 warps = 32 warps per SM
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = 148.0 * 32 * iters * instr_per_iter;
    const double cycles = ms * 1e-3 * 1.965e9;
    printf("%-28s %8.3f ms  %.3f warp-instr/cycle/SM (%d per iteration)\n", name, ms, warp_instr / cycles / 148.0, instr_per_iter);
    cudaFree(out);
}
int main() {
    run<0>("16 FFMA", 16);
    run<1>("8 FFMA2 (b, c shared)", 8);
    run<2>("8 FADD2", 8);
    run<3>("16 FADD", 16);
    run<4>("8 FFMA2 a*a+d", 8);
    run<5>("16 FFMA + fence.proxy.async", 17);
    return 0;
}
