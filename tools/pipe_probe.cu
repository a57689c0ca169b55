// pipe_probe.cu -- issue-rate microbenchmarks for the instruction forms of the force kernel
// (register-file port pressure of 2- vs 3-operand FP32 ops, FMNMX3, FSETP, MUFU.RSQ) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#define N 12
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float pa, float pb, const float *in)
{
    float v[N], a[N], b[N];
    for (int i = 0; i < N; i++) { v[i] = in[threadIdx.x + i]; a[i] = in[threadIdx.x + 32 + i]; b[i] = in[threadIdx.x + 64 + i]; }
    float c = in[threadIdx.x + 100];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (MODE == 0) v[i] = fmaf(v[i], pa, pb);            // reg, uniform/const, uniform/const
            if (MODE == 1) v[i] = fmaf(a[i], b[i], v[i]);        // 3 distinct registers
            if (MODE == 2) v[i] = fmaf(c, b[i], v[i]);           // one operand shared (reuse cache)
            if (MODE == 3) v[i] = v[i] * a[i];                   // FMUL 2 regs
            if (MODE == 4) v[i] = v[i] + a[i];                   // FADD 2 regs
            if (MODE == 5) v[i] = fmaf(v[i], v[i], a[i]);        // 2 distinct
            if (MODE == 6) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(v[i]), "f"(a[i]), "f"(b[i])); v[i] = r + 1.0f; }  // FMNMX3 + FADD imm
            if (MODE == 7) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v[i])); v[i] = r; }   // MUFU only
            if (MODE == 8) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v[i])); v[i] = fmaf(r, a[i], b[i]); a[i] = fmaf(a[i], c, r); b[i] = fmaf(b[i], c, r); v[i] = fmaf(v[i], c, 1.0f);}  // 1 MUFU : 4 FFMA
            if (MODE == 9) v[i] = fmaf(v[i], a[i], 1.5f);        // 2 regs + imm
        }
    }
    float s = 0;
    for (int i = 0; i < N; i++) s += v[i] + a[i] + b[i];
    if (s == 12345.678f) out[0] = s;
}
template <int MODE> void run(const char *name, int per_iter_ops, float *d, float *in)
{
    const int blocks = 148 * 8, iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, iters, 1.0000001f, 1e-9f, in);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<MODE><<<blocks, 256>>>(d, iters, 1.0000001f, 1e-9f, in);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double warp_inst = (double)per_iter_ops * N * iters * blocks * 8;
    const double cyc = ms * 1e-3 * 1.965e9 * 148 * 4;   // SMSP-cycles at max clock
    printf("%-34s %8.3f ms  %6.3f warp-inst/SMSP-cycle (at 1965 MHz)\n", name, ms, warp_inst / cyc);
}
int main()
{
    float *d, *in; cudaMalloc(&d, 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    run<0>("FFMA reg,const,const", 1, d, in);
    run<1>("FFMA 3 distinct regs", 1, d, in);
    run<2>("FFMA shared operand (reuse)", 1, d, in);
    run<3>("FMUL 2 regs", 1, d, in);
    run<4>("FADD 2 regs", 1, d, in);
    run<5>("FFMA v,v,a (2 distinct)", 1, d, in);
    run<6>("FMNMX3 + FADD", 2, d, in);
    run<7>("MUFU.RSQ chain", 1, d, in);
    run<8>("1 MUFU : 4 FFMA", 5, d, in);
    run<9>("FFMA reg,reg,imm", 1, d, in);
    return 0;
}
