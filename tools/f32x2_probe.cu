// f32x2_probe.cu -- issue/pipe-rate microbenchmarks of Blackwell's packed FP32 instructions
// (PTX fma/mul/add.rn.f32x2 -> SASS FFMA2/FMUL2/FADD2) alone and mixed with the scalar
// instructions of the force kernel (FMNMX3, MUFU.RSQ, LDS, FSETP).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_probe f32x2_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float rsq(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
#define N 8
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, const float *in)
{
    __shared__ float4 sm[256];
    sm[threadIdx.x] = make_float4(in[threadIdx.x], 1.f, 2.f, 3.f);
    __syncthreads();
    u64 v[N], a[N], b[N];
    float s[N], t[N];
    for (int i = 0; i < N; i++) {
        v[i] = pk(in[threadIdx.x + i], in[threadIdx.x + i + 1]); a[i] = pk(in[threadIdx.x + 32 + i], in[threadIdx.x + 33 + i]);
        b[i] = pk(in[threadIdx.x + 64 + i], in[threadIdx.x + 65 + i]); s[i] = in[threadIdx.x + 96 + i]; t[i] = in[threadIdx.x + 128 + i];
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (MODE == 0) v[i] = fma2(a[i], b[i], v[i]);                       // FFMA2 only
            if (MODE == 1) v[i] = mul2(v[i], a[i]);                             // FMUL2 only
            if (MODE == 2) v[i] = add2(v[i], a[i]);                             // FADD2 only
            if (MODE == 3) { v[i] = fma2(a[i], b[i], v[i]); s[i] = max3(s[i], t[i], 1.5f); }          // FFMA2 + FMNMX3
            if (MODE == 4) { v[i] = fma2(a[i], b[i], v[i]); s[i] = fmaf(s[i], t[i], 1.5f); }          // FFMA2 + FFMA
            if (MODE == 5) { v[i] = fma2(a[i], b[i], v[i]); a[i] = fma2(v[i], b[i], a[i]); s[i] = rsq(s[i]); } // 2 FFMA2 + MUFU
            if (MODE == 6) { v[i] = fma2(a[i], b[i], v[i]); a[i] = fma2(v[i], b[i], a[i]);
                             const float4 q = sm[(threadIdx.x + it + i) & 255]; s[i] += q.x; }         // 2 FFMA2 + LDS.128 + FADD
            if (MODE == 7) s[i] = fmaf(s[i], t[i], 1.5f);                                               // FFMA scalar only
            if (MODE == 8) { float x, y; upk(v[i], x, y); x = rsq(x); y = rsq(y); v[i] = fma2(pk(x, y), a[i], b[i]); } // MUFUx2 -> pack -> FFMA2
            if (MODE == 9) { v[i] = fma2(a[i], b[i], v[i]); s[i] = max3(s[i], t[i], 1.5f); t[i] = max3(t[i], s[i], 2.5f); } // FFMA2 + 2 FMNMX3
            if (MODE == 10) { v[i] = fma2(a[i], pk(s[i], s[i]), v[i]); }                                // broadcast operand
        }
    }
    float acc = 0;
    for (int i = 0; i < N; i++) { float x, y; upk(v[i], x, y); acc += x + y + s[i] + t[i]; upk(a[i], x, y); acc += x + y; }
    if (acc == 12345.678f) out[0] = acc;
}
template <int MODE> void run(const char *name, double f32_lane_ops, double warp_inst, float *d, float *in)
{
    const int blocks = 148 * 8, iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, iters, in);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<MODE><<<blocks, 256>>>(d, iters, in);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double n = (double)N * iters * blocks * 8;         // per-warp op groups
    const double cyc = ms * 1e-3 * 1.965e9 * 148 * 4;       // SMSP-cycles at max clock
    printf("%-40s %8.3f ms  %6.3f warp-inst/SMSP-clk  %6.3f fp32-lane-ops/lane-clk  (%5.1f TFLOP/s if all FMA)\n", name, ms,
           warp_inst * n / cyc, f32_lane_ops * n / cyc, f32_lane_ops * n * 32 * 2 / (ms * 1e-3) / 1e12);
}
int main()
{
    float *d, *in; cudaMalloc(&d, 4); cudaMalloc(&in, 8192); cudaMemset(in, 0, 8192);
    run<7>("FFMA scalar", 1, 1, d, in);
    run<0>("FFMA2", 2, 1, d, in);
    run<1>("FMUL2", 2, 1, d, in);
    run<2>("FADD2", 2, 1, d, in);
    run<3>("FFMA2 + FMNMX3", 2, 2, d, in);
    run<9>("FFMA2 + 2 FMNMX3", 2, 3, d, in);
    run<4>("FFMA2 + FFMA", 3, 2, d, in);
    run<5>("2 FFMA2 + MUFU", 4, 3, d, in);
    run<6>("2 FFMA2 + LDS.128 + FADD", 5, 4, d, in);
    run<8>("2 MUFU -> FFMA2", 2, 3, d, in);
    run<10>("FFMA2 with broadcast operand", 2, 1, d, in);
    return 0;
}
