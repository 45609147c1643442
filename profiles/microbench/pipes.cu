// Latency / throughput of the instructions k_deskew_bin is made of (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int OP, int CHAINS>
__global__ void k(float* out, long long* cyc, float seed) {
    float f[CHAINS];
    double d[CHAINS];
    unsigned long long p[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        f[i] = seed + i + threadIdx.x;
        d[i] = seed + i + threadIdx.x;
        p[i] = ((unsigned long long) __float_as_uint(f[i]) << 32) | __float_as_uint(f[i] + 1.f);
    }
    const float a = seed * 0.999f, b = seed * 1e-3f;
    const double da = a, db = b;
    const unsigned long long pa = ((unsigned long long) __float_as_uint(a) << 32) | __float_as_uint(a);
    const unsigned long long pb = ((unsigned long long) __float_as_uint(b) << 32) | __float_as_uint(b);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (OP == 0) f[i] = fmaf(f[i], a, b);
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
            if (OP == 2) d[i] = fma(d[i], da, db);
            if (OP == 3) { d[i] = (double) f[i]; f[i] = (float) d[i]; } // F2F.F64.F32 + F2F.F32.F64
            if (OP == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
            if (OP == 5) d[i] = d[i] * da;
            if (OP == 6) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
            if (OP == 7) { d[i] = fma(d[i], da, db); f[i] = fmaf(f[i], a, b); }                       // DFMA + FFMA, independent
            if (OP == 8) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i])); d[i] = fma(d[i], da, db); } // MUFU + DFMA
            if (OP == 9) { d[i] = fma(d[i], da, db); p[i] = p[i] * 0x9E3779B97F4A7C15ull + (p[i] >> 7); } // DFMA + integer work
            if (OP == 10) { f[i] = fmaf(f[i], a, b); p[i] = (p[i] ^ (p[i] >> 7)) + pa; }                // FFMA + ALU
        }
    }
    const long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc += f[i] + (float) d[i] + (float) p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP, int CHAINS>
void run(const char* name, int warps, float* out, long long* cyc) {
    k<OP, CHAINS><<<148, warps * 32>>>(out, cyc, 1.0001f);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double) c / ITERS;
    const int ops_per_iter = CHAINS * ((OP == 3 || OP >= 7) ? 2 : 1);
    printf("%-22s chains %d warps/SM %2d: %7.2f cyc/iter  -> %.2f cyc per dependent op, %.3f warp-instr/cyc/SMSP\n", name, CHAINS, warps, per,
           per / (OP == 3 ? 2 : 1), (double) ops_per_iter * warps / 4.0 / per);
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 8);
#define ALL(OP, NAME)               \
    run<OP, 1>(NAME, 1, out, cyc);  \
    run<OP, 1>(NAME, 4, out, cyc);  \
    run<OP, 8>(NAME, 4, out, cyc);  \
    run<OP, 8>(NAME, 16, out, cyc);
    ALL(0, "FFMA")
    ALL(1, "FFMA2 (f32x2)")
    ALL(6, "FMUL2 (f32x2)")
    ALL(2, "DFMA")
    ALL(5, "DMUL")
    ALL(3, "F2F f32->f64->f32")
    ALL(4, "MUFU.RCP")
    ALL(7, "DFMA+FFMA mixed")
    ALL(8, "MUFU+DFMA mixed")
    ALL(9, "DFMA+int mixed")
    ALL(10, "FFMA+ALU mixed")
    return 0;
}
