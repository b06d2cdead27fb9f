// lat_probe.cu - dependent-issue latency (cycles per op in a serial chain, one warp) of the instructions the fp64 step chain is made of.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/cuda/lat_probe tools/cuda/lat_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define N 512
template <int OP> __global__ void chain(double* out, long long* cyc, double a, double b, float fa, int n) {
    double x = a + threadIdx.x * 1e-9, y = b;
    float f = fa; int iv = threadIdx.x; 
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = a;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < n; ++r) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (OP == 0) x = fma(x, y, y);                                   // DFMA
            if (OP == 1) x = x + y;                                          // DADD
            if (OP == 2) x = x * y;                                          // DMUL
            if (OP == 3) x = (x < y) ? x + 0.0 : y;                          // DSETP + select (min-like), extra DADD to keep the chain
            if (OP == 4) x = (double)(float)x;                               // F2F.F32.F64 + F2F.F64.F32
            if (OP == 5) f = fmaf(f, fa, fa);                                // FFMA
            if (OP == 6) iv = iv * 3 + 1;                                    // IMAD
            if (OP == 7) x = sm[(__double2loint(x) & 63)];                   // LDS.64 dependent address
            if (OP == 8) x = sqrt(x) + y;                                    // fp64 sqrt (+DADD)
            if (OP == 9) x = y / x;                                          // compiler fp64 division (alternates between two values)
            if (OP == 10) x = (x < y) ? x : y;                               // DSETP + 2 FSEL only
            if (OP == 11) { int hi = __double2hiint(x); x = __hiloint2double(hi ^ 1, __double2loint(x)); }   // int op on halves
            if (OP == 12) x = fmin(x, y);                                    // fmin
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + f + iv;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, double a, double b) {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * sizeof(double)); cudaMalloc(&cyc, sizeof(long long));
    const int n = 8;
    chain<OP><<<1, 32>>>(out, cyc, a, b, 0.999f, n);
    chain<OP><<<1, 32>>>(out, cyc, a, b, 0.999f, n);
    long long h = 0; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s %7.2f cycles/op\n", name, (double)h / (n * N));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("DFMA", 0.5, 0.75); run<1>("DADD", 0.5, 1e-3); run<2>("DMUL", 0.5, 1.0000001);
    run<3>("DSETP+sel+DADD", 0.5, 0.75); run<10>("DSETP+2FSEL", 0.5, 0.75); run<12>("fmin(double)", 0.5, 0.75);
    run<4>("F2F f64->f32->f64", 0.5, 0.75);
    run<5>("FFMA", 0.5, 0.75); run<6>("IMAD", 0.5, 0.75); run<7>("LDS.64 dependent", 0.0, 0.0); run<11>("int on halves", 0.5, 0.7);
    run<8>("sqrt(f64)+DADD", 0.5, 0.25); run<9>("div(f64)", 0.7, 0.9);
    return 0;
}
