// division_check.cu - TEST INFRASTRUCTURE ONLY (never loaded by the product).
// Compares the division forms of citylearn_b200/csrc/unit_physics.cuh (`dvd`, `dvr` with a precomputed `Divisor`) bit for bit
// against the IEEE round-to-nearest division `__ddiv_rn` on the device, over pseudo-random and structured operands.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../citylearn_b200/csrc/unit_physics.cuh"

using namespace cl;

__device__ __forceinline__ uint64_t mix(uint64_t z) {   // splitmix64
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
// a double with a random 52-bit mantissa and an exponent in [-span, span]; sign from `neg`
__device__ __forceinline__ double make(uint64_t bits, int span, bool neg) {
    const uint64_t man = bits & 0x000fffffffffffffull;
    const int e = 1023 + (int)((bits >> 52) % (uint64_t)(2 * span + 1)) - span;
    return __longlong_as_double((long long)(((uint64_t)neg << 63) | ((uint64_t)e << 52) | man));
}

__global__ void check_kernel(uint64_t seed, long n, int mode, unsigned long long* bad, double* first_bad) {
    const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (long i = i0; i < n; i += stride) {
        const uint64_t a = mix(seed + 2 * (uint64_t)i), b = mix(seed + 2 * (uint64_t)i + 1);
        double x, y;
        switch (mode) {
            case 0: x = make(a, 40, a >> 63); y = make(b, 40, false); break;                 // physical magnitudes
            case 1: x = make(a, 480, a >> 63); y = make(b, 390, false); break;               // edge of the fast range
            case 2: x = make(a, 1000, a >> 63); y = make(b, 1000, (b >> 62) & 1); break;      // everything incl. subnormal / huge / negative divisors
            case 3: y = make(b, 20, false); x = y * (double)(int)(a % 4097) ; break;          // exact quotients
            case 4: y = make(b, 20, false); x = (a & 1) ? 0.0 : -0.0; break;                  // zero numerators
            case 5: {                                                                         // quotients next to a rounding tie: x = y * (m + 1/2 ulp-ish)
                y = make(b, 8, false);
                const double m = make(a, 8, false);
                x = fma(y, m, y * 1.1102230246251565e-16 * ((a >> 60) & 1 ? 1.0 : -1.0) * m);
                break;
            }
            default: {                                                                        // curve-like: small integers / decimals
                x = (double)(int)(a % 2001 - 1000) * 1e-3; y = (double)(1 + (int)(b % 1000)) * 1e-3; break;
            }
        }
        const double ref = __ddiv_rn(x, y);
        const double q1 = dvd<double>(x, y);
        const Divisor<double> dv = make_divisor<double>(y);
        const double q2 = dvr<double>(x, dv);
        const bool ok1 = __double_as_longlong(q1) == __double_as_longlong(ref) || (q1 != q1 && ref != ref);
        const bool ok2 = __double_as_longlong(q2) == __double_as_longlong(ref) || (q2 != q2 && ref != ref);
        if (!(ok1 && ok2)) {
            if (local == 0 && atomicAdd(bad, 0ull) == 0ull) { first_bad[0] = x; first_bad[1] = y; first_bad[2] = ref; first_bad[3] = ok1 ? q2 : q1; }
            ++local;
        }
    }
    if (local) atomicAdd(bad, local);
}

// returns the number of mismatching (x, y) pairs among n pseudo-random pairs of `mode`; first_bad_host[4] = x, y, reference, got
extern "C" long long division_check(uint64_t seed, long n, int mode, double* first_bad_host) {
    unsigned long long* bad = nullptr; double* fb = nullptr;
    if (cudaMalloc(&bad, sizeof(*bad)) != cudaSuccess || cudaMalloc(&fb, 4 * sizeof(double)) != cudaSuccess) return -1;
    cudaMemset(bad, 0, sizeof(*bad)); cudaMemset(fb, 0, 4 * sizeof(double));
    check_kernel<<<148 * 8, 256>>>(seed, n, mode, bad, fb);
    unsigned long long h = 0;
    if (cudaMemcpy(&h, bad, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    if (first_bad_host) cudaMemcpy(first_bad_host, fb, 4 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(bad); cudaFree(fb);
    return (long long)h;
}
