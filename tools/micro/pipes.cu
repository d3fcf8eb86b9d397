// Which pipe executes HMNMX2 / VIMNMX3 / IMAD / LOP3 / IDP on sm_100a?  Run under ncu and read sm__inst_executed_pipe_*.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
template <int OP>
__global__ void k(uint32_t *out, uint32_t seed, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9e3779b9u, c = a + 12345u, d = b + 777u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) {        // HMNMX2
                __half2 x = *reinterpret_cast<__half2 *>(&a), y = *reinterpret_cast<__half2 *>(&b);
                __half2 z = __hmin2(x, y), w = __hmax2(*reinterpret_cast<__half2 *>(&c), *reinterpret_cast<__half2 *>(&d));
                a = *reinterpret_cast<uint32_t *>(&z); c = *reinterpret_cast<uint32_t *>(&w); b += 0x00010001u; d ^= a;
            } else if (OP == 1) { // VIMNMX3 u16x2
                a = __vimin3_u16x2(a, b, c); c = __vimax3_u16x2(c, d, a); b += 0x00010001u; d ^= a;
            } else if (OP == 2) { // IDP4A
                a = __dp4a(a, b, c); c = __dp4a(c, d, a); b += 0x00010001u; d ^= a;
            } else {              // HFMA2
                __half2 x = *reinterpret_cast<__half2 *>(&a), y = *reinterpret_cast<__half2 *>(&b);
                __half2 z = __hfma2(x, y, *reinterpret_cast<__half2 *>(&c)); a = *reinterpret_cast<uint32_t *>(&z); b += 0x00010001u; d ^= a; c += d;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d;
}
int main() {
    uint32_t *o; cudaMalloc(&o, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int op = 0; op < 4; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (op == 0) k<0><<<148 * 8, 256>>>(o, 1, 2000); else if (op == 1) k<1><<<148 * 8, 256>>>(o, 1, 2000);
            else if (op == 2) k<2><<<148 * 8, 256>>>(o, 1, 2000); else k<3><<<148 * 8, 256>>>(o, 1, 2000);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("op %d: %.3f ms, %.1f G (2 ops of interest per unrolled step) lane-ops/s\n", op, ms, 148.0 * 8 * 256 * 2000 * 16 * 2 / ms / 1e6);
        }
    }
    return 0;
}
