// How fast are the legacy warp-level tensor-core paths that compute a Hamming distance directly on sm_100a?
//   b1 : mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.xor.popc  -> 16 x 8 descriptor pairs of 256 bits per instruction
//   s8 : mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32            -> the +-1 int8 formulation needs 8 of them per 16 x 8 pairs
// Register operands only (no memory), 4 independent accumulator chains per warp.  Prints 10^9 descriptor pairs ("matches") per second,
// to be read against the popc path's measured 466 GMatch/s (all-pairs kernel) and 558 GMatch/s (popc peak).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_b1 mma_b1.cu
#include <cstdint>
#include <cstdio>

template <int KIND>
__global__ void __launch_bounds__(256) k(int *out, int iters) {
    unsigned a[4] = {threadIdx.x * 2654435761u, threadIdx.x * 40503u + 1u, threadIdx.x ^ 0x9e3779b9u, threadIdx.x + 77u};
    unsigned b[2] = {threadIdx.x * 2246822519u, threadIdx.x * 3266489917u};
    int c[4][4] = {};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.xor.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    int s = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) s += c[u][0] + c[u][1] + c[u][2] + c[u][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    const int blocks = 148 * 8, iters = 4000;
    int *o; cudaMalloc(&o, blocks * 256 * sizeof(int));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int kind = 0; kind < 2; ++kind)
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (kind == 0) k<0><<<blocks, 256>>>(o, iters); else k<1><<<blocks, 256>>>(o, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            const double mmas = (double) blocks * 8 * iters * 4;                 // warp-level instructions
            const double pairs = kind == 0 ? mmas * 128 : mmas * 128 / 8;        // 16 x 8 pairs per b1 instruction; 8 s8 instructions per 256 dims
            if (rep) printf("%s: %.3f ms, %.1f G mma/s, %.0f GMatch/s equivalent (%s)\n", kind == 0 ? "b1 m16n8k256 xor.popc" : "s8 m16n8k32", ms,
                            mmas / ms / 1e6, pairs / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
