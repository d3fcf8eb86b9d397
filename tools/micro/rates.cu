// Issue-rate micro-benchmark (sm_100a): warp instructions per clock and SM for the integer / SIMD instructions the kernels lean on.
// Eight independent chains per thread, 1024 threads per SM.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rates rates.cu
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#define CH 8
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, int iters, uint32_t one) {
    uint32_t a[CH], b = threadIdx.x * 2654435761u + seed, c = b ^ 0x9e3779b9u;
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = b * (i + 3) + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                // asm volatile keeps one instruction per chain step (the compiler otherwise fuses consecutive min / add steps)
                if (OP == 0) { if (u & 1) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); else asm volatile("max.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c)); }   // VIMNMX.U32 (min / max alternate: no 3-input fusion)
                else if (OP == 1) { if (u & 1) asm volatile("min.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); else asm volatile("max.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(c)); }   // VIMNMX.U16x2
                else if (OP == 2) a[i] = __vimin3_u16x2(a[i], b, c);                                                     // VIMNMX3.U16x2
                else if (OP == 3) a[i] = __vimin3_u32(a[i], b, c);                                                       // VIMNMX3.U32
                else if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(a[i]) : "r"(b), "r"(c));          // LOP3
                else if (OP == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));                             // IADD3
                else if (OP == 6) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(one), "r"(b));            // IMAD
                else if (OP == 7) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));                // PRMT
                else if (OP == 8) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(0));   // VABSDIFF4
                else if (OP == 9) asm volatile("shf.r.wrap.b32 %0, %0, %1, 8;" : "+r"(a[i]) : "r"(b));                   // SHF
                else if (OP == 10) { if (u & 1) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); else asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(c)); }   // HMNMX2
                else if (OP == 11) { if (u & 1) asm volatile("min.f32 %0, %0, %1;" : "+f"(*reinterpret_cast<float *>(&a[i])) : "f"(__uint_as_float(b))); else asm volatile("max.f32 %0, %0, %1;" : "+f"(*reinterpret_cast<float *>(&a[i])) : "f"(__uint_as_float(c))); }   // FMNMX
                else if (OP == 12) { uint32_t t; asm volatile("popc.b32 %0, %1;" : "=r"(t) : "r"(a[i])); a[i] = t; }      // POPC
                else if (OP == 13) { asm volatile("min.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(one), "r"(c)); }
                else if (OP == 14) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(a[i]) : "r"(b), "r"(c)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(one), "r"(c)); }
                else if (OP == 15) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));            // IDP4A
                else if (OP == 16) { if (u & 1) asm volatile("min.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); else asm volatile("max.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(a[i]) : "r"(b), "r"(c)); }   // VIMNMX.U16x2 + LOP3
            }
            b += 0x00010001u; c ^= b;
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int OP> float run(uint32_t *o, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<OP><<<148 * 4, 256>>>(o, 1, iters, 1);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}
int main() {
    uint32_t *o; cudaMalloc(&o, 148 * 4 * 256 * 4);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 4000;
    const char *names[] = {"VIMNMX.U32", "VIMNMX.U16x2", "VIMNMX3.U16x2", "VIMNMX3.U32", "LOP3", "IADD3", "IMAD", "PRMT", "VABSDIFF4", "SHF", "HMNMX2", "FMNMX",
                           "POPC", "VIMNMX.U16x2+IMAD (2 instr/step)", "LOP3+IMAD (2 instr/step)", "IDP4A", "VIMNMX.U16x2+LOP3 (2 instr/step)"};
    float ms[17];
    ms[0] = run<0>(o, iters); ms[1] = run<1>(o, iters); ms[2] = run<2>(o, iters); ms[3] = run<3>(o, iters); ms[4] = run<4>(o, iters); ms[5] = run<5>(o, iters);
    ms[6] = run<6>(o, iters); ms[7] = run<7>(o, iters); ms[8] = run<8>(o, iters); ms[9] = run<9>(o, iters); ms[10] = run<10>(o, iters); ms[11] = run<11>(o, iters);
    ms[12] = run<12>(o, iters); ms[13] = run<13>(o, iters); ms[14] = run<14>(o, iters); ms[15] = run<15>(o, iters); ms[16] = run<16>(o, iters);
    for (int op = 0; op < 17; ++op) {
        const double steps = (double) iters * 4 * CH;                         // chain steps per thread
        const double warp_steps_per_sm = steps * (4 * 256 / 32);
        printf("%-20s %.3f ms  %.2f chain steps / clk / SM (at the nominal %d MHz)\n", names[op], ms[op], warp_steps_per_sm / (ms[op] * 1e-3 * clk * 1e3), clk / 1000);
    }
    return 0;
}
