// Micro-benchmark: issue cost of the packed / scalar fp32 and ALU instructions the encode kernels are made of, at the
// occupancy the BC7 kernel runs at (1 CTA x 384 threads per SM = 3 warps per scheduler) and at 16 warps per SM.
// Prints SM cycles per warp-instruction per scheduler (SMSP).   nvcc -arch=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define CHAINS 8

__constant__ float c_magic = 12582912.0f;
__constant__ float2 c_magic2 = { 12582912.0f, 12582912.0f };

template<int KIND>
__global__ void k(float *out, unsigned long long *cycles, float a0, float b0)
{
    unsigned long long r[CHAINS];
    unsigned long long a, b, c;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a0 * 0.5f));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b0 * 0.25f));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(b0 + 1.0f), "f"(a0 + 2.0f));
    float s[CHAINS];
    double dd[CHAINS];
    const double da = (double)a0 * 1.0000001, db = (double)b0 * 0.3;
    for (int i = 0; i < CHAINS; i++)
        dd[i] = (double)threadIdx.x + i;
    for (int i = 0; i < CHAINS; i++)
    {
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r[i]) : "f"(a0 + i + threadIdx.x), "f"(b0 - i - threadIdx.x));
        s[i] = a0 * i + threadIdx.x;
    }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++)
    {
#pragma unroll
        for (int i = 0; i < CHAINS; i++)
        {
            if (KIND == 0)       // FFMA2, three 64-bit register operands
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(a), "l"(b));
            else if (KIND == 1)  // FADD2 reg + reg
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(a));
            else if (KIND == 2)  // scalar FFMA 3-reg
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(a0), "f"(b0));
            else if (KIND == 3)  // FFMA2 + FMNMX pair
            {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(a), "l"(b));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(a0));
            }
            else if (KIND == 4)  // FMNMX only
                asm volatile("min.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(a0));
            else if (KIND == 5)  // FFMA2 acc form: d = x*x + d (two distinct regs)
                asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(r[i]) : "l"(a));
            else if (KIND == 6)  // scalar FADD
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(a0));
            else if (KIND == 7)  // FFMA2 + FADD2 alternating (different chains)
            {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(a), "l"(b));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(c) : "l"(r[(i + 4) % CHAINS]));
            }
            else if (KIND == 8)  // LOP3 / integer ALU
            {
                unsigned int &u = reinterpret_cast<unsigned int &>(s[i]);
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(__float_as_uint(a0) + threadIdx.x), "r"(__float_as_uint(b0)));
            }
            else if (KIND == 9)  // FFMA2 + LOP3
            {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(a), "l"(b));
                unsigned int &u = reinterpret_cast<unsigned int &>(s[i]);
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(__float_as_uint(a0) + threadIdx.x), "r"(__float_as_uint(b0)));
            }
            else if (KIND == 11) // FMUL2 reg, reg
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(a));
            else if (KIND == 12) // FFMA2 with an immediate addend
                asm volatile("{ .reg .b64 k; mov.b64 k, {0f4B400000, 0f4B400000}; fma.rn.f32x2 %0, %0, %1, k; }" : "+l"(r[i]) : "l"(a));
            else if (KIND == 13) // FADD2 with an immediate
                asm volatile("{ .reg .b64 k; mov.b64 k, {0f4B400000, 0f4B400000}; add.rn.f32x2 %0, %0, k; }" : "+l"(r[i]));
            else if (KIND == 14) // FFMA2 + PRMT
            {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(a), "l"(b));
                unsigned int &u = reinterpret_cast<unsigned int &>(s[i]);
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(u) : "r"(__float_as_uint(a0) + threadIdx.x), "r"(__float_as_uint(b0)));
            }
            else if (KIND == 15) // FFMA2 + IADD (VIADD / IADD3)
            {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(a), "l"(b));
                unsigned int &u = reinterpret_cast<unsigned int &>(s[i]);
                asm volatile("add.u32 %0, %0, %1;" : "+r"(u) : "r"(__float_as_uint(a0) + threadIdx.x));
            }
            else if (KIND == 16) // FADD2 + FMNMX
            {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(a));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(a0));
            }
            else if (KIND == 17) // FADD2 + FADD2 with scalar-broadcast operand {x, x}
            {
                unsigned long long bb;
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(s[i]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(bb));
            }
            else if (KIND == 18) // FFMA2 + 2 FMNMX
            {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(a), "l"(b));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(a0));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(b0));
            }
            else if (KIND == 19) // scalar FFMA + FMNMX
            {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(a0), "f"(b0));
                float &t = reinterpret_cast<float &>(r[i]);
                asm volatile("min.f32 %0, %0, %1;" : "+f"(t) : "f"(a0));
            }
            else if (KIND == 20) // scalar FFMA + LOP3
            {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(a0), "f"(b0));
                unsigned int &u = reinterpret_cast<unsigned int &>(r[i]);
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(__float_as_uint(a0) + threadIdx.x), "r"(__float_as_uint(b0)));
            }
            else if (KIND == 21) // FADD2 with a scalar operand read from constant memory
            {
                const float m = c_magic;
                unsigned long long bb;
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(m));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(bb));
            }
            else if (KIND == 22) // FADD2 with a register pair holding the constant (loaded once through an opaque asm)
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(c));
            else if (KIND == 23) // FFMA2 reg, scalar.F32 multiplier, reg
            {
                unsigned long long bb;
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(s[i]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(bb), "l"(b));
            }
            else if (KIND == 24) // DFMA alone
            {
                double &d = reinterpret_cast<double &>(r[i]);
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(da), "d"(db));
            }
            else if (KIND == 25) // FFMA2 + DFMA
            {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(a), "l"(b));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(da), "d"(db));
            }
            else if (KIND == 26) // 2 FFMA2 + DFMA
            {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(a), "l"(b));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(c) : "l"(a), "l"(r[(i + 3) % CHAINS]));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(da), "d"(db));
            }
            else if (KIND == 27) // F2F f32 -> f64
            {
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(dd[i]) : "f"(s[i]));
                s[i] += 1.0f;
            }
            else if (KIND == 10) // 2 FFMA2 + 1 scalar FFMA
            {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(a), "l"(b));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(a0), "f"(b0));
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < CHAINS; i++)
    {
        float x, y;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(r[i]));
        acc += x + y + s[i] + (float)dd[i];
    }
    float x, y;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(c));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + x + y;
    if (threadIdx.x == 0)
        cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template<int KIND>
void run(const char *name, int perIter, int threads)
{
    float *out;
    unsigned long long *cyc, h[148];
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    k<KIND><<<148, threads>>>(out, cyc, 1.0001f, 0.9999f);
    k<KIND><<<148, threads>>>(out, cyc, 1.0001f, 0.9999f);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += (double)h[i];
    avg /= 148;
    const double warpsPerSmsp = threads / 32 / 4.0;
    const double instrPerSmsp = (double)ITER * CHAINS * perIter * warpsPerSmsp;
    printf("%-44s threads %4d: %.3f cycles per warp-instruction per SMSP\n", name, threads, avg / instrPerSmsp);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    for (int threads : { 384 })
    {
        run<0>("FFMA2 (3 x 64-bit regs)", 1, threads);
        run<5>("FFMA2 d = x*x + d", 1, threads);
        run<1>("FADD2 reg, reg", 1, threads);
        run<2>("FFMA scalar 3-reg", 1, threads);
        run<6>("FADD scalar", 1, threads);
        run<4>("FMNMX", 1, threads);
        run<8>("LOP3", 1, threads);
        run<3>("FFMA2 + FMNMX (per instruction)", 2, threads);
        run<9>("FFMA2 + LOP3 (per instruction)", 2, threads);
        run<7>("FFMA2 + FADD2 (per instruction)", 2, threads);
        run<10>("FFMA2 + FFMA scalar (per instruction)", 2, threads);
        run<24>("DFMA", 1, threads);
        run<25>("FFMA2 + DFMA (per instruction)", 2, threads);
        run<26>("2 FFMA2 + DFMA (per instruction)", 3, threads);
        run<27>("F2F.F64.F32 (+ FADD)", 2, threads);
        run<21>("FADD2 reg, c[magic] scalar", 1, threads);
        run<22>("FADD2 reg, reg pair (hoisted constant)", 1, threads);
        run<23>("FFMA2 reg, scalar.F32, reg", 1, threads);
        run<11>("FMUL2 reg, reg", 1, threads);
        run<12>("FFMA2 reg, reg, imm", 1, threads);
        run<13>("FADD2 reg, imm", 1, threads);
        run<14>("FFMA2 + PRMT (per instruction)", 2, threads);
        run<15>("FFMA2 + IADD (per instruction)", 2, threads);
        run<16>("FADD2 + FMNMX (per instruction)", 2, threads);
        run<17>("FADD2 reg, scalar.F32", 1, threads);
        run<18>("FFMA2 + 2 FMNMX (per instruction)", 3, threads);
        run<19>("FFMA + FMNMX scalar (per instruction)", 2, threads);
        run<20>("FFMA scalar + LOP3 (per instruction)", 2, threads);
    }
    return 0;
}
