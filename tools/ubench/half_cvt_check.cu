// Checks half_bits_to_float (hardware conversion + the exponent-0 correction) against the reference's TwosCLHalfToFloat
// formula on the device, all 65536 patterns.  nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false half_cvt_check.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include "../../convectionkernels_b200/csrc/bc6h_core.cuh"
using namespace cvttb200;
__global__ void k(uint32_t *bad, uint32_t *first)
{
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= 65536) return;
    const float a = twoscl_half_to_float((int)u);
    const float b = half_bits_to_float(u);                                       // what the kernels use
    const float c = __half2float(__ushort_as_half((unsigned short)u));           // plain IEEE: differs for exponent 0 (twice the value)
    const bool inDomain = (u & 0x7c00u) != 0x7c00u;
    if (inDomain && !(a == b))
    {
        const uint32_t i = atomicAdd(bad, 1u);
        if (i < 16) { first[i * 4] = u; first[i * 4 + 1] = __float_as_uint(a); first[i * 4 + 2] = __float_as_uint(b); first[i * 4 + 3] = __float_as_uint(c); }
    }
}
int main()
{
    uint32_t *bad, *first;
    cudaMallocManaged(&bad, 4); cudaMallocManaged(&first, 16 * 16);
    *bad = 0;
    k<<<256, 256>>>(bad, first);
    cudaDeviceSynchronize();
    printf("mismatching patterns (exponent < 31): %u\n", *bad);
    for (uint32_t i = 0; i < *bad && i < 16; i++) printf("  u=%04x twoscl=%08x kernel=%08x ieee=%08x\n", first[i*4], first[i*4+1], first[i*4+2], first[i*4+3]);
    return 0;
}
