// tools/fetch_granularity_probe.cu -- how many bytes does one 4-byte global load pull from DRAM?
// (development aid; `nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe tools/fetch_granularity_probe.cu`,
//  then `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./probe`)
//
// Every thread reads 4 bytes at a 512-byte stride over a 4 GiB buffer (8.4 M requests).  Measured on
// B200: ld.global.nc / .cg / .cv / .cs / L1::no_allocate all read 1.07 GB = 128 bytes per request
// (188 us); ld.global.nc.L2::64B reads 0.54 GB = 64 bytes per request (104 us).  At a 128-byte
// stride the plain load reads the whole buffer.  The direct-gather warp kernel therefore loads with
// the .L2::64B form (bev_b200/csrc/warp_u8c3.cuh: ldg_sparse).
#include <cstdio>
#include <cuda_runtime.h>

template <int V> __device__ __forceinline__ unsigned ld(const unsigned *p)
{
    unsigned v;
    if (V == 0) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (V == 1) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (V == 2) asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    if (V == 3) asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <int V, int STRIDE> __global__ void probe(const unsigned *__restrict__ p, size_t n, unsigned *out)
{
    unsigned acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc += ld<V>(p + i * (STRIDE / 4));
    if (acc == 0x12345678u) *out = acc;
}
int main()
{
    const size_t bytes = 4ull << 30;
    unsigned *p, *o;
    cudaMalloc(&p, bytes);
    cudaMalloc(&o, 4);
    cudaMemset(p, 1, bytes);
    probe<0, 128><<<148 * 16, 256>>>(p, bytes / 128, o);
    probe<0, 512><<<148 * 16, 256>>>(p, bytes / 512, o);
    probe<1, 512><<<148 * 16, 256>>>(p, bytes / 512, o);
    probe<2, 512><<<148 * 16, 256>>>(p, bytes / 512, o);
    probe<3, 512><<<148 * 16, 256>>>(p, bytes / 512, o);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
