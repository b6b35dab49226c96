// tools/ubench/intpeak.cu -- measured integer / packed-byte issue peaks of the SM ("INT32 peak must be
// micro-benchmarked on the box").  Every kernel keeps 8 independent dependency chains per thread so that the pipes,
// not latencies, are the limit; 148 x 8 CTAs of 256 threads.  Prints G-ops/s for the whole chip:
//   imad      32-bit multiply-add                      (1 op  = 1 IMAD)
//   vsadu4    packed-byte sum of absolute differences  (1 op  = 1 VABSDIFF4 with accumulate = 4 samples)
//   dp4a      packed-byte dot product                  (1 op  = 1 IDP4A = 4 MACs)
//   vavgu4    packed-byte rounded average              (lowered by the compiler to a LOP3/IADD sequence)
//   shfl      warp shuffle                             (1 op = 1 SHFL of a warp)
// Build and run:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/intpeak tools/ubench/intpeak.cu && /tmp/intpeak
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define INNER 64
template <int OP> __global__ void __launch_bounds__(256) k(unsigned *out, unsigned a, unsigned b, int iters)
{
    unsigned x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x * 2654435761u + c;
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int j = 0; j < INNER; j++)
#pragma unroll
            for (int c = 0; c < CHAINS; c++)
            {
                if (OP == 0) x[c] = x[c] * a + b;
                else if (OP == 2) x[c] = __vsadu4(x[c] ^ (unsigned)j, a) + x[c];
                else if (OP == 3) x[c] = (unsigned)__dp4a((int)(x[c] ^ (unsigned)j), (int)a, (int)x[c]);
                else if (OP == 4) x[c] = __vavgu4(x[c], a ^ (unsigned)j);
                else x[c] = __shfl_xor_sync(0xffffffffu, x[c], 1 + (j & 15));
            }
    }
    unsigned s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s ^= x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> static void run(const char *name, unsigned *d, double per_op_samples)
{
    const int grid = 148 * 8, iters = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<grid, 256>>>(d, 3u, 0x01010101u, 4);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++)
    {
        cudaEventRecord(e0);
        k<OP><<<grid, 256>>>(d, 3u, 0x01010101u, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double ops = (double)grid * 256 * iters * INNER * CHAINS;
    printf("{\"op\": \"%s\", \"ms\": %.4f, \"gops\": %.1f, \"gsamples\": %.1f}\n", name, best, ops / best * 1e-6, ops * per_op_samples / best * 1e-6);
}

int main()
{
    unsigned *d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("imad", d, 1);
    run<2>("vsadu4", d, 4);
    run<3>("dp4a", d, 4);
    run<4>("vavgu4", d, 4);
    run<5>("shfl", d, 1);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
