/*
 * shim_check.cu -- the candidate re-check kernel (h264_wave.h, wave_mb_check) built from the same
 * macroblock code as shim_cuda.cu but with ONE warp per macroblock (MB_WARPS 1: the tasks of a
 * macroblock run one after the other, exactly like the host emulation).  A re-check is almost
 * entirely the candidate stage, which is single-warp work anyway; with one warp per CTA three times
 * as many re-checks are resident per SM.
 */
#define MB_WARPS 1
#include <cuda_runtime.h>
#include <stdint.h>
#include "h264_common.h"
#include "h264_pixel.h"
#include "h264_mbenc.h"
#include "h264_wave.h"

__global__ void __launch_bounds__(32) k_check1(const FrameParams *fps, int njobs, int pass)
{
    __shared__ MBWork work;
    __shared__ FrameParams sfp;
    for (int i = threadIdx.x; i < (int)(sizeof(FrameParams) / 4); i += blockDim.x)
        ((uint32_t *)&sfp)[i] = ((const uint32_t *)(fps + blockIdx.y))[i];
    if (threadIdx.x == 0) { work.pf_inp_tag = 0; work.pf_win_tag = 0; }
    __syncwarp();
    const FrameParams *fp = &sfp;
    if (fp->fsync[FS_STATE] != pass) return;
    const int nmbx = fp->nmbx, nmb = fp->nmbx * fp->nmby;
    for (int n = blockIdx.x; n < nmb; n += gridDim.x)
    {
        int y = n / nmbx;
        wave_mb_check(fp, &work, n - y * nmbx, y, pass);
    }
}

void h264b200_launch_check1(const FrameParams *fps, int njobs, int pass, cudaStream_t st)
{
    k_check1<<<dim3(148 * 4, njobs), 32, 0, st>>>(fps, njobs, pass);
}
