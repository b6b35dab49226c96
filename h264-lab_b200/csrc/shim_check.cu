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

#ifndef CHECK_MIN_BLOCKS
#define CHECK_MIN_BLOCKS 24
#endif
__global__ void __launch_bounds__(32, CHECK_MIN_BLOCKS) k_check1(const FrameParams *fps, int njobs, int pass)
{
    __shared__ MBWork work;
    __shared__ FrameParams sfp;
    for (int i = threadIdx.x; i < (int)(sizeof(FrameParams) / 4); i += blockDim.x)
        ((uint32_t *)&sfp)[i] = ((const uint32_t *)(fps + blockIdx.y))[i];
    if (threadIdx.x == 0) { work.pf_inp_tag = 0; work.pf_win_tag = 0; work.map_tag[0] = work.map_tag[1] = 0; work.pf_enable = 0; }
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

/* Speculative motion estimation ahead of the wavefront (h264_wave.h, me_prepass_mb): one warp per macroblock, every
 * macroblock of every P frame of the submission, no dependencies between them.  round 0 predicts the context from the
 * previous frame's motion field, later rounds from the field the round before predicted. */
#ifndef ME_MIN_BLOCKS
#define ME_MIN_BLOCKS 20
#endif
__global__ void __launch_bounds__(32, ME_MIN_BLOCKS) k_me(const FrameParams *fps, int njobs, int round)
{
    __shared__ MBWork work;
    __shared__ FrameParams sfp;
    for (int i = threadIdx.x; i < (int)(sizeof(FrameParams) / 4); i += blockDim.x)
        ((uint32_t *)&sfp)[i] = ((const uint32_t *)(fps + blockIdx.y))[i];
    if (threadIdx.x == 0) { work.pf_inp_tag = 0; work.pf_win_tag = 0; work.map_tag[0] = work.map_tag[1] = 0; work.pf_enable = 0; }
    __syncwarp();
    const FrameParams *fp = &sfp;
    if (!fp->use_me || fp->slice_type != SLICE_P) return;
    const int nmbx = fp->nmbx, nmb = fp->nmbx * fp->nmby;
    if (round == 0)
    {
        for (int n = blockIdx.x; n < nmb; n += gridDim.x)
        {
            int y = n / nmbx;
            me_prepass_mb(fp, &work, n - y * nmbx, y, 0);
        }
        return;
    }
    /* refinement rounds: only the macroblocks k_me_scan listed */
    const int cnt = fp->me_count[round];
    for (int i = blockIdx.x; i < cnt; i += gridDim.x)
    {
        const int n = fp->me_list[i], y = n / nmbx;
        me_prepass_mb(fp, &work, n - y * nmbx, y, round);
    }
}

/* one THREAD per macroblock: list the macroblocks whose record has to be recomputed in refinement round `round` */
__global__ void __launch_bounds__(256) k_me_scan(const FrameParams *fps, int njobs, int round)
{
    const FrameParams *fp = fps + blockIdx.y;
    if (!fp->use_me || fp->slice_type != SLICE_P) return;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= fp->nmbx * fp->nmby) return;
    if (me_record_stale(fp, n)) fp->me_list[atomicAdd(fp->me_count + round, 1)] = n;
}

/* Intra modes of every macroblock of a finished sweep 0 (h264_wave.h, wave_mb_intra_check): one warp per macroblock,
 * independent of each other; winners are queued for repair pass 1. */
__global__ void __launch_bounds__(32) k_intra_check(const FrameParams *fps, int njobs)
{
    __shared__ MBWork work;
    __shared__ FrameParams sfp;
    for (int i = threadIdx.x; i < (int)(sizeof(FrameParams) / 4); i += blockDim.x)
        ((uint32_t *)&sfp)[i] = ((const uint32_t *)(fps + blockIdx.y))[i];
    if (threadIdx.x == 0) { work.pf_inp_tag = 0; work.pf_win_tag = 0; work.map_tag[0] = work.map_tag[1] = 0; work.pf_enable = 0; }
    __syncwarp();
    const FrameParams *fp = &sfp;
    if (!fp->spec_no_intra || fp->slice_type != SLICE_P || fp->fsync[FS_STATE] != 1) return;
    const int nmbx = fp->nmbx, nmb = fp->nmbx * fp->nmby;
    for (int n = blockIdx.x; n < nmb; n += gridDim.x)
    {
        int y = n / nmbx;
        wave_mb_intra_check(fp, &work, n - y * nmbx, y);
    }
}

void h264b200_launch_intra_check(const FrameParams *fps, int njobs, int max_nmb, cudaStream_t st)
{
    k_intra_check<<<dim3((max_nmb + 3) / 4, njobs), 32, 0, st>>>(fps, njobs);
}

void h264b200_launch_me(const FrameParams *fps, int njobs, int max_nmb, int round, cudaStream_t st)
{
    if (round == 0) { k_me<<<dim3((max_nmb + 3) / 4, njobs), 32, 0, st>>>(fps, njobs, 0); return; }
    k_me_scan<<<dim3((max_nmb + 255) / 256, njobs), 256, 0, st>>>(fps, njobs, round);
    k_me<<<dim3((max_nmb + 15) / 16, njobs), 32, 0, st>>>(fps, njobs, round);
}

void h264b200_launch_check1(const FrameParams *fps, int njobs, int pass, cudaStream_t st)
{
    k_check1<<<dim3(148 * 8, njobs), 32, 0, st>>>(fps, njobs, pass);
}
