/*
 * shim_cuda.cu -- sm_100a kernels + the C-ABI shim (include/h264b200_shim.h).
 *
 * One submission encodes one frame for each of n independent encoder instances
 * (closed-GOP segments or streams).  Per submission, on the submitting thread's lane
 * (a CUDA stream pair; see `Lane`):
 *
 *   H2D inputs [-> k_denoise     : optional temporal noise suppressor in front of the encoder]
 *              -> k_sadmap       : P frames: per macroblock, quadrant SADs at 15 x 15 full-sample offsets and 5 x 5
 *                                  quarter-sample positions (h264_sadmap.h) -- what the motion search looks up
 *              -> k_me (x3, + k_me_scan; shim_check.cu): speculative motion estimation of every macroblock on a
 *                                  PREDICTED context, result stored with the context as its key (h264_wave.h)
 *              -> k_encode_rows  : sweep 0: macroblock decisions + transform/quant/recon.  P frames: decide / work
 *                                  pipeline inside the row's CTA (h264_fast.h; records staged by TMA bulk copies).
 *                                  Persistent-style wavefront: one CTA of 4 warps per macroblock
 *                                  ROW; rows are claimed from an atomic ticket so that a
 *                                  claimed row's predecessor is always running; a row may
 *                                  process macroblock x when the row above has finished
 *                                  x+2 macroblocks (acquire/release on per-row counters).
 *                                  The first n tickets follow the mv_clusters trajectory.
 *              -> k_intra_check (second stream) || k_check1: intra modes of the macroblocks the sweep decided without
 *                                  them / candidate stage of those whose speculated cluster candidates were wrong
 *              -> k_after_check, k_repair_round x3, k_encode_rows (repair wave),
 *                 k_replay       : the exact-wavefront machinery of h264_wave.h (twice, more
 *                                  passes only after a host check)
 *              -> k_deblock_rows : in-loop filter, same wavefront on its own counters, luma and
 *                                  chroma as independent wavefronts;   || on the lane's second stream:
 *                 k_cavlc        : one thread per macroblock: syntax + CAVLC bit strings,
 *                 k_scan         : per frame, exclusive prefix sum of the bit lengths,
 *                 k_pack         : scatter the strings into the slice payload.
 *              -> k_borders      : guard-band replication of the new reference picture.
 *              -> k_hpel         : its three half-sample planes.
 *   D2H payload.
 *
 * Build-time knobs (A/B-tested on B200): ENC_MIN_BLOCKS / ENC_MIN_BLOCKS_I (CTAs per SM of k_encode_rows /
 * k_encode_rows_i, i.e. their register budgets), H264_INL (h264_common.h: which big leaves are inlined), PROG_STRIDE, POLL_NS.
 *
 * The per-macroblock code is in h264_*.h (shared with the test-only host emulation).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <mutex>
#include <atomic>

#include "h264_common.h"
#include "h264_pixel.h"
#include "h264_mbenc.h"
#include "h264_wave.h"
#include "h264_fast.h"
#include "h264_cavlc.h"
#include "h264_deblock.h"
#include "h264_denoise.h"
#include "../../include/h264b200_shim.h"

#ifndef ENC_MIN_BLOCKS_I
#define ENC_MIN_BLOCKS_I 3      /* k_encode_rows_i: submissions without a P frame */
#endif
#ifndef ENC_MIN_BLOCKS
#define ENC_MIN_BLOCKS 2        /* 232 registers, no spills; with the decide / work fast path the latency of the complete path
                                   counts for more than resident row slots (3: 168 registers, -1 %; 4: 128 registers, -5 %) */
#endif
/* Row progress counters sit PROG_STRIDE ints apart: one 128-byte line (one L2 slice entry) per
 * row, so that the pollers of neighbouring rows do not queue up on the same line. */
#ifndef PROG_STRIDE
#define PROG_STRIDE 32
#endif
#ifndef POLL_NS
#define POLL_NS 20
#endif
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "h264b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -3; } } while (0)

/* ------------------------------------------------------------------------------ */
/* wavefront helpers                                                                */
/* ------------------------------------------------------------------------------ */
/* Row progress counters: the producer's samples / records are ordinary stores by all threads of
 * its CTA, made visible by ONE release (bar.sync orders them before thread 0's st.release.gpu,
 * cumulativity does the rest); the consumer's thread 0 polls with ld.acquire.gpu and the CTA
 * barrier extends the acquire to the other threads (they share the SM's L1). */
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
/* Polling uses RELAXED loads (served by L2, they leave the SM's L1 alone); one acquire fence
 * follows when the awaited value has been seen.  Polling with ld.acquire would invalidate the
 * L1 of every co-resident CTA at each iteration. */
__device__ __forceinline__ int ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acquire() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
/* After relaxed polls have seen the awaited value: ONE acquiring load of the same counter (it reads that value or a later
 * one, so it synchronises with a release that covers everything awaited).  Unlike a fence it does not wait for the
 * polling thread's own outstanding stores -- the write-back of the macroblock it has just finished. */
__device__ __forceinline__ void acquire_counter(const int *p) { (void)ld_acquire(p); }
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
/* two counters published together: ONE release fence, then two relaxed stores (each of them a release pattern with it) */
__device__ __forceinline__ void st_release2(int *p, int v, int *q, int u)
{
    asm volatile("fence.acq_rel.gpu;\n\tst.relaxed.gpu.global.s32 [%0], %1;\n\tst.relaxed.gpu.global.s32 [%2], %3;" ::"l"(p), "r"(v), "l"(q), "r"(u) : "memory");
}
__device__ __forceinline__ void wait_row(const int *progress_above, int need)      /* one warp */
{
    if (LANE_ID == 0) { while (ld_relaxed(progress_above) < need) __nanosleep(32); acquire_counter(progress_above); }
    __syncwarp();
}
__device__ __forceinline__ void publish_row(int *progress, int done)               /* one warp */
{
    __syncwarp();
    if (LANE_ID == 0) st_release(progress, done);
}
__device__ __forceinline__ void wait_row_cta(const int *progress_above, int need)  /* whole CTA */
{
    if (threadIdx.x == 0) { while (ld_relaxed(progress_above) < need) __nanosleep(POLL_NS); acquire_counter(progress_above); }
    __syncthreads();
}
__device__ __forceinline__ void publish_row_cta(int *progress, int done)
{
    __syncthreads();
    if (threadIdx.x == 0) st_release(progress, done);
}

/* sync area layout per submission: [0] ticket of k_encode_rows, [1] ticket of k_deblock_rows.
 *
 * k_encode_rows: one sweep (h264_wave.h) over the frames of the submission; one CTA of
 * MB_WARPS warps per macroblock row.  Rows are claimed from an atomic ticket, so the row a
 * CTA waits for (same frame, row - 1, an earlier ticket) is always running or finished:
 * no co-residency assumption, no deadlock.  Row progress counters are monotonic: sweep p of
 * a row counts from p*nmbx.  k_replay then replays the cluster trajectory of every frame and
 * publishes FS_DONE or the number of the repair sweep the host has to launch. */
/* The mv_clusters recurrence (clusters_update, H:5263-5278) over `avail` consecutive records held one per lane.  The
 * serial chain is kept as short as it can be: the state stays unpacked in four registers, every lane prepares x, y and
 * the norm of ITS record once, and all broadcasts are independent of the state, so they run ahead of the chain
 * (cluster -> norms -> tests -> selects).  Same arithmetic as clusters_update, including the 16-bit wrap of mv_pack. */
__device__ __forceinline__ void replay_chunk(int32_t c[2], int avail, int mv0, int flags, int u0, int u1, int &t0, int &t1, int &ndirty)
{
    const unsigned FULLM = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int mx = mv_x(mv0), my = mv_y(mv0), mn = mx * mx + my * my;
    int c0x = mv_x(c[0]), c0y = mv_y(c[0]), c1x = mv_x(c[1]), c1y = mv_y(c[1]);
#pragma unroll 8
    for (int i = 0; i < 32; i++)
    {
        if (i >= avail) break;
        const int f = __shfl_sync(FULLM, flags, i), a0 = __shfl_sync(FULLM, u0, i), a1 = __shfl_sync(FULLM, u1, i);
        const int x = __shfl_sync(FULLM, mx, i), y = __shfl_sync(FULLM, my, i), norm = __shfl_sync(FULLM, mn, i);
        const int r0 = mv_pack((c0x + 1) & ~3, (c0y + 1) & ~3), r1 = mv_pack((c1x + 1) & ~3, (c1y + 1) & ~3);
        if (lane == i) { t0 = r0; t1 = r1; }
        if ((f & SPEC_USED_CL) && (r0 != a0 || r1 != a1)) ndirty++;
        const int n0 = c0x * c0x + c0y * c0y, n1 = c1x * c1x + c1y * c1y;
        const int upd = f & SPEC_UPDATES;
        const int lo = upd && norm < n1, hi = upd && norm >= n0;
        const int b0x = (int)(int16_t)((63 * c0x + x + 32) >> 6), b0y = (int)(int16_t)((63 * c0y + y + 32) >> 6);
        const int b1x = (int)(int16_t)((63 * c1x + x + 32) >> 6), b1y = (int)(int16_t)((63 * c1y + y + 32) >> 6);
        c0x = lo ? b0x : c0x; c0y = lo ? b0y : c0y;
        c1x = hi ? b1x : c1x; c1y = hi ? b1y : c1y;
    }
    c[0] = mv_pack(c0x, c0y); c[1] = mv_pack(c1x, c1y);
}

/* Trajectory follower of sweep 0 (one warp per frame): consumes the macroblocks of the frame in
 * raster order as they finish, replays mv_clusters_update, publishes the running state for
 * the macroblocks still to start (their speculation) and, at the end, does what
 * wave_end_of_pass(pass 0) does: true candidates per macroblock, dirty count, FS_DONE or 1. */
__device__ void trajectory_follower(const FrameParams *fp)
{
    const int lane = threadIdx.x;
    const int nmbx = fp->nmbx, nmb = fp->nmbx * fp->nmby;
    volatile int32_t *fs = (volatile int32_t *)fp->fsync;
    if (fp->slice_type != SLICE_P)
    {
        if (lane == 0) { fs[FS_PASSES] = 1; __threadfence(); fs[FS_STATE] = FS_DONE; }
        return;
    }
    int32_t c[2];
    c[0] = fp->clusters[0]; c[1] = fp->clusters[1];
    int ndirty = 0, n = 0;
    /* the fast path publishes a macroblock's vectors and speculation record ("decided") before its pixel work is done */
    const int *prog = fp->use_me ? fp->row_progress_mv : fp->row_progress;
#ifdef H264_FASTPROF
    unsigned long long fgt0, fwait = 0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(fgt0));
#endif
    while (n < nmb)
    {
#ifdef H264_FASTPROF
        unsigned long long fa_, fb_;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(fa_));
#endif
        /* how many macroblocks of the raster order are finished from n on (at most 32) */
        int row = n / nmbx, x = n - row * nmbx, avail = 0;
        if (lane == 0)
        {
            int p;
            while ((p = ld_relaxed(prog + row * PROG_STRIDE)) <= x) __nanosleep(200);
            acquire_counter(prog + row * PROG_STRIDE);
            avail = min(p - x, 32 - (n & 31));       /* chunks end at multiples of 32 (checkpoints) */
        }
        avail = __shfl_sync(0xffffffffu, avail, 0);
#ifdef H264_FASTPROF
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(fb_));
        fwait += fb_ - fa_;
        if (lane == 0 && fp->prof && (n % nmbx) == 0) fp->prof[68 * 8 + 2 * (n / nmbx)] = (int)(fb_ - fgt0);
#endif
        if (!(n & 31) && lane < 2) fp->cl_ckpt[2 * (n >> 5) + lane] = c[lane];
        int mv0 = 0, flags = 0, u0 = 0, u1 = 0;
        if (lane < avail)
        {
            const MBSpec *sp = fp->spec + n + lane;
            mv0 = sp->mv0; flags = sp->flags; u0 = sp->cl_used[0]; u1 = sp->cl_used[1];
        }
        int t0 = 0, t1 = 0;
        replay_chunk(c, avail, mv0, flags, u0, u1, t0, t1, ndirty);
        if (lane < avail) { fp->cl_true[2 * (n + lane)] = t0; fp->cl_true[2 * (n + lane) + 1] = t1; }
        n += avail;
        if (lane == 0) { fs[FS_LIVE] = c[0]; fs[FS_LIVE + 1] = c[1]; fs[FS_LIVE + 2] = n; }
    }
    __syncwarp();
#ifdef H264_FASTPROF
    if (lane == 0 && fp->prof) { unsigned long long fe_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(fe_)); fp->prof[68 * 8 + 200] = (int)(fe_ - fgt0); fp->prof[68 * 8 + 201] = (int)fwait; }
#endif
    if (lane == 0)
    {
        fs[FS_CL_END] = c[0]; fs[FS_CL_END + 1] = c[1];
        fs[FS_NDIRTY] = ndirty;
        /* a frame whose sweep 0 left the intra modes out is never finished here: k_intra_check may still queue re-encodes,
         * k_after_check(1) finishes the frame when neither it nor the candidate re-check found anything */
        const int more = ndirty || fp->spec_no_intra;
        if (!more) { fp->clusters[0] = c[0]; fp->clusters[1] = c[1]; fs[FS_PASSES] = 1; }
        __threadfence();
        fs[FS_STATE] = more ? 1 : FS_DONE;
    }
}

/* Trajectory follower of a repair wave (pass > 0, one warp per frame): what k_replay would do after the
 * wave -- replay the trajectory from the first macroblock whose cluster-relevant result changed -- done
 * WHILE the wave runs, gated by the row counters like the follower of sweep 0.  Every change of this
 * pass is a causal successor of a change made by the parallel rounds that ran before the wave, so the
 * resume point is known when the wave starts; k_replay re-checks that and falls back to its own
 * replay otherwise (FS_REPLAYED / FS_REPLAY_TF). */
__device__ void repair_follower(const FrameParams *fp, int pass)
{
    const int lane = threadIdx.x;
    volatile int32_t *fs = (volatile int32_t *)fp->fsync;
    if (fp->slice_type != SLICE_P || fs[FS_STATE] != pass || fs[FS_TRAJ_CHANGED] == 0) return;
    const int nmbx = fp->nmbx, nmb = fp->nmbx * fp->nmby, base = pass * nmbx;
    const int tf = fs[FS_TRAJ_FIRST];
    const int first_block = tf > 0 ? (0x3fffffff - tf) >> 5 : 0;
    int32_t c[2];
    c[0] = fp->clusters[0]; c[1] = fp->clusters[1];
    if (first_block > 0) { c[0] = fp->cl_ckpt[2 * first_block]; c[1] = fp->cl_ckpt[2 * first_block + 1]; }
    int ndirty = 0, n = 32 * first_block;
    while (n < nmb)
    {
        int row = n / nmbx, x = n - row * nmbx, avail = 0;
        if (lane == 0)
        {
            int p;
            while ((p = ld_relaxed(fp->row_progress + row * PROG_STRIDE) - base) <= x) __nanosleep(100);
            acquire_counter(fp->row_progress + row * PROG_STRIDE);
            avail = min(p - x, 32 - (n & 31));
        }
        avail = __shfl_sync(0xffffffffu, avail, 0);
        if (!(n & 31) && lane < 2) fp->cl_ckpt[2 * (n >> 5) + lane] = c[lane];
        int mv0 = 0, flags = 0, u0 = 0, u1 = 0;
        if (lane < avail)
        {
            const MBSpec *sp = fp->spec + n + lane;
            mv0 = sp->mv0; flags = sp->flags; u0 = sp->cl_used[0]; u1 = sp->cl_used[1];
        }
        int t0 = 0, t1 = 0;
        replay_chunk(c, avail, mv0, flags, u0, u1, t0, t1, ndirty);
        if (lane < avail) { fp->cl_true[2 * (n + lane)] = t0; fp->cl_true[2 * (n + lane) + 1] = t1; }
        n += avail;
    }
    __syncwarp();
    if (lane == 0)
    {
        fs[FS_CL_END] = c[0]; fs[FS_CL_END + 1] = c[1];
        fs[FS_NDIRTY] = ndirty;
        fs[FS_REPLAY_TF] = tf;
        __threadfence();
        fs[FS_REPLAYED] = pass;
    }
}

__device__ int g_d_no_fast = 0;      /* developer knob H264B200_NO_FAST (set from the host): no decide / work fast path */
/* one thread: bulk copies (TMA engine) of the motion-estimation records of macroblocks x0 .. x0 + FAST_BATCH - 1 of row y into
 * work.me_stage[], all completing on one barrier phase */
__device__ __forceinline__ void me_stage_issue(const FrameParams *fp, MBWork *w, int x0, int y)
{
    const int num = min(FAST_BATCH, fp->nmbx - x0);
    w->me_first = x0; w->me_num = num > 0 ? num : 0;
    if (num <= 0) return;
    const unsigned bar = smem_u32(&w->me_bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(num * ME_WORDS * 4)) : "memory");
    for (int k = 0; k < num; k++)
    {
        const uint32_t *src = fp->sadmap + (size_t)(y * fp->nmbx + x0 + k) * SM_WORDS + SM_ME_OFF;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(w->me_stage[k])), "l"(src), "r"((unsigned)(ME_WORDS * 4)), "r"(bar) : "memory");
    }
    w->me_cnt++;
}

static __device__ __forceinline__ void encode_rows_body(const FrameParams *fps, int njobs, int *tickets, int pass)
{
    __shared__ MBWork work;
    __shared__ FrameParams sfp;
    __shared__ int s_item;
    if (threadIdx.x == 0)
    {
        s_item = atomicAdd(&tickets[0], 1); work.scal[9] = 0; work.pf_inp_tag = 0; work.pf_win_tag = 0;
        /* SAD-map staging (bulk copies into work.maps[], h264_mbenc.h map_prefetch): only the row loop of sweep 0 stages ahead */
        work.map_tag[0] = work.map_tag[1] = 0; work.map_cnt[0] = work.map_cnt[1] = 0; work.pf_enable = pass == 0;
        work.me_cnt = 0; work.me_first = -1; work.me_num = 0; work.batch_n = 0;
        mbar_init(&work.map_bar[0], 1); mbar_init(&work.map_bar[1], 1); mbar_init(&work.me_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int item = s_item;
    {   /* the first njobs tickets are the trajectory followers */
        if (item < njobs)
        {
            if (threadIdx.x < 32) { if (pass == 0) trajectory_follower(fps + item); else repair_follower(fps + item, pass); }
            return;
        }
        item -= njobs;
    }
    const int job = item % njobs, row = item / njobs;
    for (int i = threadIdx.x; i < (int)(sizeof(FrameParams) / 4); i += blockDim.x)
        ((uint32_t *)&sfp)[i] = ((const uint32_t *)(fps + job))[i];
    __syncthreads();
    const FrameParams *fp = &sfp;
    if (row >= fp->nmby) return;
    if (pass > 0 && fp->fsync[FS_STATE] != pass) return;      /* frame already exact (or failed) */
    const int nmbx = fp->nmbx;
    int *progress = fp->row_progress;
    const int base = pass * nmbx;
    if (pass > 0 && fp->fsync[FS_WAVE_TAGS] == 0)
    {   /* nothing was tagged for this wave (the last parallel round changed nothing): no macroblock can change in it,
         * every row says so at once -- no chain of round trips down the frame -- and leaves */
        if (threadIdx.x == 0) { st_release(fp->row_clean + row, pass); st_release(progress + row * PROG_STRIDE, base + nmbx); }
        return;
    }
    if (pass > 0)
    {
        /* Repair sweeps touch few macroblocks.  A row in which nothing is queued, below a row that
         * went through the sweep without anything to do, has nothing to do either: it says so
         * and leaves, so that "nothing to do" travels down the frame at one global round trip
         * per row instead of one per macroblock. */
        int has_need = 0;
        for (int i = threadIdx.x; i < nmbx; i += blockDim.x) has_need |= fp->need_reenc[row * nmbx + i] == REPAIR_TAG(pass, REPAIR_ROUNDS);
        has_need = __syncthreads_or(has_need);
        if (!has_need)
        {
            __shared__ int s_clean;
            if (threadIdx.x == 0)
            {
                int clean = 1;
                if (row > 0)
                    for (;;)
                    {
                        if (ld_relaxed(fp->row_clean + row - 1) == pass) break;
                        if (ld_relaxed(progress + (row - 1) * PROG_STRIDE) >= base + 1) { clean = ld_relaxed(fp->row_clean + row - 1) == pass; break; }
                        __nanosleep(20);
                    }
                fence_acquire();
                s_clean = clean;
                if (clean) { st_release(fp->row_clean + row, pass); st_release(progress + row * PROG_STRIDE, base + nmbx); }
            }
            __syncthreads();
            if (s_clean) return;
        }
    }
    if (pass == 0 && fp->slice_type == SLICE_P && fp->use_me && fp->spec_no_intra && !g_d_no_fast)
    {
        /* Fast path of P frames (h264_fast.h) as a pipeline inside the CTA.  Warp 0 DECIDES macroblock after macroblock from
         * the motion-estimation records (staged by bulk copies), publishing "decided" progress for the row below and
         * handing every decision to the ring work.fd[]; warps 1..3 WORK the ring off, one macroblock each (decision q goes
         * to warp 1 + q % 3).  The wavefront's serial chain is the decide step alone.  A macroblock that cannot be decided
         * from its record is encoded by the whole CTA (wave_mb_first) once the ring has drained. */
        int *prog_mv = fp->row_progress_mv;
        const int pw = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
        volatile int32_t *v_qdec = &work.q_dec, *v_cmd = &work.cmd;
        volatile int32_t *v_done = work.q_done;
        if (threadIdx.x == 0) { work.pf_enable = 0; work.q_dec = 0; work.cmd = 0; work.slow_x = 0; for (int k = 0; k < MB_WARPS; k++) work.q_done[k] = 0; }
        __syncthreads();
        /* state of the deciding warp */
        int x = 0, q = 0, seen_mv = 0, x_base = 0, q_base = 0, pub_full = 0, p_early = -1;
        int my = pw - 1;                                   /* next decision of a working warp */
        const int row_thr = fp->have_cost_stat ? fp->cost_stat[2 + row] : 0;
#ifdef H264_FASTPROF
        unsigned long long gt0, gt_poll = 0, gt_ring = 0, gt_dec = 0, gt_slow = 0, gt_pub = 0, gt_stage = 0, gt_tmp;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0));
#define GT(v) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v))
#else
#define GT(v) do { } while (0)
#endif
        for (;;)
        {
            if (pw == 0)
            {
                /* decisions worked off completely: everything below the smallest decision still open at one of the three warps */
#define FAST_PREFIX() min(q, min(min(0 + 3 * v_done[1], 1 + 3 * v_done[2]), 2 + 3 * v_done[3]))
/* the working warps make their stores visible device-wide (__threadfence) BEFORE they count a decision as done; the
 * deciding warp only has to order its reads of those counters before its own release (a CTA-scope fence: both sides
 * are in this CTA), the release store is cumulative */
#define FAST_PUBLISH_FULL_IF(k_) do { const int xf_ = x_base + (FAST_PREFIX() - q_base); if (xf_ >= pub_full + (k_)) { pub_full = xf_; __threadfence_block(); if (lane == 0) st_release(progress + row * PROG_STRIDE, xf_); } } while (0)
#define FAST_PUBLISH_FULL() FAST_PUBLISH_FULL_IF(1)
                int go_slow = 0;
                while (x < nmbx && !go_slow)
                {
                    if (x < work.me_first || x >= work.me_first + work.me_num) { if (lane == 0) me_stage_issue(fp, &work, x, row); __syncwarp(); }
                    const int first = work.me_first, num = work.me_num;
#ifdef H264_FASTPROF
                    { unsigned long long a_, b_; GT(a_);
#endif
                    mbar_wait(&work.me_bar, (unsigned)(work.me_cnt - 1) & 1u);
#ifdef H264_FASTPROF
                    GT(b_); gt_stage += b_ - a_; gt_ring += 0; }
#endif
                    for (int nb = x - first; nb < num; nb++)
                    {
                        const int need = min(x + 2, nmbx);
                        if (row > 0 && seen_mv < need)
                        {
                            int p = 0;
#ifdef H264_FASTPROF
                            unsigned long long a_, b_; GT(a_);
#endif
                            if (lane == 0) p = p_early >= 0 ? p_early : ld_relaxed(prog_mv + (row - 1) * PROG_STRIDE);
                            p = __shfl_sync(0xffffffffu, p, 0);
                            while (p < need)
                            {
                                FAST_PUBLISH_FULL();
                                __nanosleep(POLL_NS);
                                if (lane == 0) p = ld_relaxed(prog_mv + (row - 1) * PROG_STRIDE);
                                p = __shfl_sync(0xffffffffu, p, 0);
                            }
                            if (lane == 0) acquire_counter(prog_mv + (row - 1) * PROG_STRIDE);
                            __syncwarp();
                            seen_mv = p;
#ifdef H264_FASTPROF
                            GT(b_); gt_poll += b_ - a_;
#endif
                        }
#ifdef H264_FASTPROF
                        { unsigned long long a_, b_; GT(a_);
#endif
                        while (q - FAST_PREFIX() >= FAST_RING - 1) { FAST_PUBLISH_FULL(); __nanosleep(32); }
#ifdef H264_FASTPROF
                        GT(b_); gt_ring += b_ - a_; }
#endif
#ifdef H264_FASTPROF
                        unsigned long long d0_, d1_, d2_; GT(d0_);
#endif
                        if (!fast_decide(fp, &work, x, row, work.me_stage[x - first], q % FAST_RING, row_thr)) { go_slow = 1; break; }
#ifdef H264_FASTPROF
                        GT(d1_); gt_dec += d1_ - d0_;
#endif
                        q++;
                        __threadfence_block();
                        if (lane == 0) *v_qdec = q;
                        x++;
                        /* the progress of the row above that the NEXT macroblock needs is asked for before this one is
                         * published: the release waits for this warp's stores, the load's round trip runs beside it */
                        p_early = -1;
                        if (row > 0 && seen_mv < min(x + 2, nmbx) && lane == 0) p_early = ld_relaxed(prog_mv + (row - 1) * PROG_STRIDE);
                        publish_row(prog_mv + row * PROG_STRIDE, x);
                        /* complete progress (what slow macroblocks and intra neighbours of the row below wait for) costs a
                         * release of its own: while deciding, only every second step; every wait of this warp publishes at once */
                        FAST_PUBLISH_FULL_IF(2);
#ifdef H264_FASTPROF
                        GT(d2_); gt_pub += d2_ - d1_;
#endif
                    }
                }
                /* drain the ring: the complete path (or the end of the row) needs every earlier macroblock finished */
                while (FAST_PREFIX() < q) { FAST_PUBLISH_FULL(); __nanosleep(32); }
                FAST_PUBLISH_FULL();
                if (lane == 0) { work.slow_x = x; __threadfence_block(); *v_cmd = go_slow ? 1 : 2; }
            } else
            {
                const int k = pw - 1;
                int cnt = v_done[pw];
                for (;;)
                {
                    int have;
                    while (!(have = (*v_qdec > my)) && *v_cmd == 0) __nanosleep(32);
                    if (!have) have = *v_qdec > my;          /* a decision made just before the command was raised */
                    if (!have) break;
                    __threadfence_block();
                    fast_work(fp, &work, work.fd[my % FAST_RING].x, row, my % FAST_RING, pw);
                    __threadfence();                        /* reconstruction and records before the completion is seen */
                    cnt++;
                    if (lane == 0) v_done[pw] = cnt;
                    my += 3;
                    (void)k;
                }
            }
            __syncthreads();
            const int cmd = work.cmd, sx = work.slow_x;
            if (cmd == 2) break;
#ifdef H264_FASTPROF
            GT(gt_tmp);
#endif
            /* complete path for macroblock sx, whole CTA */
            if (row > 0) wait_row_cta(progress + (row - 1) * PROG_STRIDE, min(sx + 2, nmbx));
            wave_mb_first(fp, &work, sx, row);
            if (threadIdx.x < 16) work.last_mv[threadIdx.x] = fp->mbi[row * nmbx + sx].mv[threadIdx.x];
            __syncthreads();
            if (threadIdx.x == 0)
            {
                st_release2(prog_mv + row * PROG_STRIDE, sx + 1, progress + row * PROG_STRIDE, sx + 1);
                work.cmd = 0;
                atomicAdd(fp->fsync + FS_SLOW, 1);
            }
            mb_store_coefs(fp, &work);
            if (pw == 0) { x = sx + 1; x_base = x; q_base = q; pub_full = x; }
#ifdef H264_FASTPROF
            { unsigned long long b_; GT(b_); gt_slow += b_ - gt_tmp; }
#endif
        }
#ifdef H264_FASTPROF
        if (threadIdx.x == 0 && fp->prof)
        {
            unsigned long long b_; GT(b_);
            int *pr = fp->prof + row * 8;
            pr[0] = (int)(gt0 & 0x7fffffff); pr[1] = (int)(b_ - gt0); pr[2] = (int)gt_poll; pr[3] = (int)gt_ring; pr[4] = (int)gt_slow; pr[5] = q; pr[6] = (int)gt_dec; pr[7] = (int)gt_pub; pr[3] = (int)gt_ring + 0; pr[2] = (int)gt_poll; if (fp->prof) fp->prof[68 * 8 + 300 + row] = (int)gt_stage;
        }
#endif
        if (threadIdx.x == 0) { st_release(progress + row * PROG_STRIDE, nmbx); atomicAdd(fp->fsync + FS_FAST, q); }
        return;
    }
    if (pass == 0)
    {
        for (int x = 0; x < nmbx; x++)
        {
            if (row > 0) wait_row_cta(progress + (row - 1) * PROG_STRIDE, base + min(x + 2, nmbx));
            wave_mb_first(fp, &work, x, row);
            __syncthreads();
            if (threadIdx.x == 0) st_release2(fp->row_progress_mv + row * PROG_STRIDE, x + 1, progress + row * PROG_STRIDE, base + x + 1);
            mb_store_coefs(fp, &work);
        }
        return;
    }
    /* Repair sweep: as far as the row above has got, the macroblocks of this row are examined
     * 128 at a time (one thread each: queued for a re-encode, or a causal neighbour changed in
     * this sweep?); a stretch with nothing to do costs one step, the sweep only serialises at
     * the macroblocks that are actually re-encoded. */
    __shared__ int s_above, s_first;
    int x = 0;
    while (x < nmbx)
    {
        if (threadIdx.x == 0)
        {
            int p = nmbx;
            if (row > 0)
            {
                const int need = base + min(x + 2, nmbx);
                while ((p = ld_relaxed(progress + (row - 1) * PROG_STRIDE)) < need) __nanosleep(20);
                acquire_counter(progress + (row - 1) * PROG_STRIDE);
                p -= base;
            }
            s_above = p;
            s_first = 0x7fffffff;
        }
        __syncthreads();
        const int xe = s_above >= nmbx ? nmbx - 1 : s_above - 2;       /* last macroblock whose top-right neighbour is final */
        const int cnt = min(xe - x + 1, (int)blockDim.x);
        if ((int)threadIdx.x < cnt)
        {
            const int xx = x + threadIdx.x, n = row * nmbx + xx;
            const int has_l = xx > 0, has_t = row > 0, has_tl = row > 0 && xx > 0, has_tr = row > 0 && xx < nmbx - 1;
            const int f0 = fp->need_reenc[n];
            const int f1 = fp->changed_pass[has_l ? n - 1 : n];
            const int f2 = fp->changed_pass[has_t ? n - nmbx : n];
            const int f3 = fp->changed_pass[has_tl ? n - nmbx - 1 : n];
            const int f4 = fp->changed_pass[has_tr ? n - nmbx + 1 : n];
            if ((f0 == REPAIR_TAG(pass, REPAIR_ROUNDS)) | (has_l & (f1 == pass)) | (has_t & (f2 == pass)) | (has_tl & (f3 == pass)) | (has_tr & (f4 == pass)))
                atomicMin(&s_first, (int)threadIdx.x);
        }
        __syncthreads();
        const int first = s_first;
        if (first == 0x7fffffff)
        {
            x += cnt;
            publish_row_cta(progress + row * PROG_STRIDE, base + x);
            continue;
        }
        x += first;
        wave_mb_repair(fp, &work, x, row, pass);
        if (threadIdx.x == 0) atomicAdd(fp->fsync + FS_WAVE_REENC, 1);
        x++;
        publish_row_cta(progress + row * PROG_STRIDE, base + x);
        mb_store_coefs(fp, &work);
    }
}

/* Two register budgets of the same row loop.  P frames: two CTAs per SM (232 registers, no spills) -- with the decide /
 * work fast path the latency of the complete path counts for more than resident row slots (3 per SM: 168 registers,
 * -1 %; 4: 128 registers, -5 %).  Submissions of I frames only (every macroblock takes the complete path, nothing
 * waits on a vector chain): three per SM, +13 % on the all-intra configuration. */
__global__ void __launch_bounds__(MB_WARPS * 32, ENC_MIN_BLOCKS) k_encode_rows(const FrameParams *fps, int njobs, int *tickets, int pass)
{
    encode_rows_body(fps, njobs, tickets, pass);
}
__global__ void __launch_bounds__(MB_WARPS * 32, ENC_MIN_BLOCKS_I) k_encode_rows_i(const FrameParams *fps, int njobs, int *tickets, int pass)
{
    encode_rows_body(fps, njobs, tickets, pass);
}

/* parallel repair round r of pass `pass` (h264_wave.h): every CTA looks at a strip of macroblocks
 * (one flag per thread), then re-encodes the ones tagged for this round */
__global__ void __launch_bounds__(MB_WARPS * 32, ENC_MIN_BLOCKS) k_repair_round(const FrameParams *fps, int njobs, int pass, int r)
{
    __shared__ MBWork work;
    __shared__ FrameParams sfp;
    __shared__ int s_list[MB_WARPS * 32], s_cnt;
    for (int i = threadIdx.x; i < (int)(sizeof(FrameParams) / 4); i += blockDim.x)
        ((uint32_t *)&sfp)[i] = ((const uint32_t *)(fps + blockIdx.y))[i];
    if (threadIdx.x == 0) { work.scal[9] = 0; work.pf_inp_tag = 0; work.pf_win_tag = 0; work.map_tag[0] = work.map_tag[1] = 0; work.pf_enable = 0; }
    __syncthreads();
    const FrameParams *fp = &sfp;
    if (fp->fsync[FS_STATE] != pass) return;
    const int nmbx = fp->nmbx, nmb = fp->nmbx * fp->nmby, tag = REPAIR_TAG(pass, r);
    for (int base = 0; base < nmb; base += gridDim.x * blockDim.x)
    {
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        const int n = base + threadIdx.x * gridDim.x + blockIdx.x;
        if (n < nmb && fp->need_reenc[n] == tag) s_list[atomicAdd(&s_cnt, 1)] = n;
        __syncthreads();
        const int cnt = s_cnt;
        for (int i = 0; i < cnt; i++)
        {
            const int m = s_list[i], y = m / nmbx;
            wave_mb_round(fp, &work, m - y * nmbx, y, pass, r);
            mb_store_coefs(fp, &work);
        }
        __syncthreads();
    }
}

__global__ void k_after_check(const FrameParams *fps, int njobs, int pass)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < njobs) wave_after_check(fps + j, pass);
}

/* one warp per frame: end-of-sweep bookkeeping (trajectory replay, convergence test) */
__global__ void __launch_bounds__(32) k_replay(const FrameParams *fps, int njobs, int pass)
{
    __shared__ MBWork work;
    const FrameParams *fp = fps + blockIdx.x;
    if (fp->fsync[FS_STATE] != pass) return;
    int next = wave_end_of_pass(fp, &work, pass);
    if (next != FS_DONE && next > fp->max_passes) { if (threadIdx.x == 0) atomicOr(&fp->out_info[1], 4); next = FS_DONE; }
    if (threadIdx.x == 0) fp->fsync[FS_STATE] = next;
}

/* In-loop filter, x+2y wavefront, one warp per macroblock row.  The first njobs tickets do
 * something unrelated that only has to happen once the frame is exact: they replay the cluster
 * update over this frame's final motion field from its END state, i.e. the trajectory
 * speculated for the NEXT frame (h264_wave.h "predict"). */
__global__ void __launch_bounds__(32) k_deblock_rows(const FrameParams *fps, int njobs, int *tickets)
{
    __shared__ int s_item;
    __shared__ DeblockTile tile[2];
    if (threadIdx.x == 0) s_item = atomicAdd(&tickets[1], 1);
    __syncwarp();
    int item = s_item;
    if (item < njobs)
    {
        const FrameParams *fp = fps + item;
        if (fp->fsync[FS_STATE] == FS_DONE && fp->slice_type == SLICE_P) wave_replay(fp, (MBWork *)0, 1);
        return;
    }
    item -= njobs;
    const int part = item & 1;                 /* 0: luma wavefront, 1: chroma wavefront (independent) */
    item >>= 1;
    const int job = item % njobs, row = item / njobs;
    const FrameParams *fp = fps + job;
    if (row >= fp->nmby || fp->disable_deblock || fp->fsync[FS_STATE] != FS_DONE) return;
    const int nmbx = fp->nmbx;
    int *progress = part ? fp->row_progress_dfc : fp->row_progress_df;
    /* software pipeline over the macroblocks of the row: tile[x & 1] is filtered while tile[(x + 1) & 1] fills */
    deblock_mb(fp, &tile[0], 0, row, part, 0);
    int seen = 0;                              /* progress of the row above as far as this warp has acquired it */
    for (int x = 0; x < nmbx; x++)
    {
        DeblockTile *cur = &tile[x & 1], *nxt = &tile[(x + 1) & 1];
        /* the progress of the row above is asked for BEFORE the next tile's prefetch and looked at after it: in the steady
         * state of the wavefront the row above is exactly far enough, and the round trip hides behind the prefetch */
        const int need = min(x + 2, nmbx), must = row > 0 && seen < need;
        int p = seen;
        if (must && LANE_ID == 0) p = ld_relaxed(progress + (row - 1) * PROG_STRIDE);
        int bs_next = 0;                       /* the next macroblock's boundary strengths: loads now, store after the filters */
        if (x + 1 < nmbx) { deblock_prefetch(fp, nxt, x + 1, row, part); bs_next = deblock_bs_item(fp, x + 1, row, LANE_ID); }
        if (must)
        {
            /* relaxed polls, then ONE acquiring load of the counter: unlike a fence it does not wait for this lane's own
             * write-back stores of the previous macroblock */
            if (LANE_ID == 0) { while (p < need) { __nanosleep(32); p = ld_relaxed(progress + (row - 1) * PROG_STRIDE); } acquire_counter(progress + (row - 1) * PROG_STRIDE); }
            seen = __shfl_sync(0xffffffffu, p, 0);
        }
        deblock_mb(fp, cur, x, row, part, 1);
        nxt->bs[LANE_ID] = (uint8_t)bs_next;
        if (x + 1 < nmbx) deblock_handover(cur, nxt, part);
        publish_row(progress + row * PROG_STRIDE, x + 1);
    }
}

__global__ void k_borders(const FrameParams *fps, int njobs)
{
    const FrameParams *fp = fps + blockIdx.y;
    if (fp->fsync[FS_STATE] != FS_DONE) return;
    for (int pl = 0; pl < 3; pl++)
    {
        long n = border_samples(fp, pl);
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
            extend_border_sample(fp, pl, i);
    }
}

/* start-of-frame state: row progress counters (encode, deblock luma / chroma, clean flags), out_info, fsync with the
 * cluster state the trajectory starts from (FS_LIVE), and the tickets of the submission */
__global__ void __launch_bounds__(256) k_frame_init(const FrameParams *fps, int njobs, int *tickets)
{
    const FrameParams *fp = fps + blockIdx.y;
    const int nprog = (4 * PROG_STRIDE + 1) * fp->nmby;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nprog; i += gridDim.x * blockDim.x) fp->row_progress[i] = 0;
    if (blockIdx.x == 0)
    {
        if (threadIdx.x < 16) fp->out_info[threadIdx.x] = 0;
        if (threadIdx.x < FS_WORDS) fp->fsync[threadIdx.x] = (threadIdx.x == FS_LIVE || threadIdx.x == FS_LIVE + 1) ? fp->clusters[threadIdx.x - FS_LIVE] : 0;
        if (blockIdx.y == 0 && threadIdx.x < 16) tickets[threadIdx.x] = 0;
        if (threadIdx.x < 8) fp->me_count[threadIdx.x] = 0;
    }
}

/* temporal noise suppressor (h264_denoise.h): thread = 4 samples of one plane; blockIdx.y = job.  Jobs that do not
 * ask for it have dn_out == NULL. */
__global__ void __launch_bounds__(256) k_denoise(const FrameParams *fps, int njobs)
{
    const FrameParams *fp = fps + blockIdx.y;
    if (!fp->dn_out[0]) return;
    for (int pl = 0; pl < 3; pl++)
    {
        const int w = pl ? fp->width >> 1 : fp->width, h = pl ? fp->height >> 1 : fp->height;
        if (w <= 2 || h <= 2) continue;
        const int wpr = (w + 3) >> 2;
        const long n = (long)wpr * h;
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        {
            const int y = (int)(i / wpr), x0 = (int)(i - (long)y * wpr) * 4;
            denoise_word(fp->dn_src[pl], fp->dn_src_stride[pl], fp->dn_prev[pl], fp->dn_out[pl], fp->dn_stride[pl], w, h, x0, y);
        }
    }
}

/* half-sample planes of the new reference picture (after deblocking and border extension) */
__global__ void __launch_bounds__(256) k_hpel(const FrameParams *fps, int njobs)
{
    const FrameParams *fp = fps + blockIdx.y;
    if (fp->fsync[FS_STATE] != FS_DONE || !fp->update_ref) return;
    const long nwords = fp->luma_bytes >> 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (long)gridDim.x * blockDim.x)
        hpel_plane_word(fp, i);
}

__global__ void k_cavlc(const FrameParams *fps, int njobs)
{
    const FrameParams *fp = fps + blockIdx.y;
    const int nmb = fp->nmbx * fp->nmby;
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n > nmb || fp->fsync[FS_STATE] != FS_DONE) return;
    fp->mb_nbits[n] = cavlc_mb(fp, n);
}

/* one block per frame: exclusive scan of mb_nbits -> mb_bitoff, totals, and zero the payload */
__global__ void __launch_bounds__(1024) k_scan(const FrameParams *fps, int njobs)
{
    const FrameParams *fp = fps + blockIdx.x;
    if (fp->fsync[FS_STATE] != FS_DONE) return;
    const int cnt = fp->nmbx * fp->nmby + 1;
    __shared__ int part[1024];
    const int tid = threadIdx.x;
    const int chunk = (cnt + 1023) / 1024;
    const int lo = tid * chunk, hi = min(lo + chunk, cnt);
    int s = 0, maxb = 0;
    for (int i = lo; i < hi; i++) { int b = fp->mb_nbits[i]; s += b; maxb = max(maxb, b); }
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1)
    {
        int v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int base = fp->hdr_bits + part[tid] - s;
    for (int i = lo; i < hi; i++) { int b = fp->mb_nbits[i]; fp->mb_bitoff[i] = base; base += b; }
    const int total = fp->hdr_bits + part[1023];
    if (maxb > MB_BITS_WORDS * 32 - 64) atomicOr(&fp->out_info[1], 1);
    if ((total + 95) / 32 > fp->out_cap_words) atomicOr(&fp->out_info[1], 2);
    if (tid == 0)
    {
        fp->out_info[0] = total;
        int run = 0;
        if (fp->slice_type == SLICE_P) for (int k = cnt - 2; k >= 0 && fp->mbi[k].type == MBT_SKIP; k--) run++;
        fp->out_info[2] = run;
    }
    const int nw = min((total + 95) / 32, fp->out_cap_words);
    for (int i = tid; i < nw; i += 1024) fp->out_words[i] = 0;
}

__global__ void k_pack(const FrameParams *fps, int njobs)
{
    const FrameParams *fp = fps + blockIdx.y;
    const int nmb = fp->nmbx * fp->nmby;
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n > nmb || fp->out_info[1] || fp->fsync[FS_STATE] != FS_DONE) return;
    int nb = fp->mb_nbits[n];
    if (nb) pack_mb(fp, n, nb, fp->mb_bitoff[n]);
}

/* ------------------------------------------------------------------------------ */
/* SAD-map pre-pass (h264_sadmap.h)                                                 */
/* ------------------------------------------------------------------------------ */
/* One CTA per macroblock, one THREAD per tabulated position: 225 full-sample offsets around the centre, then 169
 * quarter-sample positions around the best of them.  The reference window (30 rows x 36 bytes) and, for the second part,
 * the windows of the planes G, b, h, j (20 rows x 24 bytes each) sit in shared memory; a thread walks the 16 rows of its
 * block: five aligned words per row and source, four funnel shifts, four packed-byte SADs against the input row
 * (one 16-byte broadcast load).  No shuffles, no divergence, ~250 / ~500 instructions per thread. */
#define SMW_STRIDE 10
#define SMP_STRIDE 7
__global__ void __launch_bounds__(256) k_sadmap(const FrameParams *fps, int njobs)
{
    const FrameParams *fp = fps + blockIdx.y;
    if (!fp->use_sadmap || fp->slice_type != SLICE_P) return;
    const int nmbx = fp->nmbx, nmb = nmbx * fp->nmby, st = fp->stride[0];
    const int xmax = nmbx * 16 + 12, ymax = fp->nmby * 16 + 15;
    __shared__ __align__(16) uint32_t s_inp[64];
    __shared__ uint32_t s_win[30 * SMW_STRIDE];
    __shared__ uint32_t s_pl[4][20 * SMP_STRIDE];
    __shared__ int s_red[8];
    const int t = threadIdx.x;
    for (int n = blockIdx.x; n < nmb; n += gridDim.x)
    {
        const int mby = n / nmbx, mbx = n - mby * nmbx;
        uint32_t *rec = fp->sadmap + (size_t)n * SM_WORDS;
        int cx, cy;
        sadmap_center(fp, n, &cx, &cy);
        if (t < 64) s_inp[t] = sadmap_inp_word(fp, mbx, mby, t >> 2, t & 3);
        const int wx0 = mbx * 16 + cx - SM_R, wy0 = mby * 16 + cy - SM_R, ax0 = wx0 & ~3, sh0 = wx0 - ax0;
        for (int i = t; i < 30 * 9; i += 256)
        {
            const int r = i / 9, wd = i - r * 9;
            const int y = min(max(wy0 + r, -16), ymax), x = min(max(ax0 + 4 * wd, -16), xmax);
            s_win[r * SMW_STRIDE + wd] = *(const uint32_t *)(fp->ref[0] + (long)y * st + x);
        }
        __syncthreads();
        int key = 0x7FFFFFFF;
        if (t < SM_INT_ENTRIES)
        {
            const int dy = t / SM_N, dx = t - dy * SM_N;
            uint32_t lo = SM_INVALID, hi = SM_INVALID;
            if (sadmap_block_inside(fp, wx0 + dx, wy0 + dy))
            {
                const int col = sh0 + dx, sh = (col & 3) * 8;
                const uint32_t *row = s_win + dy * SMW_STRIDE + (col >> 2);
                int q0 = 0, q1 = 0, q2 = 0, q3 = 0;
#pragma unroll
                for (int r = 0; r < 16; r++)
                {
                    const uint32_t w0 = row[0], w1 = row[1], w2 = row[2], w3 = row[3], w4 = row[4];
                    const uint4 in = *(const uint4 *)(s_inp + 4 * r);
                    const int l = (int)(__vsadu4(__funnelshift_r(w0, w1, sh), in.x) + __vsadu4(__funnelshift_r(w1, w2, sh), in.y));
                    const int rr = (int)(__vsadu4(__funnelshift_r(w2, w3, sh), in.z) + __vsadu4(__funnelshift_r(w3, w4, sh), in.w));
                    if (r < 8) { q0 += l; q1 += rr; } else { q2 += l; q3 += rr; }
                    row += SMW_STRIDE;
                }
                lo = (uint32_t)q0 | ((uint32_t)q1 << 16); hi = (uint32_t)q2 | ((uint32_t)q3 << 16);
                key = ((q0 + q1 + q2 + q3) << 8) | t;
            }
            *(uint2 *)(rec + SM_INT_OFF + 2 * t) = make_uint2(lo, hi);
        }
        /* position with the smallest 16x16 SAD (first in scan order on ties): centre of the quarter map */
#pragma unroll
        for (int o = 16; o; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
        if ((t & 31) == 0) s_red[t >> 5] = key;
        __syncthreads();
        int bt = s_red[0];
#pragma unroll
        for (int k = 1; k < 8; k++) bt = min(bt, s_red[k]);
        bt = bt == 0x7FFFFFFF ? SM_R * SM_N + SM_R : (bt & 255);
        const int bx = cx + (bt % SM_N) - SM_R, by = cy + (bt / SM_N) - SM_R;
        const int ox = mbx * 16 + bx - 2, oy = mby * 16 + by - 2, pax0 = ox & ~3, psh0 = ox - pax0;
        for (int i = t; i < 4 * 20 * 6; i += 256)
        {
            const int pl = i / 120, k = i - pl * 120, r = k / 6, wd = k - r * 6;
            const int y = min(max(oy + r, -16), ymax), x = min(max(pax0 + 4 * wd, -16), xmax);
            const pix_t *plane = pl == 0 ? fp->ref[0] : fp->hp[pl - 1];
            s_pl[pl][r * SMP_STRIDE + wd] = *(const uint32_t *)(plane + (long)y * st + x);
        }
        __syncthreads();
        if ((t & ~31) < 4 * SM_Q_ENTRIES)              /* whole warps (the shuffles below): surplus lanes redo position 0 and store nothing */
        {
            /* quarter map: thread = (position, group of four rows); the four threads of a position sit in one warp and
             * combine their partial quadrant sums with two shuffles */
            const int act = t < 4 * SM_Q_ENTRIES;
            const int pq = act ? t >> 2 : 0, g = t & 3;
            const int qyi = pq / SM_QN, qxi = pq - qyi * SM_QN, qx = qxi - SM_QR, qy = qyi - SM_QR;
            const int aqx = (mbx * 16 + bx) * 4 + qx, aqy = (mby * 16 + by) * 4 + qy;
            const int inside = sadmap_block_inside(fp, aqx >> 2, aqy >> 2);
            int ql = 0, qr = 0;
            {
                /* the two sources of the position (interp_luma_word): plane, column / row offset inside the staged windows */
                const int dx = qx & 3, dy = qy & 3, fx = (qx >> 2) + 2, fy = (qy >> 2) + 2, pos = 1 << (dx + 4 * dy);
                int pa = 0, ca = fx, ra = fy, pb = -1, cb = 0, rb = 0;
                if (pos != 1)
                {
                    pa = -1;
                    if (pos & 0xe0ee) { pa = 1; ca = fx; ra = fy + ((pos & 0xe000) ? 1 : 0); }
                    if (pos & 0xbbb0) { const int c2 = fx + ((pos & 0x8880) ? 1 : 0); if (pa >= 0) { pb = 2; cb = c2; rb = fy; } else { pa = 2; ca = c2; ra = fy; } }
                    if (pos & 0x4e40) { if (pa >= 0) { pb = 3; cb = fx; rb = fy; } else { pa = 3; ca = fx; ra = fy; } }
                    if ((pos & 0xfafa) && pb < 0) { pb = 0; cb = fx + ((dx + 1) >> 2); rb = fy + ((dy + 1) >> 2); }
                }
                const int cola = psh0 + ca, sha = (cola & 3) * 8, colb = psh0 + cb, shb = (colb & 3) * 8;
                const uint32_t *rowa = s_pl[pa] + (ra + 4 * g) * SMP_STRIDE + (cola >> 2);
                const uint32_t *rowb = s_pl[pb < 0 ? 0 : pb] + (rb + 4 * g) * SMP_STRIDE + (colb >> 2);
#pragma unroll
                for (int r = 0; r < 4; r++)
                {
                    uint32_t a0 = __funnelshift_r(rowa[0], rowa[1], sha), a1 = __funnelshift_r(rowa[1], rowa[2], sha);
                    uint32_t a2 = __funnelshift_r(rowa[2], rowa[3], sha), a3 = __funnelshift_r(rowa[3], rowa[4], sha);
                    if (pb >= 0)
                    {
                        a0 = __vavgu4(a0, __funnelshift_r(rowb[0], rowb[1], shb)); a1 = __vavgu4(a1, __funnelshift_r(rowb[1], rowb[2], shb));
                        a2 = __vavgu4(a2, __funnelshift_r(rowb[2], rowb[3], shb)); a3 = __vavgu4(a3, __funnelshift_r(rowb[3], rowb[4], shb));
                    }
                    const uint4 in = *(const uint4 *)(s_inp + 4 * (4 * g + r));
                    ql += (int)(__vsadu4(a0, in.x) + __vsadu4(a1, in.y)); qr += (int)(__vsadu4(a2, in.z) + __vsadu4(a3, in.w));
                    rowa += SMP_STRIDE; rowb += SMP_STRIDE;
                }
            }
            /* rows 0-7 (groups 0, 1): q0 | q1;  rows 8-15 (groups 2, 3): q2 | q3 */
            uint32_t half = (uint32_t)ql | ((uint32_t)qr << 16);
            half += __shfl_xor_sync(0xffffffffu, half, 1);
            const uint32_t other = __shfl_down_sync(0xffffffffu, half, 2);
            if (g == 0 && act) *(uint2 *)(rec + SM_Q_OFF + 2 * pq) = inside ? make_uint2(half, other) : make_uint2(SM_INVALID, SM_INVALID);
        }
        if (t == 0) { rec[0] = (uint32_t)mv_pack(cx, cy); rec[1] = (uint32_t)mv_pack(bx, by); rec[2] = 1; rec[3] = 0; rec[SM_ME_OFF + ME_KEY + 15] = 0; }
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------------ */
/* host side                                                                        */
/* ------------------------------------------------------------------------------ */
struct h264b200_ctx
{
    int device;
    int width, height, nmbx, nmby, nmb;
    int stride[2];
    int inp_stride[3];
    pix_t *d_frames[2];
    pix_t *d_hpel;                /* half-sample planes b, h, j of the current reference picture */
    size_t luma_bytes;
    size_t plane_off[3];
    pix_t *d_inp[3];
    /* two input buffers (d_inp points at the one the current frame uses).  h264b200_prefetch_input names the NEXT frame;
     * its copy into the other buffer is issued by the submission of the current frame, behind that submission's own
     * (small) uploads, on a copy stream: it overlaps the current frame's kernels. */
    pix_t *d_inb[2][3]; int inb_cur;
    struct { const unsigned char *yuv[3]; int stride[3]; } want;   /* named by prefetch_input, not copied yet */
    int want_valid;
    struct { const unsigned char *yuv[3]; int stride[3]; cudaEvent_t ev; int ttl; } stg;   /* staged in d_inb[1 - inb_cur] */
    pix_t *d_clip; int clip_frames;
    pix_t *d_dn[2]; int dn_cur;   /* temporal noise suppressor: previous / new filtered picture (layout of d_inp), allocated on first use */
    int cur;
    int last_dec;                 /* index (into d_frames) of the picture the last encoded frame was reconstructed into */
    cudaStream_t copy_stream;     /* stream of the staged input copies (per context: it belongs to the context's device) */
    MBInfo *d_mbi;
    int16_t *d_coef;
    uint32_t *d_mb_bits;
    int *d_mb_nbits, *d_mb_bitoff;
    uint32_t *d_out_words;
    int out_cap_words;
    int *d_out_info;
    int32_t *d_clusters;
    int cost_stat_valid;          /* a P frame has been finished: d_cost_stat holds its statistic */
    int *d_cost_stat;             /* [2 + nmby] inter-cost thresholds derived from the previous P frame (FrameParams::cost_stat) */
    MBSpec *d_spec; int32_t *d_cl_true; int32_t *d_cl_ckpt; int *d_changed_pass; int *d_need_reenc; int *d_fsync;
    int have_traj; int stats[8]; long long dbg[8];
    int *d_prof;
    uint32_t *d_sadmap;           /* [nmb][SM_WORDS] SAD-map records of the frame being encoded (h264_sadmap.h) */
    int32_t *d_me_field;          /* [nmb][16] motion field predicted by the motion-estimation pre-pass (h264_wave.h) */
    int *d_me_list;               /* [8 + nmb] per-round counters, then the list of macroblocks a refinement round recomputes */
    int *d_progress;              /* 2 * nmby */
    uint32_t *h_out_words;        /* pinned */
    int *h_out_info;              /* pinned */
};

void h264b200_launch_check1(const FrameParams *fps, int njobs, int pass, cudaStream_t st);    /* shim_check.cu */
void h264b200_launch_me(const FrameParams *fps, int njobs, int max_nmb, int round, cudaStream_t st);
void h264b200_launch_intra_check(const FrameParams *fps, int njobs, int max_nmb, cudaStream_t st);

/* Submission lanes.  Every host thread that submits work gets its own lane: a CUDA stream pair, its
 * FrameParams staging, tickets and events.  Encoder instances are independent (reference: "distinct
 * encoders are independent", SURVEY 8(b) Threading), so threads driving different encoders run
 * concurrently on the device: the latency-bound kernels of one lane (re-check, repair rounds, in-loop
 * filter) overlap the macroblock sweep of another.  More threads than lanes share lanes (mutex). */
#define MAX_LANES 64
struct Lane
{
    std::mutex lock;
    cudaStream_t stream, stream2;
    FrameParams *d_fps, *h_fps;
    int fps_cap;
    int *d_tickets;
    int *d_info, *h_info;         /* gathered per-job results of a submission (device / pinned host), 12 ints per job */
    cudaEvent_t ev[6], ev_fork, ev_join;
    cudaEvent_t ev_x[7];           /* [0], [1]: around the entropy-coding kernels on stream2; [2], [3], [4]: around the two pre-passes */
    int ev_ok;
    float last_ms[8];
};
static Lane g_lanes[MAX_LANES];
static std::atomic<int> g_lane_next(0);
static std::atomic<long> g_launches(0);
/* a lane's streams, events and staging belong to ONE device: a thread gets one lane per device it submits to */
#define MAX_DEVICES 16
static thread_local Lane *t_lanes[MAX_DEVICES];
static thread_local Lane *t_lane = NULL;
static Lane *lane_get(int device)
{
    Lane *&l = t_lanes[device & (MAX_DEVICES - 1)];
    if (!l) l = &g_lanes[g_lane_next.fetch_add(1) % MAX_LANES];
    t_lane = l;
    return l;
}
/* every entry point binds the calling thread to the context's device first (contexts are created on the device that
 * is current or named at creation; the caller's current device may be anything afterwards) */
static Lane *lane_enter(const h264b200_ctx *c);
#define g_stream (t_lane->stream)
#define g_stream2 (t_lane->stream2)
#define g_d_fps (t_lane->d_fps)
#define g_h_fps (t_lane->h_fps)
#define g_fps_cap (t_lane->fps_cap)
#define g_d_tickets (t_lane->d_tickets)
#define g_d_info (t_lane->d_info)
#define g_h_info (t_lane->h_info)
#define g_ev (t_lane->ev)
#define g_ev_ok (t_lane->ev_ok)
#define g_ev_fork (t_lane->ev_fork)
#define g_ev_join (t_lane->ev_join)
#define g_last_ms (t_lane->last_ms)
#define g_ev_x (t_lane->ev_x)

static int g_enc_dyn_smem = 0;     /* developer knob H264B200_ENC_SMEM: extra dynamic shared memory per CTA of k_encode_rows
                                       (limits the CTAs resident per SM, to study cache contention) */
static std::once_flag g_knob_once;
static int g_no_sadmap = 0;        /* developer knob H264B200_NO_SADMAP: pixel flavour of the searches everywhere (A/B runs) */
static int g_no_me = 0;            /* developer knob H264B200_NO_ME_PREPASS: no speculative motion estimation ahead of the wavefront */
static int g_me_rounds = ME_ROUNDS; /* developer knob H264B200_ME_ROUNDS */
static int g_thr_eighths = 13;     /* developer knob H264B200_THR: in-sweep intra evaluation from thr/8 of the mean inter cost */
static int g_no_intra_spec = 0;    /* developer knob H264B200_NO_INTRA_SPEC: intra modes evaluated inside sweep 0 as the reference does */
static int ensure_globals(int njobs)
{
    std::call_once(g_knob_once, []() {
        g_no_sadmap = getenv("H264B200_NO_SADMAP") != NULL;
        g_no_me = getenv("H264B200_NO_ME_PREPASS") != NULL;
        if (getenv("H264B200_THR")) g_thr_eighths = atoi(getenv("H264B200_THR"));
        g_no_intra_spec = getenv("H264B200_NO_INTRA_SPEC") != NULL;
        if (getenv("H264B200_NO_FAST")) { int one = 1; cudaMemcpyToSymbol(g_d_no_fast, &one, sizeof(one)); }
        if (getenv("H264B200_ME_ROUNDS")) g_me_rounds = atoi(getenv("H264B200_ME_ROUNDS"));
        const char *e = getenv("H264B200_ENC_SMEM");
        if (e) { g_enc_dyn_smem = atoi(e); cudaFuncSetAttribute(k_encode_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, g_enc_dyn_smem); cudaFuncSetAttribute(k_encode_rows_i, cudaFuncAttributeMaxDynamicSharedMemorySize, g_enc_dyn_smem); }
    });
    if (!g_stream) CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    if (!g_d_tickets) CK(cudaMalloc(&g_d_tickets, 64));
    if (!g_ev_ok) { for (int i = 0; i < 6; i++) CK(cudaEventCreate(&g_ev[i])); for (int i = 0; i < 7; i++) CK(cudaEventCreate(&g_ev_x[i])); g_ev_ok = 1; }
    if (njobs > g_fps_cap)
    {
        if (g_d_fps) cudaFree(g_d_fps);
        if (g_h_fps) cudaFreeHost(g_h_fps);
        int cap = njobs < 16 ? 16 : njobs * 2;
        CK(cudaMalloc(&g_d_fps, sizeof(FrameParams) * cap));
        CK(cudaMallocHost(&g_h_fps, sizeof(FrameParams) * cap));
        if (g_d_info) cudaFree(g_d_info);
        if (g_h_info) cudaFreeHost(g_h_info);
        CK(cudaMalloc(&g_d_info, sizeof(int) * 20 * cap));
        CK(cudaMallocHost(&g_h_info, sizeof(int) * 20 * cap));
        g_fps_cap = cap;
    }
    return 0;
}

extern "C" void h264b200_ctx_destroy(h264b200_ctx *c);
extern "C" int h264b200_ctx_create(h264b200_ctx **out, int width, int height, int device)
{
    int ndev = 0;
    *out = NULL;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return -1;
    if (device >= 0) CK(cudaSetDevice(device));
    int dev = 0;
    CK(cudaGetDevice(&dev));
    h264b200_ctx *c = (h264b200_ctx *)calloc(1, sizeof(*c));
    if (!c) return -3;
    /* a failed allocation releases everything allocated before it (h264b200_ctx_destroy copes with NULL members) */
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "h264b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); h264b200_ctx_destroy(c); return -3; } } while (0)
    c->device = dev;
    c->width = width; c->height = height;
    c->nmbx = (width + 15) >> 4; c->nmby = (height + 15) >> 4; c->nmb = c->nmbx * c->nmby;
    const int w = c->nmbx * 16, h = c->nmby * 16;
    c->stride[0] = w + 32; c->stride[1] = (w + 32) / 2;
    const size_t ysz = (size_t)c->stride[0] * (h + 32), csz = (size_t)c->stride[1] * (h / 2 + 16);
    c->plane_off[0] = (size_t)c->stride[0] * 16 + 16;
    c->plane_off[1] = ysz + (size_t)c->stride[1] * 8 + 8;
    c->plane_off[2] = ysz + csz + (size_t)c->stride[1] * 8 + 8;
    for (int i = 0; i < 2; i++)
    {
        CKC(cudaMalloc(&c->d_frames[i], ysz + 2 * csz + 256));
        CKC(cudaMemset(c->d_frames[i], 0, ysz + 2 * csz + 256));
    }
    c->luma_bytes = ysz;
    CKC(cudaMalloc(&c->d_hpel, 3 * ysz + 256));
    CKC(cudaMemset(c->d_hpel, 0, 3 * ysz + 256));
    c->inp_stride[0] = (width + 63) & ~63;
    c->inp_stride[1] = c->inp_stride[2] = (width / 2 + 63) & ~63;
    {   /* one allocation, planes back to back: a tightly packed I420 frame is then a single copy */
        const size_t s0 = (size_t)c->inp_stride[0] * height, s1 = (size_t)c->inp_stride[1] * (height / 2);
        for (int b = 0; b < 2; b++)
        {
            CKC(cudaMalloc(&c->d_inb[b][0], s0 + 2 * s1 + 256));
            c->d_inb[b][1] = c->d_inb[b][0] + s0;
            c->d_inb[b][2] = c->d_inb[b][1] + s1;
        }
        CKC(cudaEventCreateWithFlags(&c->stg.ev, cudaEventDisableTiming));
        for (int k = 0; k < 3; k++) c->d_inp[k] = c->d_inb[0][k];
    }
    CKC(cudaMalloc(&c->d_mbi, sizeof(MBInfo) * c->nmb));
    CKC(cudaMemset(c->d_mbi, 0, sizeof(MBInfo) * c->nmb));
    CKC(cudaMalloc(&c->d_coef, sizeof(int16_t) * COEF_PER_MB * (size_t)c->nmb));
    CKC(cudaMemset(c->d_coef, 0, sizeof(int16_t) * COEF_PER_MB * (size_t)c->nmb));
    CKC(cudaMalloc(&c->d_mb_bits, sizeof(uint32_t) * MB_BITS_WORDS * (size_t)(c->nmb + 1)));
    CKC(cudaMalloc(&c->d_mb_nbits, sizeof(int) * (c->nmb + 2)));
    CKC(cudaMalloc(&c->d_mb_bitoff, sizeof(int) * (c->nmb + 2)));
    c->out_cap_words = c->nmb * 160 + 1024;
    CKC(cudaMalloc(&c->d_out_words, sizeof(uint32_t) * (size_t)c->out_cap_words));
    CKC(cudaMalloc(&c->d_out_info, 64));
    CKC(cudaMalloc(&c->d_clusters, 16));
    CKC(cudaMemset(c->d_clusters, 0, 16));
    CKC(cudaMalloc(&c->d_cost_stat, sizeof(int) * (2 + c->nmby)));
    CKC(cudaMemset(c->d_cost_stat, 0, sizeof(int) * (2 + c->nmby)));
    CKC(cudaMalloc(&c->d_progress, sizeof(int) * (4 * PROG_STRIDE + 1) * c->nmby));
    CKC(cudaMalloc(&c->d_spec, sizeof(MBSpec) * c->nmb));
    CKC(cudaMemset(c->d_spec, 0, sizeof(MBSpec) * c->nmb));
    CKC(cudaMalloc(&c->d_cl_true, sizeof(int32_t) * 2 * c->nmb));
    CKC(cudaMemset(c->d_cl_true, 0, sizeof(int32_t) * 2 * c->nmb));
    CKC(cudaMalloc(&c->d_cl_ckpt, sizeof(int32_t) * 2 * (c->nmb / 32 + 2)));
    CKC(cudaMemset(c->d_cl_ckpt, 0, sizeof(int32_t) * 2 * (c->nmb / 32 + 2)));
    CKC(cudaMalloc(&c->d_changed_pass, sizeof(int) * c->nmb));
    CKC(cudaMalloc(&c->d_need_reenc, sizeof(int) * c->nmb));
    CKC(cudaMemset(c->d_need_reenc, 0, sizeof(int) * c->nmb));
    CKC(cudaMalloc(&c->d_fsync, sizeof(int) * FS_WORDS));
    CKC(cudaMalloc(&c->d_sadmap, sizeof(uint32_t) * SM_WORDS * (size_t)c->nmb + 256));
    CKC(cudaMemset(c->d_sadmap, 0, sizeof(uint32_t) * SM_WORDS * (size_t)c->nmb + 256));
    CKC(cudaMalloc(&c->d_me_field, sizeof(int32_t) * 16 * (size_t)c->nmb));
    CKC(cudaMemset(c->d_me_field, 0, sizeof(int32_t) * 16 * (size_t)c->nmb));
    CKC(cudaMalloc(&c->d_me_list, sizeof(int) * (8 + (size_t)c->nmb)));
    CKC(cudaMemset(c->d_me_list, 0, sizeof(int) * (8 + (size_t)c->nmb)));
#if defined(H264_PROFILE) || defined(H264_FASTPROF)
    CKC(cudaMalloc(&c->d_prof, sizeof(int) * 20 * c->nmb));
    CKC(cudaMemset(c->d_prof, 0, sizeof(int) * 20 * c->nmb));
#endif
    CKC(cudaMallocHost(&c->h_out_words, sizeof(uint32_t) * (size_t)c->out_cap_words));
    CKC(cudaMallocHost(&c->h_out_info, 128));
    /* The clears above run on the legacy default stream, which does not order itself against the lanes' NON-BLOCKING
     * streams: without this wait the first submission could start before (or while) they execute -- seen as rare wrong
     * first frames when several host threads create and drive encoders at the same time (tools/stress_threads.py). */
    CKC(cudaStreamSynchronize(0));
#undef CKC
    *out = c;
    return 0;
}

extern "C" void h264b200_ctx_destroy(h264b200_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < 2; i++) cudaFree(c->d_frames[i]);
    cudaFree(c->d_hpel);
    for (int b = 0; b < 2; b++) cudaFree(c->d_inb[b][0]);
    if (c->stg.ev) cudaEventDestroy(c->stg.ev);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->d_prof) cudaFree(c->d_prof);
    cudaFree(c->d_sadmap); cudaFree(c->d_me_field); cudaFree(c->d_me_list);
    if (c->d_clip) cudaFree(c->d_clip);
    for (int i = 0; i < 2; i++) if (c->d_dn[i]) cudaFree(c->d_dn[i]);
    cudaFree(c->d_mbi); cudaFree(c->d_coef); cudaFree(c->d_mb_bits); cudaFree(c->d_mb_nbits); cudaFree(c->d_mb_bitoff);
    cudaFree(c->d_out_words); cudaFree(c->d_out_info); cudaFree(c->d_clusters); cudaFree(c->d_progress); cudaFree(c->d_cost_stat);
    cudaFree(c->d_spec); cudaFree(c->d_cl_true); cudaFree(c->d_cl_ckpt); cudaFree(c->d_changed_pass); cudaFree(c->d_need_reenc); cudaFree(c->d_fsync);
    if (c->h_out_words) cudaFreeHost(c->h_out_words);
    if (c->h_out_info) cudaFreeHost(c->h_out_info);
    free(c);
}

extern "C" void h264b200_ctx_reset(h264b200_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaMemset(c->d_clusters, 0, 16);
    cudaStreamSynchronize(0);           /* default-stream clear before anything a lane's non-blocking stream does with it */
    c->cur = 0; c->last_dec = 0;
    c->have_traj = 0; c->cost_stat_valid = 0;
    c->stg.ttl = 0; c->want_valid = 0;
    /* the noise suppressor starts from an all-zero "previous picture" (H:6345-6349) */
    for (int i = 0; i < 2; i++) if (c->d_dn[i]) { cudaFree(c->d_dn[i]); c->d_dn[i] = NULL; }
    c->dn_cur = 0;
}

static void build_fp(const h264b200_job *job, FrameParams *fp)
{
    h264b200_ctx *c = job->ctx;
    const h264b200_frame_params &p = job->p;
    memset(fp, 0, sizeof(*fp));
    fp->width = c->width; fp->height = c->height; fp->nmbx = c->nmbx; fp->nmby = c->nmby;
    fp->slice_type = p.slice_type; fp->qp = p.qp; fp->speed = p.speed; fp->disable_deblock = p.disable_deblock;
    fp->lambda_q4 = p.lambda_q4; fp->lambda_mv_q4 = p.lambda_mv_q4; fp->lambda_i4_q4 = p.lambda_i4_q4;
    fp->lambda_i16_q4 = p.lambda_i16_q4; fp->skip_thr_inter = p.skip_thr_inter; fp->skip_thr_i4x4 = p.skip_thr_i4x4;
    fp->mvlim_x0 = -14 * 4; fp->mvlim_y0 = -14 * 4;                                /* H:6322-6324 */
    fp->mvlim_x1 = (c->nmbx * 16 - 2) * 4; fp->mvlim_y1 = (c->nmby * 16 - 2) * 4;
    for (int i = 0; i < 2; i++)
    {
        fp->df_alpha[i] = p.df_alpha[i]; fp->df_beta[i] = p.df_beta[i];
        for (int k = 0; k < 4; k++) fp->df_tc0[i][k] = p.df_tc0[i][k];
    }
    memcpy(fp->qdat, p.qdat, sizeof(fp->qdat));
    for (int i = 0; i < 3; i++)
    {
        if (job->preloaded_index >= 0)
        {
            size_t fs = (size_t)c->width * c->height * 3 / 2, ys = (size_t)c->width * c->height;
            const pix_t *b = c->d_clip + fs * job->preloaded_index;
            fp->inp[i] = i == 0 ? b : (i == 1 ? b + ys : b + ys + ys / 4);
            fp->inp_stride[i] = i ? c->width / 2 : c->width;
        } else { fp->inp[i] = c->d_inp[i]; fp->inp_stride[i] = c->inp_stride[i]; }
        if (p.denoise && c->d_dn[0])
        {   /* the macroblock path reads the filtered picture (H:6692-6693); the filter reads what was the input */
            const size_t s0 = (size_t)c->inp_stride[0] * c->height, s1 = (size_t)c->inp_stride[1] * (c->height / 2);
            const size_t off = i == 0 ? 0 : (i == 1 ? s0 : s0 + s1);
            fp->dn_src[i] = fp->inp[i]; fp->dn_src_stride[i] = fp->inp_stride[i];
            fp->dn_prev[i] = c->d_dn[c->dn_cur] + off;
            fp->dn_out[i] = c->d_dn[c->dn_cur ^ 1] + off;
            fp->dn_stride[i] = c->inp_stride[i];
            fp->inp[i] = fp->dn_out[i]; fp->inp_stride[i] = c->inp_stride[i];
        }
        fp->dec[i] = c->d_frames[c->cur] + c->plane_off[i];
        fp->ref[i] = c->d_frames[c->cur ^ 1] + c->plane_off[i];
    }
    for (int i = 0; i < 3; i++) fp->hp[i] = c->d_hpel + i * c->luma_bytes + c->plane_off[0];
    fp->hp_out = c->d_hpel; fp->dec_base = c->d_frames[c->cur]; fp->luma_bytes = (int)c->luma_bytes;
    fp->update_ref = job->update_ref;
    fp->stride[0] = c->stride[0]; fp->stride[1] = c->stride[1];
    fp->mbi = c->d_mbi; fp->coef = c->d_coef;
    fp->clusters = c->d_clusters;
    fp->cost_stat = c->d_cost_stat;
    fp->thr_eighths = g_thr_eighths;
    fp->have_cost_stat = c->cost_stat_valid;
    fp->row_progress = c->d_progress; fp->row_progress_df = c->d_progress + PROG_STRIDE * c->nmby; fp->row_progress_dfc = c->d_progress + 2 * PROG_STRIDE * c->nmby; fp->row_progress_mv = c->d_progress + 3 * PROG_STRIDE * c->nmby; fp->row_clean = c->d_progress + 4 * PROG_STRIDE * c->nmby;
    fp->mb_bits = c->d_mb_bits; fp->mb_nbits = c->d_mb_nbits; fp->mb_bitoff = c->d_mb_bitoff;
    fp->out_words = c->d_out_words; fp->out_cap_words = c->out_cap_words; fp->out_info = c->d_out_info;
    fp->hdr_bits = p.hdr_bits;
    fp->spec = c->d_spec; fp->cl_true = c->d_cl_true; fp->cl_ckpt = c->d_cl_ckpt; fp->changed_pass = c->d_changed_pass; fp->need_reenc = c->d_need_reenc; fp->fsync = c->d_fsync;
    fp->max_passes = 4096;
    fp->prof = c->d_prof;
    fp->spec_from_prev = (p.slice_type == SLICE_P && c->have_traj && !getenv("H264B200_NO_PREV_TRAJ"));
    fp->sadmap = c->d_sadmap;
    fp->use_sadmap = p.slice_type == SLICE_P && !g_no_sadmap;
    fp->use_me = fp->use_sadmap && !g_no_me;
    fp->spec_no_intra = p.slice_type == SLICE_P && !g_no_intra_spec;
    fp->me_field = c->d_me_field;
    fp->me_count = c->d_me_list; fp->me_list = c->d_me_list + 8;
}

extern "C" int h264b200_preload(h264b200_ctx *c, int nframes, const unsigned char *frames)
{
    if (!c) return -3;
    std::lock_guard<std::mutex> guard(lane_enter(c)->lock);
    if (ensure_globals(1)) return -3;
    size_t fs = (size_t)c->width * c->height * 3 / 2;
    if (c->d_clip) { cudaFree(c->d_clip); c->d_clip = NULL; c->clip_frames = 0; }
    CK(cudaMalloc(&c->d_clip, fs * nframes + 256));
    CK(cudaMemcpy(c->d_clip, frames, fs * nframes, cudaMemcpyHostToDevice));
    c->clip_frames = nframes;
    return 0;
}

/* second stream of the lane (entropy coding beside the in-loop filter, intra verification beside the candidate re-check) */
static int ensure_stream2(void)
{
    if (!g_stream2)
    {
        CK(cudaStreamCreateWithFlags(&g_stream2, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&g_ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&g_ev_join, cudaEventDisableTiming));
    }
    return 0;
}

/* kernels that follow the macroblock sweeps; frames that are not FS_DONE are skipped inside */
static int launch_post(const FrameParams *d_fps, int n, int max_rows, int max_nmb, int cap_words, cudaStream_t st,
                       cudaEvent_t ev_mid)
{
    /* the entropy-coding kernels only read the macroblock records: they run on a second stream next to the
     * in-loop filter (a latency-bound wavefront that leaves most of the chip idle) */
    (void)cap_words;            /* every job carries its own capacity (FrameParams::out_cap_words): pictures of a batch may differ in size */
    if (ensure_stream2()) return -3;
    CK(cudaEventRecord(g_ev_fork, st));
    CK(cudaStreamWaitEvent(g_stream2, g_ev_fork, 0));
    if (ev_mid) CK(cudaEventRecord(g_ev_x[0], g_stream2));
    k_cavlc<<<dim3((max_nmb + 1 + 63) / 64, n), 64, 0, g_stream2>>>(d_fps, n);
    k_scan<<<n, 1024, 0, g_stream2>>>(d_fps, n);
    k_pack<<<dim3((max_nmb + 1 + 127) / 128, n), 128, 0, g_stream2>>>(d_fps, n);
    if (ev_mid) CK(cudaEventRecord(g_ev_x[1], g_stream2));
    CK(cudaEventRecord(g_ev_join, g_stream2));
    if (!ev_mid) CK(cudaMemsetAsync(g_d_tickets + 1, 0, 4, st));      /* host-driven extra passes: the slot was used before */
    k_deblock_rows<<<2 * n * max_rows + n, 32, 0, st>>>(d_fps, n, g_d_tickets);
    k_borders<<<dim3(64, n), 256, 0, st>>>(d_fps, n);
    k_hpel<<<dim3(148, n), 256, 0, st>>>(d_fps, n);
    if (ev_mid) CK(cudaEventRecord(ev_mid, st));
    CK(cudaStreamWaitEvent(st, g_ev_join, 0));
    g_launches += 6;
    return 0;
}

/* out_info[0..3] and fsync[0..7] of every job of the submission, gathered into one buffer: ONE device->host copy
 * instead of two tiny ones per job (every small transfer on a stream costs ~10 us of latency) */
__global__ void k_gather_info(const FrameParams *fps, int n, int *out)
{
    const int j = blockIdx.x, t = threadIdx.x;
    if (j < n && t < 20) out[j * 20 + t] = t < 4 ? fps[j].out_info[t] : (t < 12 ? fps[j].fsync[t - 4] : fps[j].fsync[FS_FAST + t - 12]);
}
static int fetch_info(int n, h264b200_job *jobs, const int *idx, cudaStream_t st, const FrameParams *d_fps)
{
    k_gather_info<<<n, 32, 0, st>>>(d_fps, n, g_d_info);
    CK(cudaMemcpyAsync(g_h_info, g_d_info, sizeof(int) * 20 * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    for (int k = 0; k < n; k++) memcpy(jobs[idx ? idx[k] : k].ctx->h_out_info, g_h_info + 20 * k, sizeof(int) * 20);
    g_launches += 1;
    return 0;
}

/* host planes -> the ctx's device input layout (rows padded to 64 bytes), one copy when the frame is tightly packed */
static int upload_input(h264b200_ctx *c, pix_t *const dst[3], const unsigned char *const yuv[3], const int stride[3], cudaStream_t st)
{
    const size_t ysz = (size_t)c->width * c->height;
    if (c->inp_stride[0] == c->width && c->inp_stride[1] == c->width / 2 &&
        stride[0] == c->width && stride[1] == c->width / 2 && stride[2] == c->width / 2 &&
        yuv[1] == yuv[0] + ysz && yuv[2] == yuv[1] + ysz / 4)
    {
        CK(cudaMemcpyAsync(dst[0], yuv[0], ysz * 3 / 2, cudaMemcpyHostToDevice, st));
        return 0;
    }
    for (int pl = 0; pl < 3; pl++)
    {
        int w = pl ? c->width / 2 : c->width, h = pl ? c->height / 2 : c->height;
        CK(cudaMemcpy2DAsync(dst[pl], c->inp_stride[pl], yuv[pl], stride[pl], w, h, cudaMemcpyHostToDevice, st));
    }
    return 0;
}

static std::atomic<long> g_prefetch_hits(0);
extern "C" long h264b200_prefetch_hits(void) { return g_prefetch_hits; }      /* frames whose staged copy was used */
extern "C" int h264b200_prefetch_input(h264b200_ctx *c, const unsigned char *const yuv[3], const int stride[3])
{
    if (!c || !yuv || !yuv[0]) return -3;
    for (int i = 0; i < 3; i++) { c->want.yuv[i] = yuv[i]; c->want.stride[i] = stride[i]; }
    c->want_valid = 1;
    return 0;
}

static Lane *lane_enter(const h264b200_ctx *c)
{
    cudaSetDevice(c->device);
    return lane_get(c->device);
}

/* one submission: n <= max_jobs_per_submission() frames of contexts that live on the lane's device */
static int encode_chunk(int n, h264b200_job *jobs)
{
    if (ensure_globals(2 * n)) return -3;
    cudaStream_t st = g_stream;
    int max_rows = 0, max_nmb = 0, cap = 0x7fffffff, any_denoise = 0;
    for (int i = 0; i < n; i++)
    {
        h264b200_ctx *c = jobs[i].ctx;
        jobs[i].status = 0;
        if (jobs[i].p.denoise)
        {
            any_denoise = 1;
            for (int k = 0; k < 2; k++)
                if (!c->d_dn[k])
                {
                    const size_t sz = (size_t)c->inp_stride[0] * c->height + 2 * (size_t)c->inp_stride[1] * (c->height / 2) + 256;
                    CK(cudaMalloc(&c->d_dn[k], sz));
                    CK(cudaMemsetAsync(c->d_dn[k], 0, sz, st));      /* on the lane's stream: ordered before the kernels that read it */
                }
        }
        jobs[i].status = 0;
        if (jobs[i].preloaded_index < 0 && c->stg.ttl > 0)
        {   /* the frame staged while the previous one was encoded: switch to its buffer, wait for its copy, no H2D here */
            const int hit = c->stg.yuv[0] == jobs[i].yuv[0] && c->stg.yuv[1] == jobs[i].yuv[1] && c->stg.yuv[2] == jobs[i].yuv[2] &&
                            c->stg.stride[0] == jobs[i].stride[0] && c->stg.stride[1] == jobs[i].stride[1] && c->stg.stride[2] == jobs[i].stride[2];
            c->stg.ttl = 0;
            if (hit)
            {
                g_prefetch_hits++;
                c->inb_cur ^= 1;
                for (int k = 0; k < 3; k++) c->d_inp[k] = c->d_inb[c->inb_cur][k];
                CK(cudaStreamWaitEvent(st, c->stg.ev, 0));
                jobs[i].status = 1;      /* "input already on the device" until the copy loop below */
            } else CK(cudaStreamWaitEvent(st, c->stg.ev, 0));    /* unused copy: let it finish before its buffer can be reused */
        }
        build_fp(&jobs[i], &g_h_fps[i]);
        max_rows = c->nmby > max_rows ? c->nmby : max_rows;
        max_nmb = c->nmb > max_nmb ? c->nmb : max_nmb;
        cap = c->out_cap_words < cap ? c->out_cap_words : cap;
    }
    CK(cudaEventRecord(g_ev[0], st));
    for (int i = 0; i < n; i++)
    {
        h264b200_ctx *c = jobs[i].ctx;
        if (jobs[i].preloaded_index >= 0)
        {
            if (jobs[i].preloaded_index >= c->clip_frames) { jobs[i].status = -3; return -3; }
            continue;
        }
        if (jobs[i].status == 1) { jobs[i].status = 0; continue; }          /* staged ahead of time */
        if (upload_input(c, c->d_inp, jobs[i].yuv, jobs[i].stride, st)) { jobs[i].status = -3; return -3; }
    }
    CK(cudaMemcpyAsync(g_d_fps, g_h_fps, sizeof(FrameParams) * n, cudaMemcpyHostToDevice, st));
    /* per-frame synchronisation state of every job (row counters, output info, fsync + the live cluster state) and the
     * tickets: one small kernel instead of four memsets / copies per job */
    k_frame_init<<<dim3(8, n), 256, 0, st>>>(g_d_fps, n, g_d_tickets);
    g_launches += 1;
    if (any_denoise) { k_denoise<<<dim3(296, n), 256, 0, st>>>(g_d_fps, n); g_launches += 1; }
    /* SAD maps of every macroblock of every P frame of the submission: dependency-free, ahead of the wavefront */
    int any_p = 0, any_pslice = 0;
    for (int i = 0; i < n; i++) { any_p |= g_h_fps[i].use_sadmap; any_pslice |= g_h_fps[i].slice_type == SLICE_P; }
    CK(cudaEventRecord(g_ev_x[2], st));
    if (any_p) { k_sadmap<<<dim3(max_nmb, n), 256, 0, st>>>(g_d_fps, n); g_launches += 1; }
    CK(cudaEventRecord(g_ev_x[3], st));
    /* ... and the motion estimation itself on predicted contexts (round 0 + refinement rounds); the wavefront verifies */
    int any_me = 0;
    for (int i = 0; i < n; i++) any_me |= g_h_fps[i].use_me;
    if (any_me) for (int r = 0; r < g_me_rounds && r < 8; r++) { h264b200_launch_me(g_d_fps, n, max_nmb, r, st); g_launches += r ? 2 : 1; }
    CK(cudaEventRecord(g_ev_x[4], st));
    /* sweep 0 of every frame, then -- optimistically -- everything that follows it */
    CK(cudaEventRecord(g_ev[1], st));
    (any_pslice ? k_encode_rows : k_encode_rows_i)<<<n * max_rows + n, MB_WARPS * 32, g_enc_dyn_smem, st>>>(g_d_fps, n, g_d_tickets, 0);
    {   /* inputs of the NEXT frames named by h264b200_prefetch_input: their copies start now, behind this submission's
         * own uploads, and run under its kernels */
        for (int i = 0; i < n; i++)
        {
            h264b200_ctx *c = jobs[i].ctx;
            if (!c->want_valid) continue;
            c->want_valid = 0;
            if (!c->copy_stream) CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
            if (upload_input(c, c->d_inb[c->inb_cur ^ 1], c->want.yuv, c->want.stride, c->copy_stream)) return -3;
            CK(cudaEventRecord(c->stg.ev, c->copy_stream));
            for (int k = 0; k < 3; k++) { c->stg.yuv[k] = c->want.yuv[k]; c->stg.stride[k] = c->want.stride[k]; }
            c->stg.ttl = 1;
        }
    }
    {   /* P frames: sweep 0 decided among the inter modes only; the intra modes of every macroblock are verified now */
        int any_spec = 0;
        for (int i = 0; i < n; i++) any_spec |= g_h_fps[i].spec_no_intra;
        /* ... on the second stream, beside the candidate re-check: the two read the finished sweep, tag disjoint causes
         * (both write the same tag value, k_after_check looks at the sum) and are bound by different things */
        if (ensure_stream2()) return -3;
        CK(cudaEventRecord(g_ev_fork, st));
        CK(cudaStreamWaitEvent(g_stream2, g_ev_fork, 0));
        CK(cudaEventRecord(g_ev_x[5], g_stream2));
        if (any_spec) { h264b200_launch_intra_check(g_d_fps, n, max_nmb, g_stream2); g_launches += 1; }
        CK(cudaEventRecord(g_ev_x[6], g_stream2));
        CK(cudaEventRecord(g_ev_join, g_stream2));
    }
    h264b200_launch_check1(g_d_fps, n, 1, st);
    CK(cudaStreamWaitEvent(st, g_ev_join, 0));
    k_after_check<<<(n + 63) / 64, 64, 0, st>>>(g_d_fps, n, 1);
    g_launches += 3;
    /* two repair rounds are queued unconditionally (frames that are already exact skip them on
     * the device), so that the common case needs no host round trip before the post-processing */
    for (int pass = 1; pass <= 2; pass++)
    {
        for (int r = 0; r < REPAIR_ROUNDS; r++) k_repair_round<<<dim3(296, n), MB_WARPS * 32, 0, st>>>(g_d_fps, n, pass, r);
        /* ticket slots: [0] sweep 0, [1] in-loop filter, [2], [3] the two repair waves queued here, [4] host-driven passes */
        (any_pslice ? k_encode_rows : k_encode_rows_i)<<<n * max_rows + n, MB_WARPS * 32, g_enc_dyn_smem, st>>>(g_d_fps, n, g_d_tickets + 1 + pass, pass);
        k_replay<<<n, 32, 0, st>>>(g_d_fps, n, pass);
        h264b200_launch_check1(g_d_fps, n, pass + 1, st);
        k_after_check<<<(n + 63) / 64, 64, 0, st>>>(g_d_fps, n, pass + 1);
        g_launches += 4 + REPAIR_ROUNDS;
    }
    CK(cudaEventRecord(g_ev[2], st));
    if (launch_post(g_d_fps, n, max_rows, max_nmb, cap, st, g_ev[3])) return -3;
    CK(cudaEventRecord(g_ev[4], st));
    if (fetch_info(n, jobs, NULL, st, g_d_fps)) return -3;

    /* frames whose speculated mv_clusters candidates did not verify: repair sweeps (h264_wave.h) */
    std::vector<int> dirty;
    for (int i = 0; i < n; i++)
        if (jobs[i].ctx->h_out_info[4 + FS_STATE] != FS_DONE && !(jobs[i].ctx->h_out_info[1] & 4)) dirty.push_back(i);
    FrameParams *h2 = g_h_fps + n, *d2 = g_d_fps + n;
    while (!dirty.empty())
    {
        int m = (int)dirty.size(), rows2 = 0, nmb2 = 0, pass = 0;
        for (int k = 0; k < m; k++)
        {
            h264b200_ctx *c = jobs[dirty[k]].ctx;
            h2[k] = g_h_fps[dirty[k]];
            rows2 = c->nmby > rows2 ? c->nmby : rows2;
            nmb2 = c->nmb > nmb2 ? c->nmb : nmb2;
            pass = c->h_out_info[4 + FS_STATE];       /* sweeps advance in lock step for all dirty frames */
        }
        CK(cudaMemcpyAsync(d2, h2, sizeof(FrameParams) * m, cudaMemcpyHostToDevice, st));
        for (int r = 0; r < REPAIR_ROUNDS; r++) k_repair_round<<<dim3(296, m), MB_WARPS * 32, 0, st>>>(d2, m, pass, r);
        CK(cudaMemsetAsync(g_d_tickets + 4, 0, 4, st));
        k_encode_rows<<<m * rows2 + m, MB_WARPS * 32, g_enc_dyn_smem, st>>>(d2, m, g_d_tickets + 4, pass);
        k_replay<<<m, 32, 0, st>>>(d2, m, pass);
        h264b200_launch_check1(d2, m, pass + 1, st);
        k_after_check<<<(m + 63) / 64, 64, 0, st>>>(d2, m, pass + 1);
        g_launches += 4 + REPAIR_ROUNDS;
        if (launch_post(d2, m, rows2, nmb2, cap, st, NULL)) return -3;
        if (fetch_info(m, jobs, dirty.data(), st, d2)) return -3;
        std::vector<int> still;
        for (int k = 0; k < m; k++)
        {
            h264b200_ctx *c = jobs[dirty[k]].ctx;
            int stt = c->h_out_info[4 + FS_STATE];
            if (stt != FS_DONE && !(c->h_out_info[1] & 4))
            {
                if (stt != pass + 1) { jobs[dirty[k]].status = -4; continue; }
                still.push_back(dirty[k]);
            }
        }
        dirty.swap(still);
    }

    int rc = 0;
    for (int i = 0; i < n; i++)
    {
        h264b200_ctx *c = jobs[i].ctx;
        jobs[i].out_bits = c->h_out_info[0];
        jobs[i].trailing_skip_run = c->h_out_info[2];
        jobs[i].out_words = c->h_out_words;
        c->stats[0] += c->h_out_info[4 + FS_PASSES]; c->stats[1] += c->h_out_info[4 + FS_REENC];
        c->stats[2] += c->h_out_info[4 + FS_CHECKS]; c->stats[3]++;
        c->stats[4] += c->h_out_info[12]; c->stats[5] += c->h_out_info[13];
        for (int k = 0; k < 5; k++) c->dbg[k] = c->h_out_info[14 + k];
        c->have_traj = jobs[i].p.slice_type == SLICE_P;
        if (jobs[i].p.slice_type == SLICE_P && jobs[i].status == 0) c->cost_stat_valid = 1;
        if (jobs[i].status) { if (!rc) rc = jobs[i].status; continue; }
        if (c->h_out_info[1] & 4) { jobs[i].status = -4; if (!rc) rc = -4; continue; }
        if (c->h_out_info[1]) { jobs[i].status = -2; if (!rc) rc = -2; continue; }
        CK(cudaMemcpyAsync(c->h_out_words, c->d_out_words, (size_t)((jobs[i].out_bits + 95) / 32) * 4, cudaMemcpyDeviceToHost, st));
        for (int pl = 0; pl < 3; pl++)
            if (jobs[i].recon[pl])
            {
                /* in-place mode hands us the caller's planes: copy only the visible area */
                int cw = pl ? c->width / 2 : c->width, chh = pl ? c->height / 2 : c->height;
                CK(cudaMemcpy2DAsync(jobs[i].recon[pl], jobs[i].recon_stride[pl],
                                     c->d_frames[c->cur] + c->plane_off[pl], c->stride[pl != 0], cw, chh,
                                     cudaMemcpyDeviceToHost, st));
            }
    }
    CK(cudaEventRecord(g_ev[5], st));
    CK(cudaStreamSynchronize(st));
    for (int i = 0; i < n; i++) if (jobs[i].status == 0) { jobs[i].ctx->last_dec = jobs[i].ctx->cur; if (jobs[i].update_ref) jobs[i].ctx->cur ^= 1; }
    for (int i = 0; i < n; i++) if (jobs[i].p.denoise && jobs[i].ctx->d_dn[0]) jobs[i].ctx->dn_cur ^= 1;
    cudaEventElapsedTime(&g_last_ms[0], g_ev[0], g_ev[5]);
    cudaEventElapsedTime(&g_last_ms[1], g_ev[1], g_ev[2]);
    cudaEventElapsedTime(&g_last_ms[2], g_ev[2], g_ev[3]);
    cudaEventElapsedTime(&g_last_ms[4], g_ev_x[2], g_ev_x[3]);
    cudaEventElapsedTime(&g_last_ms[5], g_ev_x[3], g_ev_x[4]);
    cudaEventElapsedTime(&g_last_ms[6], g_ev_x[5], g_ev_x[6]);
    cudaEventElapsedTime(&g_last_ms[3], g_ev_x[0], g_ev_x[1]);      /* entropy coding: events on ITS stream (it runs beside the in-loop filter) */
    return rc;
}

/* Frames that are not encoded (VBV-overflow "transparent" frames, H:6497-6508) but whose session runs the temporal
 * noise suppressor: the reference filters EVERY submitted picture before it decides what to do with it (H:6686), so
 * the filter's state has to advance: upload, k_denoise, flip -- nothing else. */
static int denoise_only_chunk(int n, h264b200_job *jobs)
{
    if (ensure_globals(2 * n)) return -3;
    cudaStream_t st = g_stream;
    for (int i = 0; i < n; i++)
    {
        h264b200_ctx *c = jobs[i].ctx;
        for (int k = 0; k < 2; k++)
            if (!c->d_dn[k])
            {
                const size_t sz = (size_t)c->inp_stride[0] * c->height + 2 * (size_t)c->inp_stride[1] * (c->height / 2) + 256;
                CK(cudaMalloc(&c->d_dn[k], sz));
                CK(cudaMemsetAsync(c->d_dn[k], 0, sz, st));      /* on the lane's stream: ordered before the kernels that read it */
            }
        if (c->stg.ttl > 0) { c->stg.ttl = 0; CK(cudaStreamWaitEvent(st, c->stg.ev, 0)); }    /* a staged copy is not used here */
        if (jobs[i].preloaded_index >= 0) { if (jobs[i].preloaded_index >= c->clip_frames) return -3; }
        else if (upload_input(c, c->d_inp, jobs[i].yuv, jobs[i].stride, st)) return -3;
        h264b200_job tmp = jobs[i];
        tmp.p.denoise = 1;
        build_fp(&tmp, &g_h_fps[i]);
    }
    CK(cudaMemcpyAsync(g_d_fps, g_h_fps, sizeof(FrameParams) * n, cudaMemcpyHostToDevice, st));
    k_denoise<<<dim3(296, n), 256, 0, st>>>(g_d_fps, n);
    g_launches += 1;
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    for (int i = 0; i < n; i++) { jobs[i].ctx->dn_cur ^= 1; jobs[i].status = 0; jobs[i].out_words = NULL; jobs[i].out_bits = 0; }
    return 0;
}

/* k_encode_rows gives its first n tickets to trajectory followers that spin until the rows of their frame have
 * finished, so a submission must leave resident CTA slots for the rows: at most half of the slots the device can hold
 * are ever followers.  Larger batches are split into consecutive submissions (results are identical: the jobs of a
 * batch are independent). */
static int max_jobs_per_submission()
{
    static std::atomic<int> cached(0);
    int v = cached.load();
    if (v) return v;
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_encode_rows, MB_WARPS * 32, g_enc_dyn_smem) != cudaSuccess) per_sm = 1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 1;
    v = per_sm * sms / 2;
    if (v < 1) v = 1;
    cached.store(v);
    return v;
}

static int encode_impl(int n, h264b200_job *jobs)
{
    if (n <= 0) return 0;
    if (!jobs) return -3;
    for (int i = 0; i < n; i++)
    {
        jobs[i].status = -3; jobs[i].out_words = NULL; jobs[i].out_bits = 0; jobs[i].trailing_skip_run = 0;
        if (!jobs[i].ctx || jobs[i].ctx->device != jobs[0].ctx->device) return -3;       /* one device per submission */
    }
    Lane *lane = lane_enter(jobs[0].ctx);
    std::lock_guard<std::mutex> guard(lane->lock);
    const int cap = max_jobs_per_submission();
    std::vector<h264b200_job> enc, dn;
    std::vector<int> enc_idx, dn_idx;
    for (int i = 0; i < n; i++)
    {
        if (jobs[i].p.denoise == 2) { dn.push_back(jobs[i]); dn_idx.push_back(i); }
        else { enc.push_back(jobs[i]); enc_idx.push_back(i); }
    }
    int rc = 0;
    for (int pass = 0; pass < 2; pass++)
    {
        std::vector<h264b200_job> &v = pass ? enc : dn;
        std::vector<int> &idx = pass ? enc_idx : dn_idx;
        for (size_t b = 0; b < v.size(); b += (size_t)cap)
        {
            const int m = (int)(v.size() - b < (size_t)cap ? v.size() - b : (size_t)cap);
            for (int k = 0; k < m; k++) v[b + k].status = 0;
            const int r = pass ? encode_chunk(m, &v[b]) : denoise_only_chunk(m, &v[b]);
            if (r == -3)
            {   /* a CUDA call failed somewhere in the submission: nothing of it can be trusted.  Every job of the chunk
                 * reports the error, queued work is drained so that the pinned staging can be reused */
                cudaStreamSynchronize(g_stream);
                if (g_stream2) cudaStreamSynchronize(g_stream2);
                cudaGetLastError();
                for (int k = 0; k < m; k++) { v[b + k].status = -3; v[b + k].out_words = NULL; }
            }
            if (r && !rc) rc = r;
            for (int k = 0; k < m; k++) jobs[idx[b + k]] = v[b + k];
        }
    }
    return rc;
}

extern "C" int h264b200_encode_frames(int n, h264b200_job *jobs) { return encode_impl(n, jobs); }

extern "C" int h264b200_get_recon(h264b200_ctx *c, unsigned char *const planes[3], const int strides[3])
{
    if (!c) return -3;
    std::lock_guard<std::mutex> guard(lane_enter(c)->lock);
    if (ensure_globals(1)) return -3;
    /* the picture the last encoded frame was reconstructed into: after a frame that updated the reference (or a transparent
     * frame, whose reconstruction IS the reference picture) that is the current reference; after a droppable frame
     * (update_ref == 0) the reference did not move and the reconstruction sits in the other picture */
    for (int pl = 0; pl < 3; pl++)
    {
        int w = c->nmbx * (pl ? 8 : 16), h = c->nmby * (pl ? 8 : 16);
        CK(cudaMemcpy2DAsync(planes[pl], strides[pl], c->d_frames[c->last_dec] + c->plane_off[pl], c->stride[pl != 0], w, h,
                             cudaMemcpyDeviceToHost, g_stream));
    }
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

/* test hook: the SAD-map / motion-estimation records (h264_sadmap.h) of the last P frame of ctx, [nmb][SM_WORDS] words */
extern "C" int h264b200_debug_get_sadmap(h264b200_ctx *c, unsigned int *out, int max_words)
{
    if (!c || !out) return -3;
    std::lock_guard<std::mutex> guard(lane_enter(c)->lock);
    const int words = SM_WORDS * c->nmb;
    if (max_words < words) return -3;
    CK(cudaMemcpy(out, c->d_sadmap, sizeof(uint32_t) * (size_t)words, cudaMemcpyDeviceToHost));
    return SM_WORDS;
}

extern "C" void h264b200_note_transparent(h264b200_ctx *c) { if (c) c->last_dec = c->cur ^ 1; }

extern "C" int h264b200_last_timing_ex(float *out_ms, int n)
{
    int dev = 0;
    cudaGetDevice(&dev);
    lane_get(dev);
    for (int i = 0; i < n; i++) out_ms[i] = i < 8 ? g_last_ms[i] : 0.f;
    return 8;
}
extern "C" void h264b200_last_timing(float out_ms[4]) { h264b200_last_timing_ex(out_ms, 4); }
/* developer builds (-DH264_PROFILE): per-MB phase cycles of the last frame, [nmb][10] ints */
extern "C" int h264b200_get_profile(h264b200_ctx *c, int *out)
{
    if (!c->d_prof) return -1;
    cudaMemcpy(out, c->d_prof, sizeof(int) * 20 * c->nmb, cudaMemcpyDeviceToHost);
    return 0;
}
extern "C" void h264b200_ctx_stats(h264b200_ctx *c, int out[4]) { for (int i = 0; i < 4; i++) out[i] = c->stats[i]; }
extern "C" int h264b200_ctx_stats_ex(h264b200_ctx *c, int *out, int n) { for (int i = 0; i < n; i++) out[i] = i < 8 ? c->stats[i] : (i < 13 ? (int)c->dbg[i - 8] : 0); return 13; }
extern "C" long h264b200_launch_count(void) { return g_launches; }
extern "C" const char *h264b200_backend_name(void) { return "cuda-sm_100a"; }
