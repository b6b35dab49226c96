/*
 * h264_mbenc.h -- macroblock mode decision, motion estimation, transform /
 * quantisation / reconstruction for one macroblock, executed by one warp
 * (SURVEY.md 8(a) rows a4-a12, a17).  Decisions replay the reference's greedy,
 * order-dependent search exactly (same candidate order, same strict comparisons,
 * same 16-bit cost cache) -- see the H:nnn citations.
 */
#pragma once
#include "h264_common.h"
#include "h264_pixel.h"
#include "h264_sadmap.h"

/* optional per-macroblock phase timing (developer builds with -DH264_PROFILE) */
#if defined(H264_PROFILE) && H264_DEVICE
#  define PROF_N 16
#  define PROF_MEMBERS long long prof_t0, prof_start; int prof[PROF_N];
#  define PROF_INIT(s) do { (s).prof_t0 = (s).prof_start = clock64(); for (int k_ = 0; k_ < PROF_N; k_++) (s).prof[k_] = 0; } while (0)
#  define PROF_MARK(s, k) do { (void)*(volatile int32_t *)&(s).w->scal[15]; long long t_ = clock64(); (s).prof[k] += (int)(t_ - (s).prof_t0); (s).prof_t0 = t_; } while (0)
#  define PROF_STORE(s, fp, n, type) do { if (threadIdx.x == 0 && (fp)->prof) { for (int k_ = 0; k_ < 12; k_++) (fp)->prof[(n) * 20 + k_] = (s).prof[k_]; \
        (fp)->prof[(n) * 20 + 16] = (type); (fp)->prof[(n) * 20 + 19] = ((s).prof[12] & 0xFFFF) | ((s).prof[13] << 16); } } while (0)
/* cycles since the start of the macroblock at which warp `wid` reaches this point (slot 12 + wid / 17 + wid) */
#  define PROF_WARP(s, fp, n, slot) do { (void)*(volatile int32_t *)&(s).w->scal[15]; if ((threadIdx.x & 31) == 0 && (fp)->prof) (fp)->prof[(n) * 20 + (slot)] = (int)(clock64() - (s).prof_start); } while (0)
/* sub-phase of the phase that is being timed: adds the cycles since the last mark to slot k WITHOUT closing the phase's own slot */
#  define PROF_SUB(s, k) do { MBState &s_ = const_cast<MBState &>(s); (void)*(volatile int32_t *)&s_.w->scal[15]; long long t_ = clock64(); s_.prof[k] += (int)(t_ - s_.prof_t0); s_.prof_t0 = t_; } while (0)
#else
#  define PROF_MEMBERS
#  define PROF_INIT(s)
#  define PROF_MARK(s, k)
#  define PROF_STORE(s, fp, n, type)
#  define PROF_WARP(s, fp, n, slot)
#  define PROF_SUB(s, k)
#endif

struct MBState   /* warp-uniform registers of the macroblock being encoded */
{
    PROF_MEMBERS
    const FrameParams *fp;
    MBWork *w;
    int mbx, mby, avail;
    int type;            /* MBT_* */
    int cost;
    int i16_mode;
    int mv_skip_pred;    /* H:674 */
    pix_t *pbest;        /* luma prediction of the chosen mode */
    SearchScratch *ss;   /* private scratch of the search warp running this code */
    int win_x0, win_y0;  /* luma coordinates of the search window's first sample; win_x0 % 4 == 0 */
    int win_ok;
    /* SAD-map record of the macroblock (h264_sadmap.h; shared-memory copy or the record in global memory, NULL: none) */
    const uint32_t *map;
    int map_cx, map_cy, map_qcx, map_qcy;
    int lut;             /* 1: the searches of this macroblock look their SADs up and build no prediction blocks; the
                            prediction of the winning mode is made from its vectors at the end (luma_pred_half) */
};

/* bind the macroblock's SAD-map record (h264_sadmap.h) */
HD void lut_bind(MBState &s, const uint32_t *rec)
{
    s.map = rec;
    s.map_cx = s.map_cy = s.map_qcx = s.map_qcy = 0;
    if (rec)
    {
        s.map_cx = mv_x((int)rec[0]); s.map_cy = mv_y((int)rec[0]);
        s.map_qcx = mv_x((int)rec[1]); s.map_qcy = mv_y((int)rec[1]);
    }
}

/* slots of MBWork::ic, the outcome of the candidate stage (inter_stage_a) */
#define IC_STATE 0        /* 0 not known yet, 1 early skip, 2 continue with the searches */
#define IC_PREF 1         /* bit t: partition mode t is worth a search (H:5224)           */
#define IC_MV_BEST 2
#define IC_SAD_BEST 3     /* incl. MV cost: the min_sad seed of the 16x16 search (H:5414) */
#define IC_MVP16 4
#define IC_MV_SKIP 5
#define IC_SAD_SKIP 6
#define IC_SIG 7          /* [7..10] candidate-stage signature for the speculation check  */
#define IC_COST0 11       /* cost of the 16x16 mode once its search has finished, -1 before */

HD int clz32(uint32_t v)
{
#if H264_DEVICE
    return __clz((int)v);
#else
    return __builtin_clz(v);
#endif
}
HD int bitsize_ue(int v) { return 2 * (32 - clz32((uint32_t)(v + 1))) - 1; }          /* H:3402 */
HD int bits_se(int v) { v = 2 * v - 1; v ^= v >> 31; return bitsize_ue(v); }          /* H:3410 */
HD int mv_cost(int mv, int mvp, int lambda_mv_q4)                                     /* H:4952 */
{
    int nb = bits_se(mv_x(mv) - mv_x(mvp)) + bits_se(mv_y(mv) - mv_y(mvp));
    return (nb * lambda_mv_q4) >> 4;
}
HD int mv_in_rect(int v, int x0, int y0, int x1, int y1)
{
    int x = mv_x(v), y = mv_y(v);
    return y >= y0 && y <= y1 && x >= x0 && x <= x1;
}

/* Exact pruning.  Every partition mode competes with the 16x16 mode, which is always searched,
 * and wins only with a strictly smaller cost (ascending order, strict '<', H:5500); Intra4x4 wins
 * only with a cost strictly below Intra16x16's and below the final inter cost (H:4827).  The 16x16
 * cost is at most lambda(1 bit) + the seed of its search (the search only lowers it), and costs
 * accumulate non-negative terms, so a partial sum that has reached that bound decides the
 * comparison: the remaining partitions / 4x4 blocks need not be evaluated.  Their side
 * effects (MVs, modes, levels) are only kept for the winning mode.  Returns the bound. */
HD int inter_cost_bound(const FrameParams *fp, const MBWork *w)
{
    int bound = ((1 * fp->lambda_q4) >> 4) + *(volatile const int32_t *)&w->ic[IC_SAD_BEST];
    const int c0 = *(volatile const int32_t *)&w->ic[IC_COST0];
    if (c0 >= 0 && c0 < bound) bound = c0;
    return bound;
}
/* upper bound of the cost the inter decision will end with: the best mode's cost, or -- when the raw
 * skip SAD is smaller -- the cost of P16x16 at the skip vector, which may be larger (H:5512-5522) */
HD int inter_final_bound(const FrameParams *fp, const MBWork *w)
{
    int bound = inter_cost_bound(fp, w);
    const int sad_skip = *(volatile const int32_t *)&w->ic[IC_SAD_SKIP];
    if (sad_skip != 0x7FFFFFFF)
    {
        const int alt = sad_skip + mv_cost(*(volatile const int32_t *)&w->ic[IC_MV_SKIP], *(volatile const int32_t *)&w->ic[IC_MVP16], fp->lambda_mv_q4);
        if (alt > bound) bound = alt;
    }
    return bound;
}


/* ------------------------------------------------------------------------------
 * loading the macroblock's inputs
 * ---------------------------------------------------------------------------- */
#if H264_DEVICE
HD void cp_async4(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
HD void cp_async_commit_wait_all()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
}
/* input samples of macroblock (mbx, mby) -> w->pf_inp, asynchronously; only for macroblocks entirely
 * inside the picture whose rows are word-aligned (else mb_load fetches synchronously) */
HD void mb_prefetch_input(const FrameParams *fp, MBWork *w, int mbx, int mby)
{
    const int ok = mbx < fp->nmbx && (mbx + 1) * 16 <= fp->width && (mby + 1) * 16 <= fp->height &&
                   !(((fp->inp_stride[0] | fp->inp_stride[1] | fp->inp_stride[2]) & 3) |
                     (((uintptr_t)fp->inp[0] | (uintptr_t)fp->inp[1] | (uintptr_t)fp->inp[2]) & 3));
    if (ok)
    {
        FOR_THREADS(i, 96)
        {
            if (i < 64) { int r = i >> 2, c = (i & 3) * 4; cp_async4(&w->pf_inp[i], fp->inp[0] + (mby * 16 + r) * fp->inp_stride[0] + mbx * 16 + c); }
            else
            {
                int k2 = i - 64, r = k2 >> 2, q = k2 & 3, pl = q >> 1, c = (q & 1) * 4;
                cp_async4(&w->pf_inp[i], fp->inp[1 + pl] + (mby * 8 + r) * fp->inp_stride[1 + pl] + mbx * 8 + c);
            }
        }
    }
    IF_THREAD0 { w->pf_inp_tag = ok ? 1 + mby * fp->nmbx + mbx : 0; }
}
/* search window for macroblock (mbx, mby) around luma position (cx, cy) -> w->win, asynchronously */
HD void win_prefetch(const FrameParams *fp, MBWork *w, int mbx, int mby, int cx, int cy)
{
    const int stride = fp->stride[0];
    const int x0 = (cx - 24) & ~3, y0 = cy - 16;
    const int xmin = -16, xmax = fp->nmbx * 16 + 12, ymin = -16, ymax = fp->nmby * 16 + 15;
    const pix_t *plane = fp->ref[0];
    FOR_THREADS(i, (WIN_W / 4) * WIN_H)
    {
        int r = i >> 4, c4 = i & 15;
        int y = imin(imax(y0 + r, ymin), ymax), x = imin(imax(x0 + 4 * c4, xmin), xmax);
        cp_async4(&w->win[i], plane + y * stride + x);
    }
    IF_THREAD0 { w->pf_win_tag = 1 + mby * fp->nmbx + mbx; w->pf_win_x0 = x0; w->pf_win_y0 = y0; }
}
/* ---- bulk copy (TMA engine, cp.async.bulk) of a SAD-map record into shared memory, completion on an mbarrier ---- */
HD unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
HD void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
HD void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tMAP_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra MAP_DONE;\n\tbra MAP_WAIT;\n\tMAP_DONE:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
/* one thread: start the copy of macroblock (mbx, mby)'s record into buffer mbx & 1 */
HD void map_prefetch(const FrameParams *fp, MBWork *w, int mbx, int mby)
{
    IF_THREAD0
    {
        if (fp->use_sadmap && fp->slice_type == SLICE_P && mbx < fp->nmbx)
        {
            const int n = mby * fp->nmbx + mbx, b = mbx & 1;
            const unsigned bar = smem_u32(&w->map_bar[b]), dst = smem_u32(w->maps[b]);
            const uint32_t *src = fp->sadmap + (size_t)n * SM_WORDS;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      /* earlier reads of the buffer before the engine overwrites it */
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(SM_WORDS * 4)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"((unsigned)(SM_WORDS * 4)), "r"(bar) : "memory");
            w->map_cnt[b]++;
            w->map_tag[b] = 1 + n;
        }
    }
}
#endif

/* me_only: what the motion-estimation pre-pass needs of it (input samples, SAD maps, MV context) -- no neighbour samples
 * of the picture under construction, no Intra4x4 modes */
HDF_mb_load void mb_load(MBState &s, int me_only = 0)
{
    const FrameParams *fp = s.fp;
    MBWork *w = s.w;
    const int mbx = s.mbx, mby = s.mby;
    const int wv = fp->width, hv = fp->height;
    const int sy = fp->stride[0], sc = fp->stride[1];
    const pix_t *dy = fp->dec[0] + (mby * 16) * sy + mbx * 16;
    const pix_t *du = fp->dec[1] + (mby * 8) * sc + mbx * 8;
    const pix_t *dv = fp->dec[2] + (mby * 8) * sc + mbx * 8;
    const MBInfo *mbi = fp->mbi + mby * fp->nmbx + mbx;
    const int nmbx = fp->nmbx, av = s.avail;
    const int inside = (mbx + 1) * 16 <= wv && (mby + 1) * 16 <= hv;
    int have_inp = 0;
#if H264_DEVICE
    /* whatever was prefetched (input of this macroblock, its search window) has to have landed */
    cp_async_commit_wait_all();
    CTA_SYNC();
    have_inp = w->pf_inp_tag == 1 + mby * nmbx + mbx;
    {   /* SAD-map record: the staged copy when there is one for this macroblock, else the record where it lies */
        const uint32_t *rec = 0;
        if (fp->use_sadmap && fp->slice_type == SLICE_P)
        {
            const int b = mbx & 1;
            rec = fp->sadmap + (size_t)(mby * nmbx + mbx) * SM_WORDS;
            if (w->map_tag[b] == 1 + mby * nmbx + mbx) { mbar_wait(&w->map_bar[b], (unsigned)(w->map_cnt[b] - 1) & 1u); rec = w->maps[b]; }
        }
        lut_bind(s, rec);
    }
#else
    lut_bind(s, (fp->use_sadmap && fp->slice_type == SLICE_P) ? fp->sadmap + (size_t)(mby * nmbx + mbx) * SM_WORDS : 0);
#endif
    {
    /* one flat list of independent loads so that all of them are in flight together:
     * [0,96) neighbour samples / MVs / modes, [96,192) input words (unless prefetched).  A neighbour item first works
     * out WHERE its value lies (a chain of cases, no memory access), then there is one load per item: lanes that
     * take different cases still have their loads in flight together -- one memory round trip for the whole context
     * instead of one per case. */
    const int nb_lo = me_only ? 71 : 0, nb_n = me_only ? 14 : 96;
    FOR_THREADS(i, nb_n + 96)
    {
        if (i >= nb_n)
        {
            const int k = i - nb_n;
            uint32_t v;
            if (have_inp) v = w->pf_inp[k];
            else if (k < 64)
            {   /* input luma; samples beyond the visible picture replicate the last column / row
                 * (pix_copy_cropped_mb H:3536) */
                int r = k >> 2, c = (k & 3) * 4;
                if (inside) v = ld4u(fp->inp[0] + (mby * 16 + r) * fp->inp_stride[0] + mbx * 16 + c);
                else
                {
                    const pix_t *row = fp->inp[0] + imin(mby * 16 + r, hv - 1) * fp->inp_stride[0];
                    v = 0;
                    for (int q = 0; q < 4; q++) v |= (uint32_t)row[imin(mbx * 16 + c + q, wv - 1)] << (8 * q);
                }
            } else
            {
                int k2 = k - 64, r = k2 >> 2, q = k2 & 3, pl = q >> 1, c = (q & 1) * 4;
                if (inside) v = ld4u(fp->inp[1 + pl] + (mby * 8 + r) * fp->inp_stride[1 + pl] + mbx * 8 + c);
                else
                {
                    const pix_t *row = fp->inp[1 + pl] + imin(mby * 8 + r, hv / 2 - 1) * fp->inp_stride[1 + pl];
                    v = 0;
                    for (int q2 = 0; q2 < 4; q2++) v |= (uint32_t)row[imin(mbx * 8 + c + q2, wv / 2 - 1)] << (8 * q2);
                }
            }
            if (k < 64) *(uint32_t *)(w->inp_y + (k >> 2) * 16 + (k & 3) * 4) = v;
            else { int k2 = k - 64; *(uint32_t *)(w->inp_c + (k2 >> 2) * 16 + (k2 & 3) * 4) = v; }
        } else
        {
            /* unfiltered neighbour samples of the picture under construction (the reference's
             * top_line context, H:4693-4714), neighbours' MVs (enc->mv_pred, H:742) and I4x4 modes */
            const int j = nb_lo + i;
            const void *src = 0;       /* NULL: not available, the default value is stored */
            void *dst = 0;
            int kind = 0;              /* 0: sample (u8 -> u8), 1: vector (i32 -> i32), 2: mode (i8 -> i32) */
            int def = 0;
            if (j < 20) { dst = &w->top_y[j]; if ((av & AVAIL_T) && (j < 16 || (av & AVAIL_TR))) src = dy - sy + j; }
            else if (j < 36) { dst = &w->left_y[j - 20]; if (av & AVAIL_L) src = dy + (j - 20) * sy - 1; }
            else if (j < 44) { dst = &w->top_c[j - 36]; if (av & AVAIL_T) src = du - sc + (j - 36); }
            else if (j < 52) { dst = &w->top_c[8 + j - 44]; if (av & AVAIL_T) src = dv - sc + (j - 44); }
            else if (j < 60) { dst = &w->left_c[j - 52]; if (av & AVAIL_L) src = du + (j - 52) * sc - 1; }
            else if (j < 68) { dst = &w->left_c[8 + j - 60]; if (av & AVAIL_L) src = dv + (j - 60) * sc - 1; }
            else if (j == 68) { dst = &w->tl[0]; if (av & AVAIL_TL) src = dy - sy - 1; }
            else if (j == 69) { dst = &w->tl[1]; if (av & AVAIL_TL) src = du - sc - 1; }
            else if (j == 70) { dst = &w->tl[2]; if (av & AVAIL_TL) src = dv - sc - 1; }
            else if (j == 71) { w->ic[IC_STATE] = 0; w->ic[IC_COST0] = -1; w->task_next = 0; w->predc_tag = 0; }
            else if (j < 76) { kind = 1; def = MV_NA; dst = &w->mvp0_left[j - 72]; if (av & AVAIL_L) src = &mbi[-1].mv[4 * (j - 72) + 3]; }
            else if (j == 76) { kind = 1; def = MV_NA; dst = &w->mvp0_tl[0]; if (av & AVAIL_TL) src = &mbi[-nmbx - 1].mv[15]; }
            else if (j < 80) { kind = 1; def = MV_NA; dst = &w->mvp0_tl[j - 76]; if (av & AVAIL_L) src = &mbi[-1].mv[4 * (j - 77) + 3]; }
            else if (j < 84) { kind = 1; def = MV_NA; dst = &w->mvp0_top[j - 80]; if (av & AVAIL_T) src = &mbi[-nmbx].mv[12 + (j - 80)]; }
            else if (j == 84) { kind = 1; def = MV_NA; dst = &w->mvp0_top[4]; if (av & AVAIL_TR) src = &mbi[-nmbx + 1].mv[12]; }
            else if (j < 89) { kind = 2; def = -1; dst = &w->nb_i4mode[j - 85]; if (av & AVAIL_L) src = &mbi[-1].i4_mode[4 * (j - 85) + 3]; }
            else if (j < 93) { kind = 2; def = -1; dst = &w->nb_i4mode[4 + j - 89]; if (av & AVAIL_T) src = &mbi[-nmbx].i4_mode[12 + (j - 89)]; }
            int v = def;
            if (src) v = kind == 1 ? *(const int32_t *)src : (kind == 2 ? (int)*(const int8_t *)src : (int)*(const pix_t *)src);
            if (dst) { if (kind == 0) *(pix_t *)dst = (pix_t)v; else *(int32_t *)dst = v; }
        }
    }
    }
    CTA_SYNC();
#if H264_DEVICE
#if MB_WARPS == 4
    mb_prefetch_input(fp, w, mbx + 1, mby);      /* for the next macroblock of the row */
    if (w->pf_enable) map_prefetch(fp, w, mbx + 1, mby);
#endif
#endif
}

/* ------------------------------------------------------------------------------
 * a6: median / directional MV predictor on the rolling context
 * (me_mv_medianpredictor_get H:3720).  x,y,wd,ht in units of 4x4 blocks.
 * ---------------------------------------------------------------------------- */
HD int med3(int a, int b, int c) { return imax(imin(imax(a, b), c), imin(a, b)); }
HDF_mvp_get int mvp_get(const int32_t *left, const int32_t *tlv, const int32_t *top, int flag, int x, int y, int wd, int ht)
{
    int a = left[y], b = top[x], c = top[x + wd], d = tlv[y];
    if (!x)
    {
        if (!(flag & AVAIL_L)) a = MV_NA;
        if (!(flag & AVAIL_TL)) d = MV_NA;
    }
    if (!y)
    {
        if (!(flag & AVAIL_T))
        {
            b = MV_NA;
            if (x + wd < 4) c = MV_NA;
            if (x > 0) d = MV_NA;
        }
        if (!(flag & AVAIL_TL) && !x) d = MV_NA;
        if (!(flag & AVAIL_TR) && x + wd == 4) c = MV_NA;
    }
    if (x + wd == 4 && (!(flag & AVAIL_TR) || y)) c = d;
    int na = a != MV_NA, nb = b != MV_NA, nc = c != MV_NA;
    int kind = 0;   /* 0 median, 1 left, 2 up, 3 up-right */
    if (na && !nb && !nc) kind = 1;
    else if (!na && nb && !nc) kind = 2;
    else if (!na && !nb && nc) kind = 3;
    if (wd == 2 && ht == 4) { if (x == 0) { if (na) kind = 1; } else { if (nc) kind = 3; } }
    else if (wd == 4 && ht == 2) { if (y == 0) { if (nb) kind = 2; } else { if (na) kind = 1; } }
    switch (kind)
    {
    case 1: return na ? a : 0;
    case 2: return nb ? b : 0;
    case 3: return nc ? c : 0;
    default:
        if (!(nb || nc)) return na ? a : 0;
        if (!na) a = 0;
        if (!nb) b = 0;
        if (!nc) c = 0;
        return mv_pack(med3(mv_x(a), mv_x(b), mv_x(c)), med3(mv_y(a), mv_y(b), mv_y(c)));
    }
}

/* me_mv_medianpredictor_put H:3696 -- must be called by one lane, followed by WSYNC */
HD void mvp_put(SearchScratch *w, int x, int y, int wd, int ht, int mv)
{
    w->mvp_tl[y] = w->mvp_top[x + wd - 1];
    for (int i = 1; i < ht; i++) w->mvp_tl[y + i] = mv;
    for (int i = 0; i < ht; i++) w->mvp_left[y + i] = mv;
    for (int i = 0; i < wd; i++) w->mvp_top[x + i] = mv;
}

/* ------------------------------------------------------------------------------
 * Search window: a WIN_W x WIN_H copy of the reference luma around the 16x16 MV predictor,
 * staged once per macroblock in shared memory.  Every SAD / interpolation asks ref_at() for
 * its block; inside the window (the common case) the samples come from shared memory,
 * otherwise straight from the frame in global memory -- identical values either way.
 * ---------------------------------------------------------------------------- */
HDF_win_load void win_load(MBState &s, int cx, int cy)
{
    const FrameParams *fp = s.fp;
    const int stride = fp->stride[0];
    const int x0 = (cx - 24) & ~3, y0 = cy - 16;
#if H264_DEVICE
    /* a window prefetched for this macroblock is used as it is when it sits close enough to the
     * wanted one (any window is exact: ref_at() goes to the frame for what the window lacks) */
    if (s.w->pf_win_tag == 1 + s.mby * fp->nmbx + s.mbx && iabs(s.w->pf_win_x0 - x0) <= 8 && iabs(s.w->pf_win_y0 - y0) <= 6)
    {
        s.win_x0 = s.w->pf_win_x0; s.win_y0 = s.w->pf_win_y0; s.win_ok = 1;
        return;
    }
#endif
    const int xmin = -16, xmax = fp->nmbx * 16 + 12, ymin = -16, ymax = fp->nmby * 16 + 15;
    const pix_t *plane = fp->ref[0];
    uint32_t *win = s.w->win;
    FOR_THREADS(i, (WIN_W / 4) * WIN_H)
    {
        int r = i >> 4, c4 = i & 15;      /* WIN_W / 4 == 16 words per row */
        int y = imin(imax(y0 + r, ymin), ymax), x = imin(imax(x0 + 4 * c4, xmin), xmax);
        win[i] = *(const uint32_t *)(plane + y * stride + x);
    }
    s.win_x0 = x0; s.win_y0 = y0; s.win_ok = 1;
#if H264_DEVICE
    IF_THREAD0 { s.w->pf_win_tag = 0; }
#endif
    CTA_SYNC();
}

/* pointer to reference sample (bx,by) for a bw x bh block read with up to 3 samples of
 * filter margin; *stride receives the row pitch to use with it */
HD const pix_t *ref_at(const MBState &s, int bx, int by, int bw, int bh, int *stride)
{
    int lx = bx - s.win_x0, ly = by - s.win_y0;
    if (s.win_ok && lx >= 3 && ly >= 3 && lx + bw + 8 <= WIN_W && ly + bh + 4 <= WIN_H)
    {
        *stride = WIN_W;
        return (const pix_t *)s.w->win + ly * WIN_W + lx;
    }
    *stride = s.fp->stride[0];
    return s.fp->ref[0] + by * s.fp->stride[0] + bx;
}

/* ------------------------------------------------------------------------------
 * a4: integer-pel greedy diamond + diagonal step + 7-probe sub-pel refinement
 * (me_search_diamond H:4973-5176).
 *   ppx,ppy: partition offset inside the macroblock
 *   inp    : input MB + partition offset (stride 16)
 *   mv     : in/out absolute quarter-pel MV
 *   rng    : search rectangle x0,y0,x1,y1
 *   buf[4] : scratch, hpel, hpel1, hpel2 destinations (stride-16 blocks)
 * Returns the best cost; *pbest receives the buffer holding the best prediction.
 * ---------------------------------------------------------------------------- */
HDF_me_search int me_search(const MBState &s, int ppx, int ppy, const pix_t *inp, int *pmv, const int *rng,
                 int mv_pred, int min_sad, int bw, int bh, pix_t *const buf[4], pix_t **pbest)
{
    const FrameParams *fp = s.fp;
    const int lam = fp->lambda_mv_q4;
    int rs;
    const pix_t *rp;
    int mv = *pmv;
    uint32_t cache[8];
    int dir, cloop, dir_prev, cost, v;
    /* costs of the 8 integer neighbours of `nb_center`, evaluated together the first time
     * the replayed search asks for one of them */
    int nb[8], nb_center = 0, nb_valid = 0;

    for (;;)   /* "restart" loop */
    {
        dir = 0; cloop = 4; dir_prev = -1;
        for (int i = 0; i < 8; i++) cache[i] = 0xffffu;
        do
        {
            int dx = dir == 0 ? 4 : (dir == 1 ? -4 : 0), dy = dir == 2 ? 4 : (dir == 3 ? -4 : 0);
            v = mv_pack(mv_x(mv) + dx, mv_y(mv) + dy);
            if (mv_in_rect(v, rng[0], rng[1], rng[2], rng[3]) && cache[dir] == 0xffffu)
            {
                if (!nb_valid || nb_center != mv)
                {
                    rp = ref_at(s, (mv_x(mv) >> 2) + ppx - 1, (mv_y(mv) >> 2) + ppy - 1, bw + 2, bh + 2, &rs);
                    sad_nb8(rp + rs + 1, rs, inp, bw, bh, nb);
                    nb_center = mv; nb_valid = 1;
                }
                cost = nb[dir] + mv_cost(v, mv_pred, lam);
                cache[dir] = (uint32_t)cost & 0xffffu;
                if (cost < min_sad)
                {
                    uint32_t corner = 0xffffu;
                    if (dir_prev >= 0) corner = cache[4 + dir];
                    cache[4] = cache[0]; cache[5] = cache[1]; cache[6] = cache[2]; cache[7] = cache[3];
                    cache[0] = cache[1] = cache[2] = cache[3] = 0xffffu;
                    if (dir_prev >= 0) cache[dir_prev ^ 1] = corner;
                    cache[dir ^ 1] = (uint32_t)min_sad & 0xffffu;
                    dir_prev = dir;
                    dir--;
                    cloop = 4 + 1;
                    mv = v;
                    min_sad = cost;
                }
            }
            dir = (dir + 1) & 3;
        } while (--cloop);

        /* one diagonal probe towards the two cheaper axis neighbours */
        {
            int pdy = cache[3] >= cache[2] ? 4 : -4;
            int sdx = cache[1] >= cache[0] ? 4 : -4;
            v = mv_pack(mv_x(mv) + sdx, mv_y(mv) + pdy);
            if (mv_in_rect(v, rng[0], rng[1], rng[2], rng[3]))
            {
                if (!nb_valid || nb_center != mv)
                {
                    rp = ref_at(s, (mv_x(mv) >> 2) + ppx - 1, (mv_y(mv) >> 2) + ppy - 1, bw + 2, bh + 2, &rs);
                    sad_nb8(rp + rs + 1, rs, inp, bw, bh, nb);
                    nb_center = mv; nb_valid = 1;
                }
                cost = nb[4 + (sdx < 0 ? 1 : 0) + (pdy < 0 ? 2 : 0)] + mv_cost(v, mv_pred, lam);
                if (cost < min_sad) { mv = v; min_sad = cost; continue; }
            }
        }
        break;
    }

    if (bw == 16 && bh == 16) PROF_SUB(s, 12);
    rp = ref_at(s, (mv_x(mv) >> 2) + ppx - 1, (mv_y(mv) >> 2) + ppy - 1, bw + 2, bh + 2, &rs);
    rp += rs + 1;                       /* integer sample position of the block */
    copy_block(rp, rs, buf[0], bw, bh);
    pix_t *best = buf[0];

    if (fp->speed < 9 && mv_in_rect(mv, fp->mvlim_x0 + 16, fp->mvlim_y0 + 16, fp->mvlim_x1 - 16, fp->mvlim_y1 - 16))
    {
        /* 7 sub-pel probes around the integer optimum (H:5083-5174): three half-sample blocks,
         * four quarter-sample averages; costs evaluated together, decision replayed in order */
        uint32_t minsad1 = cache[1], minsad2 = cache[3];
        int sqx = -1, sqy = 0, pqx = 0, pqy = -1;       /* secondary / primary quarter steps */
        if (cache[3] >= cache[2]) { pqy = 1; minsad2 = cache[2]; }
        if (cache[1] >= cache[0]) { sqx = 1; minsad1 = cache[0]; }
        if (minsad2 > minsad1) { int t; t = sqx; sqx = pqx; pqx = t; t = sqy; sqy = pqy; pqy = t; }
        const int dgx = pqx + sqx, dgy = pqy + sqy;
        pix_t *I = buf[0], *C = buf[1], *H1 = buf[2], *H2 = buf[3];
        /* the three half-sample blocks mv + 2q come straight from the half-sample planes: pure half
         * positions, i.e. plane h (0,2), b (2,0) or j (2,2) at an integer offset; the three copies
         * are issued together (one memory round trip) */
        const int st = fp->stride[0];
        const long go = (long)((mv_y(mv) >> 2) + ppy) * st + (mv_x(mv) >> 2) + ppx;
        const pix_t *src3[3];
#pragma unroll
        for (int k = 0; k < 3; k++)
        {
            const int hx = 2 * (k == 0 ? pqx : (k == 1 ? sqx : dgx)), hy = 2 * (k == 0 ? pqy : (k == 1 ? sqy : dgy));
            const pix_t *pl = (hx & 3) ? ((hy & 3) ? fp->hp[2] : fp->hp[0]) : fp->hp[1];
            src3[k] = pl + go + (long)(hy >> 2) * st + (hx >> 2);
        }
        {
            const int sh = bw == 16 ? 2 : 1;
            FOR_LANES(i, bh << sh)
            {
                const int r = i >> sh, c = (i & ((1 << sh) - 1)) * 4;
                const uint32_t v1 = ld4u(src3[0] + r * st + c), v2 = ld4u(src3[1] + r * st + c), v3 = ld4u(src3[2] + r * st + c);
                *(uint32_t *)(H1 + r * 16 + c) = v1;
                *(uint32_t *)(H2 + r * 16 + c) = v2;
                *(uint32_t *)(C + r * 16 + c) = v3;
            }
        }
        WSYNC();
        if (bw == 16 && bh == 16) PROF_SUB(s, 13);
        int sq[7];
        sad_qpel7(I, H1, H2, C, inp, bw, bh, sq);
        int vbest = mv, ibest = -1;
#pragma unroll
        for (int i = 0; i < 7; i++)
        {
            int ox = i == 0 ? 2 * pqx : (i == 1 ? pqx : (i == 2 ? 2 * sqx : (i == 3 ? sqx : (i == 4 ? dgx : (i == 5 ? 2 * dgx : pqx + dgx)))));
            int oy = i == 0 ? 2 * pqy : (i == 1 ? pqy : (i == 2 ? 2 * sqy : (i == 3 ? sqy : (i == 4 ? dgy : (i == 5 ? 2 * dgy : pqy + dgy)))));
            v = mv_pack(mv_x(mv) + ox, mv_y(mv) + oy);
            int sad_test = sq[i] + mv_cost(v, mv_pred, lam);
            if (sad_test < min_sad) { min_sad = sad_test; vbest = v; ibest = i; }
        }
        mv = vbest;
        /* materialise the winning prediction (averages in place over an operand no longer needed) */
        if (ibest == 0) best = H1;
        else if (ibest == 2) best = H2;
        else if (ibest == 5) best = C;
        else if (ibest == 1) { average_block(I, H1, I, bw, bh); best = I; }
        else if (ibest == 3) { average_block(I, H2, I, bw, bh); best = I; }
        else if (ibest == 4) { average_block(H1, H2, H1, bw, bh); best = H1; }
        else if (ibest == 6) { average_block(C, H1, C, bw, bh); best = C; }
    }
    WSYNC();
    *pmv = mv;
    *pbest = best;
    return min_sad;
}

/* ------------------------------------------------------------------------------
 * Look-up flavour of the searches (h264_sadmap.h).  The record is bound by mb_load(); lut_decide() picks the flavour
 * for the whole macroblock.  lut_part_sad() is the one place that turns "SAD of partition (ppx, ppy, bw, bh) at absolute
 * quarter-sample vector v" into a number: from the maps when the position is tabulated, from the pictures otherwise
 * (integer position: the reference picture; sub-sample position: interp_luma_planes on the half-sample planes) --
 * the same value either way.
 * ---------------------------------------------------------------------------- */
HD int lut_decide(const MBState &s, int mvp16)
{
    if (!s.map || s.fp->slice_type != SLICE_P || s.map[2] != 1u) return 0;
    const int px = (mv_x(mvp16) + 2) >> 2, py = (mv_y(mvp16) + 2) >> 2;
    return iabs(px - s.map_cx) <= SM_R - 2 && iabs(py - s.map_cy) <= SM_R - 2;
}
HD int lut_quads(const MBState &s, int vrel, uint32_t *lo, uint32_t *hi)
{
    return sadmap_lookup(s.map, s.map_cx, s.map_cy, s.map_qcx, s.map_qcy, mv_x(vrel), mv_y(vrel), lo, hi);
}
HDN int part_sad_pixels(const MBState &s, int vabs, int ppx, int ppy, int bw, int bh)
{
    const FrameParams *fp = s.fp;
    const int st = fp->stride[0];
    const pix_t *inp = s.w->inp_y + ppy * 16 + ppx;
    const long o = (long)((mv_y(vabs) >> 2) + ppy) * st + (mv_x(vabs) >> 2) + ppx;
    if (!((mv_x(vabs) | mv_y(vabs)) & 3)) return sad_frame_wh(fp->ref[0] + o, st, inp, bw, bh);
    interp_luma_planes(fp->ref[0] + o, fp->hp[0] + o, fp->hp[1] + o, fp->hp[2] + o, st, mv_x(vabs) & 3, mv_y(vabs) & 3, bw, bh, s.ss->tmpblk);
    WSYNC();
    const int v = sad_sm_wh(s.ss->tmpblk, inp, bw, bh);
    WSYNC();
    return v;
}
#if !H264_DEVICE
extern long g_emu_lut[8];       /* emulation statistics: [0] P macroblocks, [1] of them in the look-up flavour, [2] look-ups, [3] misses */
#define LUT_STAT(k) (g_emu_lut[k]++)
#else
#define LUT_STAT(k) ((void)0)
#endif
HD int lut_part_sad(const MBState &s, int vabs, int ppx, int ppy, int bw, int bh)
{
    uint32_t lo, hi;
    LUT_STAT(2);
    if (lut_quads(s, mv_pack(mv_x(vabs) - s.mbx * 64, mv_y(vabs) - s.mby * 64), &lo, &hi)) return quads_part(lo, hi, ppx, ppy, bw, bh);
    LUT_STAT(3);
    return part_sad_pixels(s, vabs, ppx, ppy, bw, bh);
}

/* me_search() with look-ups: the same walk (H:4973-5176) -- greedy diamond with the 16-bit cost cache, one diagonal
 * probe + restart, seven ordered sub-sample probes -- but every SAD is one lut_part_sad() and no prediction block is built. */
HDN int me_search_lut(const MBState &s, int ppx, int ppy, int *pmv, const int *rng, int mv_pred, int min_sad, int bw, int bh)
{
    const FrameParams *fp = s.fp;
    const int lam = fp->lambda_mv_q4;
    int mv = *pmv;
    uint32_t c0, c1, c2, c3, p0, p1, p2, p3;      /* cache[0..3], cache[4..7] of the reference, kept in registers */
    int dir, cloop, dir_prev, cost, v;
    for (;;)
    {
        dir = 0; cloop = 4; dir_prev = -1;
        c0 = c1 = c2 = c3 = p0 = p1 = p2 = p3 = 0xffffu;
        do
        {
            const int dx = dir == 0 ? 4 : (dir == 1 ? -4 : 0), dy = dir == 2 ? 4 : (dir == 3 ? -4 : 0);
            v = mv_pack(mv_x(mv) + dx, mv_y(mv) + dy);
            const uint32_t cd = dir == 0 ? c0 : (dir == 1 ? c1 : (dir == 2 ? c2 : c3));
            if (mv_in_rect(v, rng[0], rng[1], rng[2], rng[3]) && cd == 0xffffu)
            {
                cost = lut_part_sad(s, v, ppx, ppy, bw, bh) + mv_cost(v, mv_pred, lam);
                const uint32_t cc = (uint32_t)cost & 0xffffu;
                if (dir == 0) c0 = cc; else if (dir == 1) c1 = cc; else if (dir == 2) c2 = cc; else c3 = cc;
                if (cost < min_sad)
                {
                    uint32_t corner = 0xffffu;
                    if (dir_prev >= 0) corner = dir == 0 ? p0 : (dir == 1 ? p1 : (dir == 2 ? p2 : p3));
                    p0 = c0; p1 = c1; p2 = c2; p3 = c3;
                    c0 = c1 = c2 = c3 = 0xffffu;
                    if (dir_prev >= 0) { const int k = dir_prev ^ 1; if (k == 0) c0 = corner; else if (k == 1) c1 = corner; else if (k == 2) c2 = corner; else c3 = corner; }
                    { const int k = dir ^ 1; const uint32_t m = (uint32_t)min_sad & 0xffffu; if (k == 0) c0 = m; else if (k == 1) c1 = m; else if (k == 2) c2 = m; else c3 = m; }
                    dir_prev = dir;
                    dir--;
                    cloop = 4 + 1;
                    mv = v;
                    min_sad = cost;
                }
            }
            dir = (dir + 1) & 3;
        } while (--cloop);
        {
            const int pdy = c3 >= c2 ? 4 : -4, sdx = c1 >= c0 ? 4 : -4;
            v = mv_pack(mv_x(mv) + sdx, mv_y(mv) + pdy);
            if (mv_in_rect(v, rng[0], rng[1], rng[2], rng[3]))
            {
                cost = lut_part_sad(s, v, ppx, ppy, bw, bh) + mv_cost(v, mv_pred, lam);
                if (cost < min_sad) { mv = v; min_sad = cost; continue; }
            }
        }
        break;
    }
    if (fp->speed < 9 && mv_in_rect(mv, fp->mvlim_x0 + 16, fp->mvlim_y0 + 16, fp->mvlim_x1 - 16, fp->mvlim_y1 - 16))
    {
        uint32_t minsad1 = c1, minsad2 = c3;
        int sqx = -1, sqy = 0, pqx = 0, pqy = -1;
        if (c3 >= c2) { pqy = 1; minsad2 = c2; }
        if (c1 >= c0) { sqx = 1; minsad1 = c0; }
        if (minsad2 > minsad1) { int t; t = sqx; sqx = pqx; pqx = t; t = sqy; sqy = pqy; pqy = t; }
        const int dgx = pqx + sqx, dgy = pqy + sqy;
        int vbest = mv;
#pragma unroll 1
        for (int i = 0; i < 7; i++)
        {
            const int ox = i == 0 ? 2 * pqx : (i == 1 ? pqx : (i == 2 ? 2 * sqx : (i == 3 ? sqx : (i == 4 ? dgx : (i == 5 ? 2 * dgx : pqx + dgx)))));
            const int oy = i == 0 ? 2 * pqy : (i == 1 ? pqy : (i == 2 ? 2 * sqy : (i == 3 ? sqy : (i == 4 ? dgy : (i == 5 ? 2 * dgy : pqy + dgy)))));
            v = mv_pack(mv_x(mv) + ox, mv_y(mv) + oy);
            const int sad_test = lut_part_sad(s, v, ppx, ppy, bw, bh) + mv_cost(v, mv_pred, lam);
            if (sad_test < min_sad) { min_sad = sad_test; vbest = v; }
        }
        mv = vbest;
    }
    *pmv = mv;
    return min_sad;
}


#if H264_DEVICE
/* ------------------------------------------------------------------------------
 * sm_100a execution of the look-up flavour: ONE position per lane.  The host emulation runs me_search_lut() /
 * the serial candidate loop above (which document the walk); here the eight neighbours of a search centre, the seven
 * sub-sample probes and the start candidates are each costed by their own lane (look-up + MV cost), and only the
 * order-dependent replay of the reference's comparisons runs warp-uniform on the gathered costs.
 * ---------------------------------------------------------------------------- */
/* SAD of the lane's own vector v (lanes with valid == 0 return 0).  Positions that are not tabulated are computed from
 * the pictures, eight of them at a time: four lanes per position, every lane a quarter of the block's words, each word
 * predicted where it lies (interp_luma_sel: the position table of H:2079-2130 over G and the half-sample planes).
 * The eight neighbours of a search centre and the seven sub-sample probes lie within two integer samples of each
 * other: the few rows of G, b, h, j they need (19 x 24 bytes per plane) are staged in the warp's scratch first, with
 * coalesced loads that are all in flight together, and the positions are evaluated from there.  (Gathering the words
 * of eight positions straight from the pictures makes every load instruction touch ~28 cache lines: for flat content,
 * whose vectors follow the predictors rather than the SAD minimum the quarter map is centred on, all seven probes are
 * untabulated and that gather was 30 % of such a macroblock's latency.) */
HD int lut_sad_lanes(const MBState &s, int v, int valid, int ppx, int ppy, int bw, int bh)
{
    const unsigned FULLM = 0xffffffffu;
    uint32_t lo = 0, hi = 0;
    int hit = 0;
    if (valid) hit = lut_quads(s, mv_pack(mv_x(v) - s.mbx * 64, mv_y(v) - s.mby * 64), &lo, &hi);
    int sad = hit ? quads_part(lo, hi, ppx, ppy, bw, bh) : 0;
    unsigned miss = __ballot_sync(FULLM, valid && !hit);
    while (miss)
    {
        const FrameParams *fp = s.fp;
        const int lane = LANE_ID, g = lane >> 2, sub = lane & 3;
        /* lane that asked for the g-th missing position (none: a dead group) */
        int src = -1;
        { unsigned m = miss; for (int k = 0; k < g; k++) m &= m - 1; if (m) src = __ffs((int)m) - 1; }
        const int vg = __shfl_sync(FULLM, v, src & 31);
        const int wsh = bw == 16 ? 2 : 1, nw = bh << wsh;
        const int st = fp->stride[0];
        /* do the positions of this batch fit one window? */
        const int bx = mv_x(vg) >> 2, by = mv_y(vg) >> 2;
        const int xmin = __reduce_min_sync(FULLM, src >= 0 ? bx : 0x7fff), xmax = __reduce_max_sync(FULLM, src >= 0 ? bx : -0x7fff);
        const int ymin = __reduce_min_sync(FULLM, src >= 0 ? by : 0x7fff), ymax = __reduce_max_sync(FULLM, src >= 0 ? by : -0x7fff);
        const pix_t *pl0 = fp->ref[0], *pl1 = fp->hp[0], *pl2 = fp->hp[1], *pl3 = fp->hp[2];
        int pst = st, ax = mv_x(vg) + ppx * 4, ay = mv_y(vg) + ppy * 4;
        if (xmax - xmin <= 2 && ymax - ymin <= 2)
        {
            const int wx = (xmin + ppx) & ~3, wy = ymin + ppy, nitem = (ymax - ymin + bh + 1) * (PROBE_PITCH / 4);
            uint32_t *pw = (uint32_t *)s.ss->store;
            uint32_t t[16];
#pragma unroll
            for (int p = 0; p < 4; p++)
            {
                const uint32_t *srcw = (const uint32_t *)((p == 0 ? pl0 : (p == 1 ? pl1 : (p == 2 ? pl2 : pl3))) + (long)wy * st + wx);
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    const int i = lane + 32 * u, r = i / (PROBE_PITCH / 4), c = i - r * (PROBE_PITCH / 4);
                    t[p * 4 + u] = i < nitem ? srcw[r * (st >> 2) + c] : 0u;
                }
            }
            WSYNC();                        /* earlier readers of the scratch */
#pragma unroll
            for (int p = 0; p < 4; p++)
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    const int i = lane + 32 * u;
                    if (i < PROBE_ROWS * (PROBE_PITCH / 4)) pw[p * (PROBE_ROWS * PROBE_PITCH / 4) + i] = t[p * 4 + u];
                }
            WSYNC();
            pl0 = (const pix_t *)pw; pl1 = pl0 + PROBE_ROWS * PROBE_PITCH; pl2 = pl1 + PROBE_ROWS * PROBE_PITCH; pl3 = pl2 + PROBE_ROWS * PROBE_PITCH;
            pst = PROBE_PITCH; ax -= 4 * wx; ay -= 4 * wy;
        }
        int part = 0;
        if (src >= 0)
        {
            const pix_t *inp = s.w->inp_y + ppy * 16 + ppx;
            /* a lane's share is nw / 4 = 4, 8 or 16 words: four at a time, their loads issued before the first use.
             * Which planes a position averages and how its words are aligned is the same for every word of the block:
             * worked out once, for word (0, 0) */
            const pix_t *a0, *b0;
            interp_luma_sel(pl0, pl1, pl2, pl3, pst, ax, ay, &a0, &b0);
            const uint32_t *qa = (const uint32_t *)((uintptr_t)a0 & ~(uintptr_t)3), *qb = (const uint32_t *)((uintptr_t)b0 & ~(uintptr_t)3);
            const unsigned sha = (unsigned)((uintptr_t)a0 & 3) * 8, shb = (unsigned)((uintptr_t)b0 & 3) * 8;
            const int pw4 = pst >> 2;
            for (int k0 = sub; k0 < nw; k0 += 16)
            {
                uint32_t va[4], vb[4];
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    const int k = k0 + 4 * u, r = k >> wsh, c = k & ((1 << wsh) - 1), o = r * pw4 + c;
                    va[u] = __funnelshift_r(qa[o], qa[o + 1], sha);
                    vb[u] = __funnelshift_r(qb[o], qb[o + 1], shb);
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    const int k = k0 + 4 * u, r = k >> wsh, c = k & ((1 << wsh) - 1);
                    part += sad4(avg4(va[u], vb[u]), ld4_sm(inp + r * 16 + 4 * c));
                }
            }
        }
        WSYNC();                            /* the scratch may be staged again */
        part += __shfl_xor_sync(FULLM, part, 1);
        part += __shfl_xor_sync(FULLM, part, 2);
        const int rank = __popc(miss & ((1u << lane) - 1u));      /* this lane's position is the rank-th missing one */
        const int got = __shfl_sync(FULLM, part, (rank & 7) * 4);
        if (((miss >> lane) & 1u) && rank < 8) sad = got;
        for (int k = 0; k < 8 && miss; k++) miss &= miss - 1;
    }
    return sad;
}

HDN int me_search_par(const MBState &s, int ppx, int ppy, int *pmv, const int *rng, int mv_pred, int min_sad, int bw, int bh)
{
    const unsigned FULLM = 0xffffffffu;
    const FrameParams *fp = s.fp;
    const int lam = fp->lambda_mv_q4, lane = LANE_ID;
    int mv = *pmv;
    /* neighbour k = lane & 7 of a centre: 0 (+1,0) 1 (-1,0) 2 (0,+1) 3 (0,-1) 4 (+1,+1) 5 (-1,+1) 6 (+1,-1) 7 (-1,-1) */
    const int k = lane & 7;
    const int ox = (k == 0 || k == 4 || k == 6) ? 4 : ((k == 2 || k == 3) ? 0 : -4);
    const int oy = (k == 2 || k == 4 || k == 5) ? 4 : ((k == 0 || k == 1) ? 0 : -4);
    /* the reference's cost cache: cache[0..3] (this centre) and cache[4..7] (the previous one), 16 bits each, packed */
    unsigned long long cc, pc;
#define CGET(v, d) ((uint32_t)((v) >> (16 * (d))) & 0xffffu)
#define CSET(v, d, x) (v) = ((v) & ~(0xffffULL << (16 * (d)))) | ((unsigned long long)((x) & 0xffffu) << (16 * (d)))
    int dir, cloop, dir_prev, cst;
    unsigned inmask;
    for (;;)
    {
        dir = 0; cloop = 4; dir_prev = -1;
        cc = pc = ~0ULL;
        for (;;)
        {
            const int v = mv_pack(mv_x(mv) + ox, mv_y(mv) + oy);
            const int inr = mv_in_rect(v, rng[0], rng[1], rng[2], rng[3]);
            cst = lut_sad_lanes(s, v, lane < 8 && inr, ppx, ppy, bw, bh) + mv_cost(v, mv_pred, lam);
            inmask = __ballot_sync(FULLM, inr) & 0xffu;
            int moved = 0;
            do
            {
                if (((inmask >> dir) & 1u) && CGET(cc, dir) == 0xffffu)
                {
                    const int cost = __shfl_sync(FULLM, cst, dir);
                    CSET(cc, dir, (uint32_t)cost);
                    if (cost < min_sad)
                    {
                        const uint32_t corner = dir_prev >= 0 ? CGET(pc, dir) : 0xffffu;
                        pc = cc;
                        cc = ~0ULL;
                        if (dir_prev >= 0) CSET(cc, dir_prev ^ 1, corner);
                        CSET(cc, dir ^ 1, (uint32_t)min_sad);
                        mv = mv_pack(mv_x(mv) + (dir == 0 ? 4 : (dir == 1 ? -4 : 0)), mv_y(mv) + (dir == 2 ? 4 : (dir == 3 ? -4 : 0)));
                        min_sad = cost;
                        dir_prev = dir;
                        dir--;
                        cloop = 4 + 1;
                        moved = 1;
                    }
                }
                dir = (dir + 1) & 3;
            } while (--cloop && !moved);
            if (!moved) break;          /* the walk has settled on this centre; after a move its neighbours are costed anew */
        }
        {
            const int pneg = CGET(cc, 3) >= CGET(cc, 2) ? 0 : 1, sneg = CGET(cc, 1) >= CGET(cc, 0) ? 0 : 1, kk = 4 + sneg + 2 * pneg;
            if ((inmask >> kk) & 1u)
            {
                const int cost = __shfl_sync(FULLM, cst, kk);
                if (cost < min_sad) { mv = mv_pack(mv_x(mv) + (sneg ? -4 : 4), mv_y(mv) + (pneg ? -4 : 4)); min_sad = cost; continue; }
            }
        }
        break;
    }
    const uint32_t c0 = CGET(cc, 0), c1 = CGET(cc, 1), c2 = CGET(cc, 2), c3 = CGET(cc, 3);
    if (bw == 16 && bh == 16) PROF_SUB(s, 12);
#undef CGET
#undef CSET
    if (fp->speed < 9 && mv_in_rect(mv, fp->mvlim_x0 + 16, fp->mvlim_y0 + 16, fp->mvlim_x1 - 16, fp->mvlim_y1 - 16))
    {
        uint32_t minsad1 = c1, minsad2 = c3;
        int sqx = -1, sqy = 0, pqx = 0, pqy = -1;
        if (c3 >= c2) { pqy = 1; minsad2 = c2; }
        if (c1 >= c0) { sqx = 1; minsad1 = c0; }
        if (minsad2 > minsad1) { int t; t = sqx; sqx = pqx; pqx = t; t = sqy; sqy = pqy; pqy = t; }
        const int dgx = pqx + sqx, dgy = pqy + sqy;
        /* probe i = lane: 0 2p, 1 p, 2 2s, 3 s, 4 p+s, 5 2(p+s), 6 2p+s (H:5119-5161); strict '<' in that order == the
         * smallest (cost, i) */
        const int i = lane & 7;
        const int qx = i == 0 ? 2 * pqx : (i == 1 ? pqx : (i == 2 ? 2 * sqx : (i == 3 ? sqx : (i == 4 ? dgx : (i == 5 ? 2 * dgx : pqx + dgx)))));
        const int qy = i == 0 ? 2 * pqy : (i == 1 ? pqy : (i == 2 ? 2 * sqy : (i == 3 ? sqy : (i == 4 ? dgy : (i == 5 ? 2 * dgy : pqy + dgy)))));
        const int v = mv_pack(mv_x(mv) + qx, mv_y(mv) + qy);
        const int cst = lut_sad_lanes(s, v, lane < 7, ppx, ppy, bw, bh) + mv_cost(v, mv_pred, lam);
        int key = lane < 7 ? ((cst << 3) | i) : 0x7FFFFFFF;
        key = min(key, __shfl_xor_sync(FULLM, key, 4));
        key = min(key, __shfl_xor_sync(FULLM, key, 2));
        key = min(key, __shfl_xor_sync(FULLM, key, 1));
        key = __shfl_sync(FULLM, key, 0);
        if ((key >> 3) < min_sad) { min_sad = key >> 3; mv = __shfl_sync(FULLM, v, key & 7); }
    }
    if (bw == 16 && bh == 16) PROF_SUB(s, 13);
    *pmv = mv;
    return min_sad;
}
#endif

/* rows 8 * half .. 8 * half + 7 of the luma prediction of an inter macroblock, made from its final vectors (one warp):
 * what the reference has in its buffers after the winning probe (H:5116-5170) or re-interpolates (H:5520) */
HDN void luma_pred_half(const MBState &s, int half, int type, const int32_t *pmv, pix_t *dst)
{
    const FrameParams *fp = s.fp;
    FOR_LANES(i, 32)
    {
        const int r = 8 * half + (i >> 2), c = (i & 3) * 4;
        const int part = type <= 0 ? 0 : (type == 1 ? (r >> 3) : (type == 2 ? (c >> 3) : (r >> 3) * 2 + (c >> 3)));
        const int mv = pmv[part];
        *(uint32_t *)(dst + r * 16 + c) = interp_luma_word(fp, mv_x(mv) + (s.mbx * 16 + c) * 4, mv_y(mv) + (s.mby * 16 + r) * 4);
    }
    WSYNC();
}

/* me_mv_set_range H:5181 */
HD void me_set_range(const FrameParams *fp, int *pnt, int *rng, int mby_q)
{
    int x0 = fp->mvlim_x0, x1 = fp->mvlim_x1;
    int y0 = (int16_t)imax(fp->mvlim_y0, mby_q - 63 * 4), y1 = (int16_t)imin(fp->mvlim_y1, mby_q + 63 * 4);
    int px = imin(imax(mv_x(*pnt), x0), x1), py = imin(imax(mv_y(*pnt), y0), y1);
    *pnt = mv_pack(px, py);
    rng[0] = imin(imax(px - MV_RANGE_PX * 4, x0), x1);
    rng[1] = imin(imax(py - MV_RANGE_PX * 4, y0), y1);
    rng[2] = imin(imax(px + MV_RANGE_PX * 4, x0), x1);
    rng[3] = imin(imax(py + MV_RANGE_PX * 4, y0), y1);
}

/* mb_inter_partition H:5224: which partition shapes are worth a search */
HD void inter_partition_hint(const int sad[4], int mode[4])
{
    int p00 = sad[0], p01 = sad[1], p10 = sad[2], p11 = sad[3];
    int sum = p00 + p01 + p10 + p11;
    int slope = iabs((p00 - p10) + (p01 - p11)) - iabs((p00 - p01) + (p10 - p11));
    int skew = iabs(p11 - p00) - iabs(p10 - p01);
    if (slope > (sum >> 4)) mode[1] = 1;
    if (slope < -(sum >> 4)) mode[2] = 1;
    if (iabs(skew) > (sum >> 4) && iabs(slope) <= (sum >> 4)) mode[3] = 1;
}

/* does macroblock (mbx, mby) cross the picture edge (cropped sizes)? */
HD int mb_crosses_edge(const FrameParams *fp, int mbx, int mby) { return (mbx + 1) * 16 > fp->width || (mby + 1) * 16 > fp->height; }

/* chroma motion compensation of plane pl for every partition of the current MB type
 * (interpolate_chroma H:4915). mvs: per-partition MVs relative to the MB.  Warp-level. */
HDF_mc_chroma_plane void mc_chroma_plane(const MBState &s, int pl, int type, const int32_t *mvs)
{
    const FrameParams *fp = s.fp;
    int bw = (type & 2) ? 4 : 8, bh = (type & 1) ? 4 : 8;
    if (type == MBT_SKIP) bw = bh = 8;
    int sc = fp->stride[1];
    int part = 0, x = 0, y = 0;
    for (;; part++)
    {
        int ax = mv_x(mvs[part]) + s.mbx * 64, ay = mv_y(mvs[part]) + s.mby * 64;
        const pix_t *ref = fp->ref[1 + pl] + ((ay >> 3) + y) * sc + (ax >> 3) + x;
        interp_chroma_block(ref, sc, ax & 7, ay & 7, bw, bh, s.w->predc + pl * 8 + 16 * y + x);
        x = (x + bw) & 7;
        if (!x) { y = (y + bh) & 7; if (!y) break; }
    }
    WSYNC();
}

/* ------------------------------------------------------------------------------
 * a5: inter mode decision (inter_choose_mode H:5283-5524), split into the candidate stage
 * (skip test + start candidates, one warp) and one search task per partition mode.
 * ---------------------------------------------------------------------------- */
/* Candidate stage.  Needs the window loaded around mvp16.  Publishes w->ic[] (lane 0). */
/* Candidate stage.  Needs the window loaded around mvp16.  Publishes w->ic[] (lane 0). */
HDF_inter_stage_a void inter_stage_a(MBState &s, const int32_t cl[2])
{
    const FrameParams *fp = s.fp;
    MBWork *w = s.w;
    const int mbqx = s.mbx * 64, mbqy = s.mby * 64;
    int pref[4] = {1, 0, 0, 0};
    int cand[12], ncand = 0, j = 0;
    int sad4v[4];
    int sad_skip = 0x7FFFFFFF, sad_best = 0x7FFFFFFF, cand_cost_best = 0;
    int mv_best = mv_pack(MV_NA, 0);
    int state = 2;

    /* skip predictor (me_mv_medianpredictor_get_skip H:3877) */
    const int mvp16 = mvp_get(w->mvp0_left, w->mvp0_tl, w->mvp0_top, s.avail, 0, 0, 4, 4);
    int mv_skip = 0;
    if (!(~s.avail & (AVAIL_L | AVAIL_T)) && w->mvp0_left[0] != 0 && w->mvp0_top[0] != 0) mv_skip = mvp16;
    const int mv_skip_a = mv_pack(mv_x(mv_skip) + mbqx, mv_y(mv_skip) + mbqy);

    if (mv_in_rect(mv_skip_a, fp->mvlim_x0 + 16, fp->mvlim_y0 + 16, fp->mvlim_x1 - 16, fp->mvlim_y1 - 16))
    {
        uint32_t qlo, qhi;
        if (s.lut && lut_quads(s, mv_skip, &qlo, &qhi))
        {   /* the skip position is tabulated: no prediction block is needed for the test (encode_mb makes it if the
             * macroblock ends up using it) */
            sad4v[0] = (int)(qlo & 0xFFFF); sad4v[1] = (int)(qlo >> 16); sad4v[2] = (int)(qhi & 0xFFFF); sad4v[3] = (int)(qhi >> 16);
            sad_skip = sad4v[0] + sad4v[1] + sad4v[2] + sad4v[3];
        } else
        {
        if (!((mv_x(mv_skip_a) | mv_y(mv_skip_a)) & 3))
        {   /* full-pel vector: the prediction is a copy of integer samples, usually inside the search window */
            int rs;
            const pix_t *rp = ref_at(s, mv_x(mv_skip_a) >> 2, mv_y(mv_skip_a) >> 2, 16, 16, &rs);
            copy_block(rp, rs, w->skip_pred, 16, 16);
        } else
        {
            const int st = fp->stride[0];
            const long o = (long)(mv_y(mv_skip_a) >> 2) * st + (mv_x(mv_skip_a) >> 2);
            interp_luma_planes(fp->ref[0] + o, fp->hp[0] + o, fp->hp[1] + o, fp->hp[2] + o, st,
                               mv_x(mv_skip_a) & 3, mv_y(mv_skip_a) & 3, 16, 16, w->skip_pred);
        }
        WSYNC();
        sad_skip = sad_mb_quad(w->inp_y, 16, w->skip_pred, sad4v);
        }
        if (imax(imax(sad4v[0], sad4v[1]), imax(sad4v[2], sad4v[3])) < fp->skip_thr_inter)
        {
            int32_t one_mv = mv_skip;
            int ok = 1;
            /* A macroblock that crosses the picture edge: the reference copies its replicated chroma input (8 x 8, stride 8)
             * to the START of mb_pix_store (H:5333) -- where, after the swap of H:5316, the chroma prediction it is about
             * to be compared with lies (stride 16, U | V) -- so rows 0..3 of the "prediction" are input rows 2r (U) /
             * 2r + 1 (V) of the plane under test.  Reproduced, not fixed.  Only the TEST sees the damage: the chroma
             * prediction of an accepted skip is interpolated again before it becomes the reconstruction (H:5786). */
            const int crosses = mb_crosses_edge(fp, s.mbx, s.mby);
            for (int pl = 0; pl < 2 && ok; pl++)
            {
                mc_chroma_plane(s, pl, MBT_SKIP, &one_mv);
                int acc = 0;
                FOR_LANES(i, 16)
                {
                    int r = i >> 1, c = (i & 1) * 4;
                    const uint32_t pr = (crosses && r < 4) ? ld4_sm(w->inp_c + (2 * r + pl) * 16 + pl * 8 + c) : ld4_sm(w->predc + r * 16 + pl * 8 + c);
                    acc += sad4(ld4_sm(w->inp_c + r * 16 + pl * 8 + c), pr);
                }
                acc = wsum(acc);
                if (acc >= fp->skip_thr_inter) ok = 0;
            }
            if (ok) state = 1;
        }
        if (state != 1)
        {
            if (fp->speed < 1) inter_partition_hint(sad4v, pref);
            mv_best = cand[ncand++] = mv_round_fullpel(mv_skip);
            if (!((mv_x(mv_skip) | mv_y(mv_skip)) & 3))
            {
                sad_best = sad_skip;
                cand_cost_best = mv_cost(mv_skip, mvp16, fp->lambda_mv_q4);
                j = 1;
            }
        }
    }

    PROF_SUB(s, 3);
    if (state != 1)
    {
#if H264_DEVICE
        /* candidate start points (H:5370-5386), rounded to full-pel, duplicates dropped keeping first
         * occurrences (H:5198): one candidate per lane, in the reference's order; the survivors are
         * visited in lane order */
        const unsigned FULLM = 0xffffffffu;
        const int lane = LANE_ID;
        int cval = 0, cok = 0;
        if (lane == 0) { cval = mv_skip; cok = ncand; }            /* present when the skip vector was in range */
        else if (lane == 1) { cval = mvp16; cok = 1; }
        else if (lane == 2) { cval = 0; cok = 1; }
        else if (lane == 3) { cval = w->mvp0_left[0]; cok = (s.avail & AVAIL_L) && cval != MV_NA; }
        else if (lane == 4) { cval = w->mvp0_top[0]; cok = (s.avail & AVAIL_T) && cval != MV_NA; }
        else if (lane == 5) { cval = w->mvp0_top[4]; cok = (s.avail & AVAIL_TR) && cval != MV_NA; }
        else if (lane == 6) { cval = mv_pack(8 * 4, 0); cok = s.mbx <= 0; }
        else if (lane == 7) { cval = mv_pack(0, 8 * 4); cok = s.mby <= 0; }
        else if (lane == 8) { cval = cl[0]; cok = 1; }
        else if (lane == 9) { cval = cl[1]; cok = 1; }
        cval = mv_round_fullpel(cval);
        const unsigned okm = __ballot_sync(FULLM, cok);
        const unsigned same = __match_any_sync(FULLM, cval) & okm;
        unsigned keep = __ballot_sync(FULLM, cok && (same & (0u - same)) == (1u << lane));
        const unsigned keep_all = keep;
        if (j) keep &= keep - 1;                                    /* full-pel skip vector: its SAD is reused (H:5361) */
        PROF_SUB(s, 8);
        if (s.lut)
        {
            /* look-up flavour: every surviving candidate is costed by its own lane.  A full-pel skip vector is simply
             * costed again on lane 0 (same position, same numbers as the reused ones); the reference's "first strictly
             * smaller sum in list order" is the smallest (sum, lane). */
            const int mva = mv_pack(mv_x(cval) + mbqx, mv_y(cval) + mbqy);
            const int inr = ((keep_all >> lane) & 1u) && mv_in_rect(mva, fp->mvlim_x0, fp->mvlim_y0, fp->mvlim_x1, fp->mvlim_y1);
            uint32_t qlo = 0, qhi = 0;
            const int hit = inr && lut_quads(s, cval, &qlo, &qhi);
            unsigned miss = __ballot_sync(FULLM, inr && !hit);
            while (miss)
            {   /* not tabulated: from the reference picture, by the whole warp */
                const int l = __ffs((int)miss) - 1;
                miss &= miss - 1;
                const int ma = __shfl_sync(FULLM, mva, l);
                int q4[4];
                sad_mb_quad(fp->ref[0] + (long)(mv_y(ma) >> 2) * fp->stride[0] + (mv_x(ma) >> 2), fp->stride[0], w->inp_y, q4);
                if (lane == l) { qlo = (uint32_t)q4[0] | ((uint32_t)q4[1] << 16); qhi = (uint32_t)q4[2] | ((uint32_t)q4[3] << 16); }
            }
            int q4[4], pl[4] = {0, 0, 0, 0};
            q4[0] = (int)(qlo & 0xFFFF); q4[1] = (int)(qlo >> 16); q4[2] = (int)(qhi & 0xFFFF); q4[3] = (int)(qhi >> 16);
            const int sad = q4[0] + q4[1] + q4[2] + q4[3], cc = mv_cost(cval, mvp16, fp->lambda_mv_q4);
            if (inr && fp->speed < 1) inter_partition_hint(q4, pl);
            if (__ballot_sync(FULLM, pl[1])) pref[1] = 1;
            if (__ballot_sync(FULLM, pl[2])) pref[2] = 1;
            if (__ballot_sync(FULLM, pl[3])) pref[3] = 1;
            int key = inr ? (((sad + cc) << 5) | lane) : 0x7FFFFFFF;
#pragma unroll
            for (int o = 16; o; o >>= 1) key = min(key, __shfl_xor_sync(FULLM, key, o));
            if (key != 0x7FFFFFFF)
            {
                const int wl = key & 31;
                mv_best = __shfl_sync(FULLM, cval, wl); sad_best = __shfl_sync(FULLM, sad, wl); cand_cost_best = __shfl_sync(FULLM, cc, wl);
            }
            keep = 0;
        }
        while (keep)
        {
            const int cj = __shfl_sync(FULLM, cval, __ffs((int)keep) - 1);
            keep &= keep - 1;
            int mva = mv_pack(mv_x(cj) + mbqx, mv_y(cj) + mbqy);
            if (mv_in_rect(mva, fp->mvlim_x0, fp->mvlim_y0, fp->mvlim_x1, fp->mvlim_y1))
            {
                int cc = mv_cost(cj, mvp16, fp->lambda_mv_q4);
                int sad;
                uint32_t qlo, qhi;
                if (s.lut && lut_quads(s, cj, &qlo, &qhi))
                {
                    sad4v[0] = (int)(qlo & 0xFFFF); sad4v[1] = (int)(qlo >> 16); sad4v[2] = (int)(qhi & 0xFFFF); sad4v[3] = (int)(qhi >> 16);
                    sad = sad4v[0] + sad4v[1] + sad4v[2] + sad4v[3];
                } else
                {
                    int rs;
                    const pix_t *rp = ref_at(s, mv_x(mva) >> 2, mv_y(mva) >> 2, 16, 16, &rs);
                    sad = sad_mb_quad(rp, rs, w->inp_y, sad4v);
                }
                if (fp->speed < 1) inter_partition_hint(sad4v, pref);
                if (sad + cc < sad_best + cand_cost_best) { cand_cost_best = cc; sad_best = sad; mv_best = cj; }
            }
        }
#else
        /* candidate start points (H:5370-5386) */
        cand[ncand++] = mvp16;
        cand[ncand++] = 0;
        if ((s.avail & AVAIL_L) && w->mvp0_left[0] != MV_NA) cand[ncand++] = w->mvp0_left[0];
        if ((s.avail & AVAIL_T) && w->mvp0_top[0] != MV_NA) cand[ncand++] = w->mvp0_top[0];
        if ((s.avail & AVAIL_TR) && w->mvp0_top[4] != MV_NA) cand[ncand++] = w->mvp0_top[4];
        if (s.mbx <= 0) cand[ncand++] = mv_pack(8 * 4, 0);
        if (s.mby <= 0) cand[ncand++] = mv_pack(0, 8 * 4);
        cand[ncand++] = cl[0];
        cand[ncand++] = cl[1];
        {   /* round to full-pel and drop duplicates, keeping first occurrences (H:5198) */
            int k = 1;
            cand[0] = mv_round_fullpel(cand[0]);
            for (int a = 1; a < ncand; a++)
            {
                int m = mv_round_fullpel(cand[a]), i;
                for (i = 0; i < k; i++) if (m == cand[i]) break;
                if (i == k) cand[k++] = m;
            }
            ncand = k;
        }
        PROF_SUB(s, 8);
#pragma unroll 1
        for (; j < ncand; j++)
        {
            int mva = mv_pack(mv_x(cand[j]) + mbqx, mv_y(cand[j]) + mbqy);
            if (mv_in_rect(mva, fp->mvlim_x0, fp->mvlim_y0, fp->mvlim_x1, fp->mvlim_y1))
            {
                int cc = mv_cost(cand[j], mvp16, fp->lambda_mv_q4);
                int sad;
                uint32_t qlo, qhi;
                if (s.lut && lut_quads(s, cand[j], &qlo, &qhi))
                {
                    sad4v[0] = (int)(qlo & 0xFFFF); sad4v[1] = (int)(qlo >> 16); sad4v[2] = (int)(qhi & 0xFFFF); sad4v[3] = (int)(qhi >> 16);
                    sad = sad4v[0] + sad4v[1] + sad4v[2] + sad4v[3];
                } else
                {
                    int rs;
                    const pix_t *rp = ref_at(s, mv_x(mva) >> 2, mv_y(mva) >> 2, 16, 16, &rs);
                    sad = sad_mb_quad(rp, rs, w->inp_y, sad4v);
                }
                if (fp->speed < 1) inter_partition_hint(sad4v, pref);
                if (sad + cc < sad_best + cand_cost_best) { cand_cost_best = cc; sad_best = sad; mv_best = cand[j]; }
            }
        }
#endif
        PROF_SUB(s, 9);
    }
    IF_LANE0
    {
        w->ic[IC_PREF] = pref[0] | (pref[1] << 1) | (pref[2] << 2) | (pref[3] << 3);
        w->ic[IC_MV_BEST] = mv_best;
        w->ic[IC_SIG] = mv_best; w->ic[IC_SIG + 1] = sad_best; w->ic[IC_SIG + 2] = cand_cost_best;
        w->ic[IC_SIG + 3] = pref[1] | (pref[2] << 1) | (pref[3] << 2);
        w->ic[IC_SAD_BEST] = state == 1 ? 0 : sad_best + mv_cost(mv_best, mvp16, fp->lambda_mv_q4);
        w->ic[IC_MVP16] = mvp16;
        w->ic[IC_MV_SKIP] = mv_skip;
        w->ic[IC_SAD_SKIP] = sad_skip;
#if H264_DEVICE
        __threadfence_block();
        *(volatile int32_t *)&w->ic[IC_STATE] = state;
#else
        w->ic[IC_STATE] = state;
#endif
    }
    WSYNC();
}

/* Search of one partition mode (one warp, private scratch s.ss).  The prediction of the
 * whole macroblock for this mode ends up at w->mode_pred[mb_type]. */
HDF_inter_mode_search void inter_mode_search(MBState &s, int mb_type)
{
    const FrameParams *fp = s.fp;
    MBWork *w = s.w;
    SearchScratch *ss = s.ss;
    const int mbqx = s.mbx * 64, mbqy = s.mby * 64;
    const int mv_best = w->ic[IC_MV_BEST];
    int sad_best = w->ic[IC_SAD_BEST];
    const int nbits = mb_type == 0 ? 1 : (mb_type == 3 ? 12 : 4);
    int imv = 0;
    int part_sad = (nbits * fp->lambda_q4) >> 4;
    const int bw = (mb_type & 2) ? 8 : 16, bh = (mb_type & 1) ? 8 : 16;
    /* scratch tiles of the diamond / sub-pel search, and where this mode's MB prediction is assembled */
    pix_t *store = ss->store[0];
    pix_t *assembled = mb_type ? w->mode_store[mb_type - 1] : ss->store[2];
    pix_t *result = assembled;
    FOR_LANES(i, 13)
    {
        if (i < 4) ss->mvp_left[i] = w->mvp0_left[i];
        else if (i < 8) ss->mvp_tl[i - 4] = w->mvp0_tl[i - 4];
        else ss->mvp_top[i - 8] = w->mvp0_top[i - 8];
    }
    WSYNC();
    int px = 0, py = 0;
    for (;;)
    {
        int rng[4];
        int mvabs = mv_pack(mv_x(mv_best) + mbqx, mv_y(mv_best) + mbqy);
        me_set_range(fp, &mvabs, rng, mbqy + py * 4);
        int mv_pred = mvp_get(ss->mvp_left, ss->mvp_tl, ss->mvp_top, s.avail, px >> 2, py >> 2, bw >> 2, bh >> 2);
        int mv_pred_a = mv_pack(mv_x(mv_pred) + mbqx, mv_y(mv_pred) + mbqy);
        if (mb_type)
        {
            mvabs = mv_round_fullpel(mv_pred_a);
            me_set_range(fp, &mvabs, rng, mbqy + py * 4);
            if (s.lut) sad_best = lut_part_sad(s, mvabs, px, py, bw, bh) + mv_cost(mvabs, mv_pred_a, fp->lambda_mv_q4);
            else
            {
            int rs;
            const pix_t *rp = ref_at(s, (mv_x(mvabs) >> 2) + px, (mv_y(mvabs) >> 2) + py, bw, bh, &rs);
            sad_best = sad_frame_wh(rp, rs, w->inp_y + py * 16 + px, bw, bh)
                     + mv_cost(mvabs, mv_pred_a, fp->lambda_mv_q4);
            }
        }
        int sb = mb_type ? (mb_type == 2 ? 8 : 128) : 256;
        pix_t *bufs[4];
        bufs[0] = store; bufs[1] = store + sb; bufs[2] = store + (sb == 8 ? 256 : 2 * sb); bufs[3] = bufs[2] + sb;
        pix_t *dout = store;
#if H264_DEVICE
        if (s.lut) part_sad += me_search_par(s, px, py, &mvabs, rng, mv_pred_a, sad_best, bw, bh);
#else
        if (s.lut) part_sad += me_search_lut(s, px, py, &mvabs, rng, mv_pred_a, sad_best, bw, bh);
#endif
        else
        part_sad += me_search(s, px, py, w->inp_y + py * 16 + px, &mvabs, rng, mv_pred_a, sad_best,
                              bw, bh, bufs, &dout);
        if (s.lut) { }                  /* no prediction blocks in the look-up flavour (luma_pred_half makes the winner's) */
        else if (!mb_type) result = dout;
        else
        {
            const int sh = bw == 16 ? 2 : 1;
            FOR_LANES(i, bh << sh)
            {
                int r = i >> sh, c = (i & ((1 << sh) - 1)) * 4;
                *(uint32_t *)(assembled + (py + r) * 16 + px + c) = ld4_sm(dout + r * 16 + c);
            }
        }
        int mv = mv_pack(mv_x(mvabs) - mbqx, mv_y(mvabs) - mbqy);
        IF_LANE0
        {
            w->part_mvd[mb_type][imv] = mv_sub2(mv, mv_pred);
            w->part_mv[mb_type][imv] = mv;
            mvp_put(ss, px >> 2, py >> 2, bw >> 2, bh >> 2, mv);
        }
        imv++;
        WSYNC();
        px = (px + bw) & 15;
        if (!px) { py = (py + bh) & 15; if (!py) break; }
        if (mb_type && part_sad >= inter_cost_bound(fp, w)) break;      /* this mode has lost (exact pruning, see above) */
    }
    IF_LANE0
    {
#if H264_DEVICE
        __threadfence_block();
#endif
        if (!mb_type) *(volatile int32_t *)&w->ic[IC_COST0] = part_sad;
        w->mode_cost[mb_type] = part_sad;
        w->mode_pred[mb_type] = (int32_t)(result - (pix_t *)w);
    }
    WSYNC();
}

/* ------------------------------------------------------------------------------
 * a8: Intra4x4 decision + reconstruction of the 16 blocks in raster order
 * (intra_choose_4x4 H:4723, h264e_intra_choose_4x4 H:1810).  Returns the cost;
 * leaves recon in w->i4rec, levels in w->qv_y/dq_y, modes in w->i4_mode/i4_code and
 * the non-zero mask in *nz_mask_out.
 *
 * Warp mapping per 4x4 block: every predicted sample of every mode is one of 32 "source"
 * values derived from the 13 neighbouring samples Z[-4..8] = L3..L0, UL, U0..U7:
 *   S[0..3]   = Z[-4..-1] (L3..L0),  S[4..7] = Z[1..4] (U0..U3)      (H, V, HU tail)
 *   S[8..18]  = F3[k] = (Z[k-1] + 2 Z[k] + Z[k+1] + 2) >> 2,  k = -3..7
 *   S[19..28] = F2[k] = (Z[k] + Z[k+1] + 1) >> 1,             k = -4..5
 *   S[29] = (U6 + 3 U7 + 2) >> 2, S[30] = (L2 + 3 L3 + 2) >> 2, S[31] = DC
 * -- exactly 32 distinct values, one per lane
 * (ITU-T H.264 8.3.1.2.1-9 rewritten on one edge line).  The lanes build S[], then lane
 * (g, p) scores sample p for the modes of group g; nine SADs come out of five packed
 * warp reductions; transform, quantisation and reconstruction of the block run on 16
 * lanes with the 4-point butterflies exchanged through shared memory.
 * ---------------------------------------------------------------------------- */
/* source index of every predicted sample, [evaluation slot][y*4+x]; slots in the reference's
 * evaluation order DC, V, DDL, VL, H, HU, DDR, HD, VR (H:1834-1960) */
H264_TAB uint8_t i4_src_tab[9][16] = {
    {31,31,31,31,31,31,31,31,31,31,31,31,31,31,31,31},
    {4,5,6,7,4,5,6,7,4,5,6,7,4,5,6,7},
    {13,14,15,16,14,15,16,17,15,16,17,18,16,17,18,29},
    {24,25,26,27,13,14,15,16,25,26,27,28,14,15,16,17},
    {3,3,3,3,2,2,2,2,1,1,1,1,0,0,0,0},
    {21,9,20,8,20,8,19,30,19,30,0,0,0,0,0,0},
    {11,12,13,14,10,11,12,13,9,10,11,12,8,9,10,11},
    {22,11,12,13,21,10,22,11,20,9,21,10,19,8,20,9},
    {23,24,25,26,11,12,13,14,10,23,24,25,9,11,12,13},
};
H264_TAB uint8_t i4_slot_mode[9] = {2, 0, 3, 7, 1, 8, 4, 6, 5};

HD int fwd_row(int k, int a, int b, int c, int d)     /* k-th output of the forward 4-point kernel */
{
    int s = a + d, t = a - d, u = b + c, v = b - c;
    return k == 0 ? s + u : (k == 1 ? 2 * t + v : (k == 2 ? s - u : t - 2 * v));
}
HD int inv_row(int k, int a, int b, int c, int d)     /* k-th output of the inverse 4-point kernel */
{
    int e0 = a + c, e1 = a - c, e2 = (b >> 1) - d, e3 = b + (d >> 1);
    return k == 0 ? e0 + e3 : (k == 1 ? e1 + e2 : (k == 2 ? e1 - e2 : e0 - e3));
}

#if H264_DEVICE
HD int lds_u8(unsigned addr)
{
    int v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
/* sm_100a fast path of the Intra4x4 decision: same algorithm as the portable version below
 * (which documents it and is what the host emulation runs).  Differences in execution only:
 *  - a 4x4 block needs the reconstruction of its left, top, top-left and top-right blocks, so
 *    the 16 blocks are walked in anti-diagonal order t = c + 2r (10 steps instead of 16) with
 *    the two blocks of a step on the two half-warps (lane = sample y*4+x of its block);
 *  - lane j < 13 of a half-warp fetches neighbour Z[j-4] and derives F3[j-4], F2[j-4] and the
 *    two corner cases from its lane neighbours; the 31 source values go through a 32-byte
 *    shared table from which every lane picks its nine predicted samples;
 *  - the nine SADs are reduced with a halving exchange (each level a lane keeps half of its
 *    partial sums), the mode costs are formed where the sums end up and the strict-'<'-in-
 *    evaluation-order decision is a min-reduction of (cost << 4 | slot);
 *  - residual -> transform -> quantisation -> inverse -> reconstruction stay in registers. */
HDF_intra4_choose int intra4_choose(MBState &s, int *nz_mask_out, int cost16, int fixed_bound = -1)
{
    const FrameParams *fp = s.fp;
    MBWork *w = s.w;
    const int avail = s.avail;
    const int lane = LANE_ID, g = lane >> 4, l16 = lane & 15, px = l16 & 3, py = l16 >> 2;
    const unsigned FULL = 0xffffffffu;
    int cost = 0;
    int nz_mask = 0;
    const int penalty = (3 * fp->lambda_q4) >> 4;
    const int skip_thr = fp->skip_thr_i4x4;
    const int poll_skip = fp->slice_type == SLICE_P && fixed_bound < 0;

    /* padded reconstruction R[17][24]: row 0 = row above (TL at col 3, 16 + 4 samples from col 4),
     * col 3 = left column; sample (x,y) of the MB at R[(y+1)*24 + x+4] */
    pix_t *R = w->i4r;
    for (int i = lane; i < 21 + 16; i += 32)
    {
        if (i < 21) R[3 + i] = i == 0 ? w->tl[0] : w->top_y[i - 1];
        else R[(i - 21 + 1) * 24 + 3] = w->left_y[i - 21];
    }
    /* neighbour fetch offset of Z[l16-4] relative to R + (4r)*24 + 4c, with / without top-right */
    int zoff, zoff_notr;
    {
        int k = l16 - 4;
        if (l16 > 12) k = 0;
        zoff = k <= 0 ? (-k) * 24 + 3 : 4 + (k - 1);
        zoff_notr = k > 4 ? 4 + 3 : zoff;
    }
    /* source-value table of this half-warp: where this lane's values go, where its samples come from */
    pix_t *Sb = w->i4s + 32 * g;
    const int zi = l16 < 4 ? l16 : ((l16 >= 5 && l16 <= 8) ? l16 - 1 : -1);
    const int f3i = (l16 >= 1 && l16 <= 11) ? l16 + 7 : -1;
    const int f2i = l16 <= 9 ? 19 + l16 : -1;
    const int spi = l16 == 11 ? 29 : (l16 == 1 ? 30 : -1);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(Sb);
    unsigned ap[9];
#pragma unroll
    for (int k = 1; k < 9; k++) ap[k] = sbase + i4_src_tab[k][l16];
    /* block availability (block2avail H:4750), 4 bits per block */
    unsigned long long av64 = 0;
    for (int n = 0; n < 16; n++)
    {
        const int r = n >> 2, c = n & 3;
        int a = 0;
        if (c > 0 || (avail & AVAIL_L)) a |= AVAIL_L;
        if (r > 0 || (avail & AVAIL_T)) a |= AVAIL_T;
        if (r > 0 && c > 0) a |= AVAIL_TL;
        else if (r == 0 && c == 0) a |= avail & AVAIL_TL;
        else if (r == 0) a |= (avail & AVAIL_T) ? AVAIL_TL : 0;
        else a |= (avail & AVAIL_L) ? AVAIL_TL : 0;
        if (r == 0) { if (c < 3) a |= (avail & AVAIL_T) ? AVAIL_TR : 0; else a |= avail & AVAIL_TR; }
        else if (c < 3 && !((r & 1) && (c & 1))) a |= AVAIL_TR;
        av64 |= (unsigned long long)a << (4 * n);
    }
    /* the evaluation slots whose SADs end up in this lane after the halving reduction */
    const int b3 = (l16 >> 3) & 1, b2 = (l16 >> 2) & 1, b1 = (l16 >> 1) & 1;
    const int kA = b3 ? 8 : 4 * b2 + 2 * b1, kB = kA + 1;
#define I4_SLOT_MODE(k) ((k) == 0 ? 2 : ((k) == 1 ? 0 : ((k) == 2 ? 3 : ((k) == 3 ? 7 : ((k) == 4 ? 1 : ((k) == 5 ? 8 : ((k) == 6 ? 4 : ((k) == 7 ? 6 : 5))))))))
#define I4_SLOT_NEED(k) ((k) == 0 ? 0 : ((k) < 4 ? AVAIL_T : ((k) < 6 ? AVAIL_L : 7)))
    const int modeA = I4_SLOT_MODE(kA), modeB = I4_SLOT_MODE(kB);
    const int needA = I4_SLOT_NEED(kA), needB = b3 ? 16 : I4_SLOT_NEED(kB);      /* 16: unsatisfiable, there is no slot B */
    /* quantiser constants of coefficient i = v + 4u held by this lane (v = py, u = px) */
    const int ci = py + 4 * px;
    const int qcl = quant_class(ci);
    const int qmul = fp->qdat[0][qcl], dqmul = fp->qdat[0][qcl + 1], qrnd = fp->qdat[0][6];
    unsigned long long modes = 0, codes = 0;  /* chosen modes / coded values + 1 of blocks 0..15, 4 bits each */
    __syncwarp();

#pragma unroll 1
    for (int t = 0; t < 10; t++)
    {
        /* the candidate stage runs concurrently on another warp: stop as soon as it has
         * decided for an early skip (the intra result would be discarded, H:5767) */
        {
            const int st = poll_skip ? *(volatile int32_t *)&w->ic[IC_STATE] : 0;
            if (st == 1) { *nz_mask_out = 0; return 0x7FFFFFFF; }
            /* exact pruning: the blocks decided so far already cost as much as a competitor */
            int bound = cost16;
            if (st == 2) bound = imin(bound, inter_final_bound(fp, w));
            if (fixed_bound >= 0) bound = imin(bound, fixed_bound);
            if (cost + __shfl_xor_sync(FULL, cost, 16) + fp->lambda_i4_q4 >= bound) { *nz_mask_out = 0; return 0x7FFFFFFF; }
        }
        const int nA = t < 2 ? t : 4 * (t >> 1) + (t & 1) - 2;
        const int active = !g || (t >= 2 && t <= 7);
        const int n = (g && active) ? nA + 2 : nA;
        const int r = n >> 2, c = n & 3;
        const int a = (int)(av64 >> (4 * n)) & 15;

        /* most probable mode */
        const int ctx_l = c > 0 ? (int)((modes >> (4 * (n - 1))) & 15) : w->nb_i4mode[r];
        const int ctx_t = r > 0 ? (int)((modes >> (4 * (n - 4))) & 15) : w->nb_i4mode[4 + c];
        int mpred = imin(ctx_l, ctx_t);
        if (mpred < 0) mpred = 2;

        /* neighbours and input sample */
        const pix_t *Rb = R + (4 * r) * 24 + 4 * c;
        const int z = Rb[(a & AVAIL_TR) ? zoff : zoff_notr];
        const int in = w->inp_y[(4 * r + py) * 16 + 4 * c + px];

        /* source values -> table */
        int dc;
        {
            const int up = __shfl_up_sync(FULL, z, 1, 16), dn = __shfl_down_sync(FULL, z, 1, 16);
            const int t2 = z + dn;
            const int u4 = t2 + __shfl_down_sync(FULL, t2, 2, 16);
            if (zi >= 0) Sb[zi] = (pix_t)z;
            if (f3i >= 0) Sb[f3i] = (pix_t)((up + 2 * z + dn + 2) >> 2);
            if (f2i >= 0) Sb[f2i] = (pix_t)((z + dn + 1) >> 1);
            if (spi >= 0) Sb[spi] = (pix_t)(l16 == 11 ? (z + 3 * dn + 2) >> 2 : (z + 3 * up + 2) >> 2);
            const int sl = __shfl_sync(FULL, u4, 0, 16), su = __shfl_sync(FULL, u4, 5, 16);
            dc = (a & 3) == 3 ? (sl + su + 4) >> 3 : ((a & AVAIL_L) ? (sl + 2) >> 2 : ((a & AVAIL_T) ? (su + 2) >> 2 : 128));
        }
        __syncwarp();
        /* the nine predictions of this lane's sample and their absolute errors */
        int pv[9], dv[9];
        pv[0] = dc;
#pragma unroll
        for (int k = 1; k < 9; k++) pv[k] = lds_u8(ap[k]);
#pragma unroll
        for (int k = 0; k < 9; k++) dv[k] = iabs(in - pv[k]);
        int T;
        {   /* halving reduction over the 16 lanes; slots (kA, kB) end up in this lane */
            int P0 = dv[0] | (dv[1] << 16), P1 = dv[2] | (dv[3] << 16), P2 = dv[4] | (dv[5] << 16), P3 = dv[6] | (dv[7] << 16), P4 = dv[8];
            const int s0 = __shfl_xor_sync(FULL, b3 ? P0 : P4, 8);
            P1 += __shfl_xor_sync(FULL, P1, 8); P2 += __shfl_xor_sync(FULL, P2, 8); P3 += __shfl_xor_sync(FULL, P3, 8);
            if (b3) P4 += s0; else P0 += s0;
            const int a4 = __shfl_xor_sync(FULL, b3 ? P4 : (b2 ? P0 : P2), 4);
            const int b4 = __shfl_xor_sync(FULL, b2 ? P1 : P3, 4);
            const int Q0 = (b3 ? P4 : (b2 ? P2 : P0)) + a4, Q1 = (b2 ? P3 : P1) + b4;
            const int c2 = __shfl_xor_sync(FULL, b3 ? Q0 : (b1 ? Q0 : Q1), 2);
            T = (b3 ? Q0 : (b1 ? Q1 : Q0)) + c2;
            T += __shfl_xor_sync(FULL, T, 1);
        }
        /* strict '<' in evaluation order == minimum of (cost << 4 | slot) */
        int best;
        {
            const int cA = ((a & needA) == needA) ? ((((T & 0xFFFF) + (modeA != mpred ? penalty : 0)) << 4) | kA) : 0x7FFFFFFF;
            const int cB = ((a & needB) == needB) ? ((((int)((unsigned)T >> 16) + (modeB != mpred ? penalty : 0)) << 4) | kB) : 0x7FFFFFFF;
            best = imin(cA, cB);
#pragma unroll
            for (int o = 8; o; o >>= 1) best = imin(best, __shfl_xor_sync(FULL, best, o));
        }
        const int bk = best & 15, best_sad = best >> 4;
        const int mode = I4_SLOT_MODE(bk);
        const int code = mode == mpred ? 0 : (mode > mpred ? mode : mode + 1);      /* coded value + 1 */
        {   /* both halves record both blocks' modes */
            const int mo = __shfl_xor_sync(FULL, mode | (code << 4) | (n << 8), 16);
            const int no = mo >> 8;
            modes |= (unsigned long long)mode << (4 * n);
            codes |= (unsigned long long)code << (4 * n);
            if (no != n) { modes |= (unsigned long long)(mo & 15) << (4 * no); codes |= (unsigned long long)((mo >> 4) & 15) << (4 * no); }
        }
        const int do_tq = best_sad > skip_thr;

        /* prediction of the chosen mode, residual, transform, quantisation, reconstruction */
        int pred = pv[0];
#pragma unroll
        for (int k = 1; k < 9; k++) if (bk == k) pred = pv[k];
        const int res = do_tq ? in - pred : 0;
        /* forward: vertical pass -> (v = py, column px), horizontal pass -> (v = py, u = px) */
        const int f0 = __shfl_sync(FULL, res, px, 16), f1 = __shfl_sync(FULL, res, 4 + px, 16);
        const int f2_ = __shfl_sync(FULL, res, 8 + px, 16), f3_ = __shfl_sync(FULL, res, 12 + px, 16);
        const int t1 = fwd_row(py, f0, f1, f2_, f3_);
        const int g0 = __shfl_sync(FULL, t1, 4 * py, 16), g1 = __shfl_sync(FULL, t1, 4 * py + 1, 16);
        const int g2 = __shfl_sync(FULL, t1, 4 * py + 2, 16), g3 = __shfl_sync(FULL, t1, 4 * py + 3, 16);
        const int cf = (int16_t)fwd_row(px, g0, g1, g2, g3);
        const int qv = (cf * qmul + (cf < 0 ? 0xFFFF - qrnd : qrnd)) >> 16;
        const int dqv = (int16_t)(qv * dqmul);
        const int nzb = ((__ballot_sync(FULL, qv != 0) >> (16 * g)) & 0xFFFF) != 0;
        /* inverse: horizontal pass over u for this v, then vertical pass over v (all zero -> prediction) */
        const int h0 = __shfl_sync(FULL, dqv, 4 * py, 16), h1 = __shfl_sync(FULL, dqv, 4 * py + 1, 16);
        const int h2 = __shfl_sync(FULL, dqv, 4 * py + 2, 16), h3 = __shfl_sync(FULL, dqv, 4 * py + 3, 16);
        const int t2i = (int16_t)inv_row(px, h0, h1, h2, h3);
        const int v0 = __shfl_sync(FULL, t2i, px, 16), v1 = __shfl_sync(FULL, t2i, 4 + px, 16);
        const int v2 = __shfl_sync(FULL, t2i, 8 + px, 16), v3 = __shfl_sync(FULL, t2i, 12 + px, 16);
        const int rr = (int16_t)((inv_row(py, v0, v1, v2, v3) + 32) >> 6);
        const int outv = clip_u8(rr + pred);
        if (active)
        {
            R[(4 * r + py + 1) * 24 + 4 * c + px + 4] = (pix_t)outv;
            w->qv_y[n][ci] = (int16_t)qv;
            nz_mask |= nzb << (15 - n);
            cost += best_sad;
        }
        __syncwarp();
    }
#undef I4_SLOT_MODE
#undef I4_SLOT_NEED
    /* modes / coded values, and the reconstruction in the layout the rest of the MB code expects */
    if (lane < 16)
    {
        w->i4_mode[lane] = (int8_t)((modes >> (4 * lane)) & 15);
        w->i4_code[lane] = (int8_t)((int)((codes >> (4 * lane)) & 15) - 1);
    }
    for (int i = lane; i < 64; i += 32)
    {
        int rr = i >> 2, cc = (i & 3) * 4;
        *(uint32_t *)(w->i4rec + rr * 16 + cc) = *(const uint32_t *)(R + (rr + 1) * 24 + cc + 4);
    }
    __syncwarp();
    cost += __shfl_xor_sync(FULL, cost, 16);
    nz_mask |= __shfl_xor_sync(FULL, nz_mask, 16);
    *nz_mask_out = nz_mask;
    return cost + fp->lambda_i4_q4;
}
#else
HDF_intra4_choose int intra4_choose(MBState &s, int *nz_mask_out, int cost16, int fixed_bound = -1)
{
    const FrameParams *fp = s.fp;
    MBWork *w = s.w;
    const int avail = s.avail;
    int cost = fp->lambda_i4_q4;
    int nz_mask = 0;
    const int penalty = (3 * fp->lambda_q4) >> 4;
    const uint16_t *qdat = fp->qdat[0];

    for (int n = 0; n < 16; n++)
    {
        if (fixed_bound < 0 && fp->slice_type == SLICE_P && w->ic[IC_STATE] == 1) { *nz_mask_out = 0; return 0x7FFFFFFF; }
        {   /* exact pruning (see inter_cost_bound) */
            int bound = cost16;
            if (fixed_bound < 0 && fp->slice_type == SLICE_P && w->ic[IC_STATE] == 2) bound = imin(bound, inter_final_bound(fp, w));
            if (fixed_bound >= 0) bound = imin(bound, fixed_bound);
            if (cost >= bound) { *nz_mask_out = 0; return 0x7FFFFFFF; }
        }
        /* which neighbours exist for block n (block2avail H:4750) */
        const int r = n >> 2, c = n & 3;
        int a = 0;
        if (c > 0 || (avail & AVAIL_L)) a |= AVAIL_L;
        if (r > 0 || (avail & AVAIL_T)) a |= AVAIL_T;
        if (r > 0 && c > 0) a |= AVAIL_TL;
        else if (r == 0 && c == 0) a |= avail & AVAIL_TL;
        else if (r == 0) a |= (avail & AVAIL_T) ? AVAIL_TL : 0;
        else a |= (avail & AVAIL_L) ? AVAIL_TL : 0;
        if (r == 0) { if (c < 3) a |= (avail & AVAIL_T) ? AVAIL_TR : 0; else a |= avail & AVAIL_TR; }
        else if (c < 3 && !((r & 1) && (c & 1))) a |= AVAIL_TR;     /* raster blocks 4,6,8,9,10,12,14 */

        int ctx_l = c > 0 ? w->i4_mode[n - 1] : w->nb_i4mode[r];
        int ctx_t = r > 0 ? w->i4_mode[n - 4] : w->nb_i4mode[4 + c];
        int mpred = imin(ctx_l, ctx_t);
        if (mpred < 0) mpred = 2;

        const pix_t *blockin = w->inp_y + (c + r * 16) * 4;
        pix_t *block = w->i4rec + (c + r * 16) * 4;

        /* 1. the 13 neighbours Z[-4..8] -> i4z[0..12] */
        FOR_LANES(i, 13)
        {
            int k = i - 4, v = 0;
            if (k < 0) { int j = -1 - k; if (a & AVAIL_L) v = c > 0 ? w->i4rec[(r * 4 + j) * 16 + c * 4 - 1] : w->left_y[r * 4 + j]; }
            else if (k == 0)
            {
                if (a & AVAIL_TL)
                {
                    if (r > 0 && c > 0) v = w->i4rec[(r * 4 - 1) * 16 + c * 4 - 1];
                    else if (r > 0) v = w->left_y[r * 4 - 1];
                    else if (c > 0) v = w->top_y[c * 4 - 1];
                    else v = w->tl[0];
                }
            } else if (a & AVAIL_T)
            {
                int j = k - 1;
                if (j > 3 && !(a & AVAIL_TR)) j = 3;
                int x = c * 4 + j;
                v = r > 0 ? w->i4rec[(r * 4 - 1) * 16 + x] : w->top_y[x];
            }
            w->i4z[i] = (pix_t)v;
        }
        WSYNC();
        /* 2. the 32 source values */
        FOR_LANES(i, 32)
        {
            const pix_t *Z = w->i4z + 4;
            int v;
            if (i < 4) v = Z[i - 4];
            else if (i < 8) v = Z[i - 3];
            else if (i < 19) { int k = i - 11; v = (Z[k - 1] + 2 * Z[k] + Z[k + 1] + 2) >> 2; }
            else if (i < 29) { int k = i - 23; v = (Z[k] + Z[k + 1] + 1) >> 1; }
            else if (i == 29) v = (Z[7] + 3 * Z[8] + 2) >> 2;
            else if (i == 30) v = (Z[-3] + 3 * Z[-4] + 2) >> 2;
            else
            {
                int sum = 0, cnt = 0;
                if (a & AVAIL_L) { sum += Z[-1] + Z[-2] + Z[-3] + Z[-4]; cnt++; }
                if (a & AVAIL_T) { sum += Z[1] + Z[2] + Z[3] + Z[4]; cnt++; }
                v = cnt == 0 ? 128 : (cnt == 2 ? (sum + 4) >> 3 : (sum + 2) >> 2);
            }
            w->i4s[i] = (pix_t)v;
        }
        WSYNC();
        /* 3. nine SADs: lane (g,p) scores sample p for slots 0..4 (g = 0) or 5..8 (g = 1) */
        int acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, acc4 = 0;   /* slots (0,1) (2,3) (4,5) (6,7) (8) packed 16|16 */
        FOR_LANES(l, 32)
        {
            int g = l >> 4, p = l & 15;
            int in = blockin[(p >> 2) * 16 + (p & 3)];
            if (g == 0)
            {
                acc0 += iabs(in - w->i4s[i4_src_tab[0][p]]) | (iabs(in - w->i4s[i4_src_tab[1][p]]) << 16);
                acc1 += iabs(in - w->i4s[i4_src_tab[2][p]]) | (iabs(in - w->i4s[i4_src_tab[3][p]]) << 16);
                acc2 += iabs(in - w->i4s[i4_src_tab[4][p]]);
            } else
            {
                acc2 += iabs(in - w->i4s[i4_src_tab[5][p]]) << 16;
                acc3 += iabs(in - w->i4s[i4_src_tab[6][p]]) | (iabs(in - w->i4s[i4_src_tab[7][p]]) << 16);
                acc4 += iabs(in - w->i4s[i4_src_tab[8][p]]);
            }
        }
        acc0 = wsum(acc0); acc1 = wsum(acc1); acc2 = wsum(acc2); acc3 = wsum(acc3); acc4 = wsum(acc4);
        int sads[9];
        sads[0] = acc0 & 0xFFFF; sads[1] = (int)((uint32_t)acc0 >> 16);
        sads[2] = acc1 & 0xFFFF; sads[3] = (int)((uint32_t)acc1 >> 16);
        sads[4] = acc2 & 0xFFFF; sads[5] = (int)((uint32_t)acc2 >> 16);
        sads[6] = acc3 & 0xFFFF; sads[7] = (int)((uint32_t)acc3 >> 16);
        sads[8] = acc4;
        int bk = 0, best_sad = 0x7FFFFFFF;
#pragma unroll
        for (int k = 0; k < 9; k++)
        {
            int ok = k == 0 ? 1 : (k < 4 ? (a & AVAIL_T) : (k < 6 ? (a & AVAIL_L) : ((a & 7) == 7)));
            int cst = sads[k] + (i4_slot_mode[k] != mpred ? penalty : 0);
            if (ok && cst < best_sad) { best_sad = cst; bk = k; }
        }
        const int mode = i4_slot_mode[bk];
        const int do_tq = best_sad > fp->skip_thr_i4x4;

        /* 4. prediction into the reconstruction buffer, residual for the transform */
        FOR_LANES(p, 16)
        {
            int pv = w->i4s[i4_src_tab[bk][p]];
            int o = (p >> 2) * 16 + (p & 3);
            block[o] = (pix_t)pv;
            w->i4t[p] = (int16_t)((int)blockin[o] - pv);
            if (!do_tq) { w->qv_y[n][p] = 0; w->dq_y[n][p] = 0; }
            if (p == 0)
            {
                w->i4_mode[n] = (int8_t)mode;
                w->i4_code[n] = (int8_t)(mode == mpred ? -1 : (mode > mpred ? mode - 1 : mode));
            }
        }
        WSYNC();
        int nzb = 0;
        if (do_tq)
        {
            /* forward transform: vertical pass (lane = column x, vertical frequency v) ... */
            FOR_LANES(p, 16)
            {
                int x = p & 3, v = p >> 2;
                w->i4u[v * 4 + x] = (int16_t)fwd_row(v, w->i4t[x], w->i4t[4 + x], w->i4t[8 + x], w->i4t[12 + x]);
            }
            WSYNC();
            /* ... horizontal pass + quantisation (lane = coefficient i = v + 4u) */
            int nzl = 0;
            FOR_LANES(i, 16)
            {
                int v = i & 3, u = i >> 2;
                int cf = (int16_t)fwd_row(u, w->i4u[v * 4], w->i4u[v * 4 + 1], w->i4u[v * 4 + 2], w->i4u[v * 4 + 3]);
                int cl = quant_class(i);
                int rnd = cf < 0 ? 0xFFFF - qdat[6] : qdat[6];
                int q = (cf * (int)qdat[cl] + rnd) >> 16;
                w->qv_y[n][i] = (int16_t)q;
                w->dq_y[n][i] = (int16_t)(q * (int)qdat[cl + 1]);
                nzl |= q;
            }
            nzb = wor(nzl) != 0;
            WSYNC();
            if (nzb)
            {
                /* inverse: horizontal pass (lane = vertical frequency v, column x) ... */
                FOR_LANES(p, 16)
                {
                    int x = p & 3, v = p >> 2;
                    const int16_t *dq = w->dq_y[n];
                    w->i4u[v * 4 + x] = (int16_t)inv_row(x, dq[v], dq[v + 4], dq[v + 8], dq[v + 12]);
                }
                WSYNC();
                /* ... vertical pass, add prediction, clip */
                FOR_LANES(p, 16)
                {
                    int x = p & 3, y = p >> 2;
                    int rr = (int16_t)((inv_row(y, w->i4u[x], w->i4u[4 + x], w->i4u[8 + x], w->i4u[12 + x]) + 32) >> 6);
                    int o = y * 16 + x;
                    block[o] = (pix_t)clip_u8(rr + block[o]);
                }
                WSYNC();
            }
        }
        nz_mask = (nz_mask << 1) | nzb;
        cost += best_sad;
    }
    *nz_mask_out = nz_mask;
    return cost;
}
#endif

/* ------------------------------------------------------------------------------
 * a9-a11: luma transform / quant / reconstruction of a non-I4x4 macroblock
 * (mb_write H:4423-4434 with h264e_transform_sub_quant_dequant H:2619), one warp per half
 * of the macroblock (blocks 8*half .. 8*half+7; the 8x8 zeroing groups never straddle the
 * halves).  For Intra16x16 the two warps meet at named barrier 2 around the DC transform.
 * Publishes the half's block mask in w->tq_res[half] (bit 15 = block 0).
 * ---------------------------------------------------------------------------- */
HDN void luma_tq_half(MBState &s, int half, int intra16, int stage)
{
    const FrameParams *fp = s.fp;
    MBWork *w = s.w;
    const uint16_t *qdat = fp->qdat[0];
    pix_t *dec = fp->dec[0] + (s.mby * 16) * fp->stride[0] + s.mbx * 16;
    const int b0 = half * 8;
    if (stage == 0)
    {
        FOR_LANES(k, 8)
        {
            int b = b0 + k, off = (b & 3) * 4 + (b >> 2) * 64;
            fwd4x4(w->inp_y + off, 16, s.pbest + off, w->dq_y[b]);
            if (intra16) w->dc_y[b] = w->dq_y[b][0];
            else
            {
                w->zflag1[b] = (int8_t)coefs_small(w->dq_y[b], 0, qdat + 10);
                w->zflag2[b] = (int8_t)coefs_small(w->dq_y[b], 0, qdat + 18);
            }
        }
        WSYNC();
        int zmask = 0;
        if (!intra16)       /* zero_smallq H:2512: drop isolated small blocks / 8x8 groups */
        {
            for (int k = 0; k < 8; k++) if (w->zflag1[b0 + k]) zmask |= 1 << (b0 + k);
            for (int g = 0; g < 2; g++)
            {
                int f = b0 + 2 * g, m = 0x33 << f;
                if ((~zmask & m) && w->zflag2[f] && w->zflag2[f + 1] && w->zflag2[f + 4] && w->zflag2[f + 5]) zmask |= m;
            }
        }
        int nzbits = 0;
        FOR_LANES(k, 8)
        {
            int b = b0 + k, nz = 0;
            if (zmask & (1 << b)) { for (int i = 0; i < 16; i++) w->qv_y[b][i] = 0; }
            else nz = quant4x4(w->dq_y[b], w->qv_y[b], intra16, qdat);
            if (nz) nzbits |= 0x8000 >> b;
        }
        nzbits = wor(nzbits);
        IF_LANE0 { w->tq_res[half] = nzbits; }
        WSYNC();
        return;
    }
    if (stage == 1)
    {   /* Intra16x16 DC path, once for the macroblock */
        IF_LANE0
        {
            int16_t dq0[16];
            luma_dc_quant(w->dc_y, w->qdc_y, dq0, qdat);
            for (int b = 0; b < 16; b++) w->dq_y[b][0] = dq0[b];
        }
        WSYNC();
        return;
    }
    const int recon_mask = intra16 ? 0xFFFF : w->tq_res[half];
    FOR_LANES(k, 8)
    {
        int b = b0 + k, off = (b & 3) * 4 + (b >> 2) * 64;
        pix_t *o = dec + (b & 3) * 4 + (b >> 2) * 4 * fp->stride[0];
        if (recon_mask & (0x8000 >> b)) inv4x4_add(w->dq_y[b], s.pbest + off, o, fp->stride[0]);
        else copy4x4(s.pbest + off, o, fp->stride[0]);
    }
    WSYNC();
}

/* chroma transform / quant / recon of plane pl (mb_write H:4443-4491), one warp.
 * Publishes w->tq_res[2 + pl] = nz mask (bit 3 = block 0) | dc_flag << 8. */
HDN void chroma_tq_plane(MBState &s, int pl)
{
    const FrameParams *fp = s.fp;
    MBWork *w = s.w;
    const uint16_t *qdat = fp->qdat[1];
    const int sc = fp->stride[1];
    FOR_LANES(k, 4)
    {
        int b = pl * 4 + k;
        int off = pl * 8 + (k & 1) * 4 + (k >> 1) * 64;
        fwd4x4(w->inp_c + off, 16, w->predc + off, w->dq_c[b]);
        w->dc_c[b] = w->dq_c[b][0];
        int nz = 0;
        if (coefs_small(w->dq_c[b], 1, qdat + 10)) { for (int i = 0; i < 16; i++) w->qv_c[b][i] = 0; }
        else nz = quant4x4(w->dq_c[b], w->qv_c[b], 1, qdat);
        w->zflagc[b] = (int8_t)nz;
    }
    WSYNC();
    int nzm = 0;
    for (int k = 0; k < 4; k++) if (w->zflagc[pl * 4 + k]) nzm |= 8 >> k;
    IF_LANE0
    {
        int16_t dq0[4];
        int dcf = chroma_dc_quant(w->dc_c + 4 * pl, w->qdc_c + 4 * pl, dq0, qdat);
        for (int k = 0; k < 4; k++) w->dq_c[pl * 4 + k][0] = dq0[k];
        w->tq_res[2 + pl] = nzm | (dcf << 8);
    }
    WSYNC();
    const int dcf = w->tq_res[2 + pl] >> 8;
    pix_t *dec = fp->dec[1 + pl] + (s.mby * 8) * sc + s.mbx * 8;
    FOR_LANES(k, 4)
    {
        int b = pl * 4 + k;
        int off = pl * 8 + (k & 1) * 4 + (k >> 1) * 64;
        pix_t *o = dec + (k & 1) * 4 + (k >> 1) * 4 * sc;
        if (!(dcf | nzm)) copy4x4(w->predc + off, o, sc);
        else if (dcf)
        {
            if (!(nzm & (8 >> k))) for (int i = 1; i < 16; i++) w->dq_c[b][i] = 0;
            inv4x4_add(w->dq_c[b], w->predc + off, o, sc);
        } else
        {
            if (nzm & (8 >> k)) inv4x4_add(w->dq_c[b], w->predc + off, o, sc);
            else copy4x4(w->predc + off, o, sc);
        }
    }
    WSYNC();
}

#if H264_DEVICE
/* ------------------------------------------------------------------------------
 * sm_100a fast path of a9-a11 (same arithmetic as luma_tq_half / chroma_tq_plane above, which
 * document it and are what the host emulation runs): lane l owns sample row r = l & 3 of 4x4
 * block l >> 2; the horizontal 4-point kernels run inside the lane, the vertical ones across
 * the four lanes of a block with two xor-shuffles per value.  After the forward transform lane
 * r holds the coefficients of vertical frequency v = {0,2,3,1}[r], horizontal u = 0..3, i.e.
 * indices v + 4u of the reference's transposed layout (H:2391).
 * ---------------------------------------------------------------------------- */
struct TQLane { int c[4]; int v; };

HD void tq_fwd_lane(uint32_t in4, uint32_t pr4, int r, TQLane &t)
{
    const unsigned FULL = 0xffffffffu;
    int d0 = (int)(in4 & 255) - (int)(pr4 & 255), d1 = (int)((in4 >> 8) & 255) - (int)((pr4 >> 8) & 255);
    int d2 = (int)((in4 >> 16) & 255) - (int)((pr4 >> 16) & 255), d3 = (int)(in4 >> 24) - (int)(pr4 >> 24);
    int s03 = d0 + d3, e03 = d0 - d3, s12 = d1 + d2, e12 = d1 - d2;
    int h[4];
    h[0] = s03 + s12; h[1] = 2 * e03 + e12; h[2] = s03 - s12; h[3] = e03 - 2 * e12;
#pragma unroll
    for (int u = 0; u < 4; u++)
    {
        int p = __shfl_xor_sync(FULL, h[u], 3);
        int a = r < 2 ? h[u] + p : p - h[u];             /* r: 0 s03, 1 s12, 2 d12, 3 d03 */
        int q = __shfl_xor_sync(FULL, a, 1);
        t.c[u] = (int16_t)(r == 0 ? a + q : (r == 1 ? q - a : (r == 2 ? q - 2 * a : 2 * a + q)));
    }
    t.v = r == 0 ? 0 : (r == 1 ? 2 : (r == 2 ? 3 : 1));
}

/* inverse transform of the lane-distributed dequantised block + prediction -> 4 packed samples of row r */
HD uint32_t tq_inv_lane(const int d[4], uint32_t pr4, int r)
{
    const unsigned FULL = 0xffffffffu;
    int e0 = d[0] + d[2], e1 = d[0] - d[2], e2 = (d[1] >> 1) - d[3], e3 = d[1] + (d[3] >> 1);
    int t[4];
    t[0] = (int16_t)(e0 + e3); t[1] = (int16_t)(e1 + e2); t[2] = (int16_t)(e1 - e2); t[3] = (int16_t)(e0 - e3);
    uint32_t out = 0;
#pragma unroll
    for (int x = 0; x < 4; x++)
    {
        int p = __shfl_xor_sync(FULL, t[x], 1);          /* r: 0 f0|f2, 1 f2|f0, 2 f3|f1, 3 f1|f3 */
        int g = r == 0 ? t[x] + p : (r == 1 ? p - t[x] : (r == 2 ? (p >> 1) - t[x] : t[x] + (p >> 1)));
        int q = __shfl_xor_sync(FULL, g, 3);             /* r: 0 g0|g3, 1 g1|g2, 2 g2|g1, 3 g3|g0 */
        int val = (r == 0 || r == 1) ? g + q : q - g;
        int rr = (int16_t)((val + 32) >> 6);
        out |= (uint32_t)clip_u8(rr + (int)((pr4 >> (8 * x)) & 255)) << (8 * x);
    }
    return out;
}

/* 1 where every nibble of m is 0xF, at the nibble's lowest bit */
HD uint32_t nibble_all(uint32_t m) { m &= m >> 1; m &= m >> 2; return m & 0x11111111u; }
HD uint32_t nibble_any(uint32_t m) { m |= m >> 1; m |= m >> 2; return m & 0x11111111u; }

/* luma of a non-I4x4 macroblock; called by warps 0 and 1 (half = warp) */
HDF_luma_tq_fast void luma_tq_fast(MBState &s, int half, int intra16)
{
    const unsigned FULL = 0xffffffffu;
    const FrameParams *fp = s.fp;
    MBWork *w = s.w;
    const uint16_t *qdat = fp->qdat[0];
    const int lane = LANE_ID, r = lane & 3, k = lane >> 2, b = half * 8 + k;
    const int off = (b & 3) * 4 + (b >> 2) * 64 + r * 16;
    const uint32_t in4 = ld4_sm(w->inp_y + off), pr4 = ld4_sm(s.pbest + off);
    TQLane t;
    tq_fwd_lane(in4, pr4, r, t);
    const int v = t.v;
    int zero_blk = 0;
    if (intra16) { if (r == 0) w->dc_y[b] = (int16_t)t.c[0]; }
    else
    {   /* zero_smallq H:2512: blocks / 8x8 groups whose coefficients are all below the thresholds */
        const int t1a = qdat[10 + v], t1b = qdat[10 + v + 4], t2a = qdat[18 + v], t2b = qdat[18 + v + 4];
        int s1 = 1, s2 = 1;
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
            int ta = (u & 1) ? t1b : t1a, tb = (u & 1) ? t2b : t2a;
            if ((unsigned)(t.c[u] + ta) > 2u * ta) s1 = 0;
            if ((unsigned)(t.c[u] + tb) > 2u * tb) s2 = 0;
        }
        const uint32_t x1 = nibble_all(__ballot_sync(FULL, s1)), x2 = nibble_all(__ballot_sync(FULL, s2));
        const uint32_t mg = 0x00110011u << (8 * ((k & 3) >> 1));
        zero_blk = (int)((x1 >> (4 * k)) & 1) | (((x1 & mg) != mg && (x2 & mg) == mg) ? 1 : 0);
    }
    /* quantisation / dequantisation (quantize H:2567-2585) */
    const int rnd = qdat[6];
    const int cl0 = (v & 1) * 2;
    const int qm0 = qdat[cl0], dq0m = qdat[cl0 + 1], qm1 = qdat[cl0 + 2], dq1m = qdat[cl0 + 3];
    int q[4], d[4], any = 0;
#pragma unroll
    for (int u = 0; u < 4; u++)
    {
        int c = t.c[u];
        int qq = (c * ((u & 1) ? qm1 : qm0) + (c < 0 ? 0xFFFF - rnd : rnd)) >> 16;
        if (zero_blk || (intra16 && u == 0 && v == 0)) qq = 0;
        q[u] = qq; any |= qq;
        d[u] = (int16_t)(qq * ((u & 1) ? dq1m : dq0m));
        w->qv_y[b][v + 4 * u] = (int16_t)qq;
    }
    const uint32_t nzn = nibble_any(__ballot_sync(FULL, any != 0));
    int nzbits = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) if ((nzn >> (4 * j)) & 1) nzbits |= 0x8000 >> (half * 8 + j);
    if (lane == 0) w->tq_res[half] = nzbits;
    if (intra16)
    {   /* DC path (h264e_quant_luma_dc H:2344) on 16 lanes of warp 0; the halves meet at named barriers */
        bar_sync(2, 64);
        if (half == 0)
        {
            const int j = lane & 15, ji = j >> 2, jk = j & 3;
            int x = w->dc_y[j];
#pragma unroll
            for (int rep = 0; rep < 2; rep++)
            {
                int a, bb, c, dd, ss, tt, uu, ww;
                a = __shfl_sync(FULL, x, ji); bb = __shfl_sync(FULL, x, ji + 4); c = __shfl_sync(FULL, x, ji + 8); dd = __shfl_sync(FULL, x, ji + 12);
                ss = a + c; tt = a - c; uu = bb + dd; ww = bb - dd;
                int t1 = (int16_t)(jk == 0 ? ss + uu : (jk == 1 ? tt + ww : (jk == 2 ? tt - ww : ss - uu)));      /* t[4i + k] */
                a = __shfl_sync(FULL, t1, jk); bb = __shfl_sync(FULL, t1, jk + 4); c = __shfl_sync(FULL, t1, jk + 8); dd = __shfl_sync(FULL, t1, jk + 12);
                ss = a + c; tt = a - c; uu = bb + dd; ww = bb - dd;
                x = (int16_t)(ji == 0 ? ss + uu : (ji == 1 ? tt + ww : (ji == 2 ? tt - ww : ss - uu)));          /* x[k + 4m] */
                if (rep == 0)
                {
                    int qd = (int16_t)qdat[0];
                    int vq = (x * qd + (x < 0 ? (1 << 18) - 0x20000 : 0x20000)) >> 18;
                    if (lane < 16) w->qdc_y[j] = (int16_t)vq;
                    x = (int16_t)vq;
                }
            }
            if (lane < 16) w->dq_y[j][0] = (int16_t)(x * (int)(int16_t)(qdat[1] >> 2));
        }
        bar_sync(3, 64);
        if (v == 0) d[0] = w->dq_y[b][0];
    } else if (!((nzn >> (4 * k)) & 1)) { d[0] = d[1] = d[2] = d[3] = 0; }
    const uint32_t o4 = tq_inv_lane(d, pr4, r);
    pix_t *dec = fp->dec[0] + (s.mby * 16 + (b >> 2) * 4 + r) * fp->stride[0] + s.mbx * 16 + (b & 3) * 4;
    *(uint32_t *)dec = o4;
    WSYNC();
}

/* chroma plane pl; one warp, lanes 16-31 mirror lanes 0-15 */
HDF_chroma_tq_fast void chroma_tq_fast(MBState &s, int pl)
{
    const unsigned FULL = 0xffffffffu;
    const FrameParams *fp = s.fp;
    MBWork *w = s.w;
    const uint16_t *qdat = fp->qdat[1];
    const int lane = LANE_ID, l16 = lane & 15, r = l16 & 3, k = l16 >> 2, b = pl * 4 + k;
    const int off = pl * 8 + (k & 1) * 4 + (k >> 1) * 64 + r * 16;
    const uint32_t in4 = ld4_sm(w->inp_c + off), pr4 = ld4_sm(w->predc + off);
    TQLane t;
    tq_fwd_lane(in4, pr4, r, t);
    const int v = t.v;
    const int t1a = qdat[10 + v], t1b = qdat[10 + v + 4];
    int s1 = 1;
#pragma unroll
    for (int u = 0; u < 4; u++)
    {
        int ta = (u & 1) ? t1b : t1a;
        if (!(u == 0 && v == 0) && (unsigned)(t.c[u] + ta) > 2u * ta) s1 = 0;
    }
    const uint32_t x1 = nibble_all(__ballot_sync(FULL, s1));
    const int zero_blk = (int)((x1 >> (4 * k)) & 1);
    const int rnd = qdat[6];
    const int cl0 = (v & 1) * 2;
    const int qm0 = qdat[cl0], dq0m = qdat[cl0 + 1], qm1 = qdat[cl0 + 2], dq1m = qdat[cl0 + 3];
    int d[4], any = 0;
#pragma unroll
    for (int u = 0; u < 4; u++)
    {
        int c = t.c[u];
        int qq = (c * ((u & 1) ? qm1 : qm0) + (c < 0 ? 0xFFFF - rnd : rnd)) >> 16;
        if (zero_blk || (u == 0 && v == 0)) qq = 0;
        any |= qq;
        d[u] = (int16_t)(qq * ((u & 1) ? dq1m : dq0m));
        if (lane < 16) w->qv_c[b][v + 4 * u] = (int16_t)qq;
    }
    const uint32_t nzn = nibble_any(__ballot_sync(FULL, any != 0)) & 0x1111u;
    const int nzm = ((nzn & 1) ? 8 : 0) | ((nzn & 0x10) ? 4 : 0) | ((nzn & 0x100) ? 2 : 0) | ((nzn & 0x1000) ? 1 : 0);
    /* DC path (h264e_quant_chroma_dc H:2355), computed redundantly by every lane */
    int dcf, dq0k;
    {
        const int a = __shfl_sync(FULL, t.c[0], 0), bb = __shfl_sync(FULL, t.c[0], 4);
        const int c = __shfl_sync(FULL, t.c[0], 8), dd = __shfl_sync(FULL, t.c[0], 12);
        int x0 = (int16_t)(a + bb + c + dd), x1d = (int16_t)(a - bb + c - dd), x2 = (int16_t)(a + bb - c - dd), x3 = (int16_t)(a - bb - c + dd);
        const int qd = (int16_t)(qdat[0] << 1);
        x0 = (int16_t)((x0 * qd + (x0 < 0 ? (1 << 18) - 0xAAAA : 0xAAAA)) >> 18);
        x1d = (int16_t)((x1d * qd + (x1d < 0 ? (1 << 18) - 0xAAAA : 0xAAAA)) >> 18);
        x2 = (int16_t)((x2 * qd + (x2 < 0 ? (1 << 18) - 0xAAAA : 0xAAAA)) >> 18);
        x3 = (int16_t)((x3 * qd + (x3 < 0 ? (1 << 18) - 0xAAAA : 0xAAAA)) >> 18);
        if (lane < 4) w->qdc_c[4 * pl + lane] = (int16_t)(lane == 0 ? x0 : (lane == 1 ? x1d : (lane == 2 ? x2 : x3)));
        const int y0 = (int16_t)(x0 + x1d + x2 + x3), y1 = (int16_t)(x0 - x1d + x2 - x3);
        const int y2 = (int16_t)(x0 + x1d - x2 - x3), y3 = (int16_t)(x0 - x1d - x2 + x3);
        const int dqm = (int16_t)(qdat[1] >> 1);
        dcf = (y0 | y1 | y2 | y3) != 0;
        dq0k = (int16_t)((k == 0 ? y0 : (k == 1 ? y1 : (k == 2 ? y2 : y3))) * dqm);
    }
    if (lane == 0) w->tq_res[2 + pl] = nzm | (dcf << 8);
    if (!((nzn >> (4 * k)) & 1)) { d[0] = d[1] = d[2] = d[3] = 0; }
    if (v == 0) d[0] = dq0k;
    const uint32_t o4 = tq_inv_lane(d, pr4, r);
    if (lane < 16)
    {
        pix_t *dec = fp->dec[1 + pl] + (s.mby * 8 + (k >> 1) * 4 + r) * fp->stride[1] + s.mbx * 8 + (k & 1) * 4;
        *(uint32_t *)dec = o4;
    }
    WSYNC();
}
#endif

/* chroma intra prediction of plane pl (warp-level) */
HDF_intra_chroma_plane void intra_chroma_plane(MBState &s, int pl)
{
    MBWork *w = s.w;
    const pix_t *left = (s.avail & AVAIL_L) ? w->left_c : 0, *top = (s.avail & AVAIL_T) ? w->top_c : 0;
    const int mode = s.i16_mode;
    FOR_LANES(i, 16)
    {
        int r = i >> 1, xh = i & 1, yh = r >> 2;
        uint32_t v;
        if (mode == 0) v = ld4_sm(top + pl * 8 + xh * 4);
        else if (mode == 1) v = left[pl * 8 + r] * 0x01010101u;
        else
        {
            const pix_t *l = left ? left + pl * 8 + yh * 4 : 0;
            const pix_t *t = top ? top + pl * 8 + xh * 4 : 0;
            int dc;
            if (xh == yh) dc = dc_pred(l, t, 4, 2);
            else if (xh) dc = t ? dc_pred(0, t, 4, 2) : dc_pred(l, 0, 4, 2);
            else dc = l ? dc_pred(l, 0, 4, 2) : dc_pred(0, t, 4, 2);
            v = (uint32_t)dc * 0x01010101u;
        }
        *(uint32_t *)(w->predc + r * 16 + pl * 8 + xh * 4) = v;
    }
    WSYNC();
}

HD int count_nz(const int16_t *q, int i0)
{
    int n = 0;
    for (int i = i0; i < 16; i++) n += q[i] != 0;
    return n;
}

/* mv_clusters_update H:5263 */
HD void clusters_update(int32_t *cl, int mv)
{
    int x = mv_x(mv), y = mv_y(mv);
    int norm = x * x + y * y;
    int c0x = mv_x(cl[0]), c0y = mv_y(cl[0]), c1x = mv_x(cl[1]), c1y = mv_y(cl[1]);
    int n0 = c0x * c0x + c0y * c0y, n1 = c1x * c1x + c1y * c1y;
    if (norm < n1) cl[0] = mv_pack((63 * c0x + x + 32) >> 6, (63 * c0y + y + 32) >> 6);
    if (norm >= n0) cl[1] = mv_pack((63 * c1x + x + 32) >> 6, (63 * c1y + y + 32) >> 6);
}

/* The hinted partition modes (16x8, 8x16, 8x8) are independent search tasks; the warps that are
 * free (roles 1, 2, and 3 once the intra decision is over) take them from a shared counter, the
 * longest (8x8, four partitions) first.  `slot` selects the warp's private scratch. */
HDF_partition_tasks void partition_tasks(MBState &s, int slot)
{
    MBWork *w = s.w;
    const int pref = w->ic[IC_PREF];
    s.ss = &w->ss[SS_SLOT(slot)];
    for (;;)
    {
        int k = 0;
#if H264_DEVICE
        if (LANE_ID == 0) k = atomicAdd(&w->task_next, 1);
        k = __shfl_sync(0xffffffffu, k, 0);
#else
        k = w->task_next++;
#endif
        int t = -1;
        if ((pref & 8) && k-- == 0) t = 3;
        else if ((pref & 2) && k-- == 0) t = 1;
        else if ((pref & 4) && k-- == 0) t = 2;
        if (t < 0) break;
        inter_mode_search(s, t);
    }
}

/* The motion-estimation phase of a P macroblock: candidate stage on warp 0, then the 16x16 search there and the hinted
 * partition modes on warps 1 and 2 (single-warp build: one after the other).  Leaves its results in w->ic[],
 * w->mode_cost[], w->part_mv[][], w->part_mvd[][] (+ prediction blocks in the pixel flavour). */
HD void me_phase(MBState &s, const int32_t cl[2])
{
    MBWork *w = s.w;
    ON_WARP(0)
    {
        s.ss = &w->ss[0];
        inter_stage_a(s, cl);
        PROF_MARK(s, 2);
        bar_sync(1, 96);
        if (w->ic[IC_STATE] != 1) inter_mode_search(s, 0);
        PROF_MARK(s, 4);
    }
    ON_WARP(1) { bar_sync(1, 96); if (w->ic[IC_STATE] != 1) partition_tasks(s, 1); }
    ON_WARP(2) { bar_sync(1, 96); if (w->ic[IC_STATE] != 1) partition_tasks(s, 2); }
}

/* Inter decision over the results of me_phase() (H:5500-5522): early skip, else the cheapest searched mode in ascending
 * order with strict '<', else P16x16 at the skip vector when the raw skip SAD is smaller.  Every thread, same values.
 * Returns the winning partition mode of the search (what MBSpec::inter_best records). */
HD int inter_decide_p(const FrameParams *fp, const int32_t *ic, const int32_t *mode_cost, const int32_t *part_mv, const int32_t *part_mvd,
                      int *ptype, int *pcost, int32_t pmv[4], int32_t pmvd[4], int *use_skip_pred)
{
    const int mvp16 = ic[IC_MVP16], mv_skip = ic[IC_MV_SKIP];
    *use_skip_pred = 0;
    if (ic[IC_STATE] == 1)
    {
        *ptype = MBT_SKIP; *pcost = 0;
        pmv[0] = mv_skip;
        *use_skip_pred = 1;
        return 0;
    }
    int cost = 0xffffff, best_type = 0;
    for (int t = 0; t < 4; t++)
        if ((ic[IC_PREF] >> t) & 1)
            if (mode_cost[t] < cost) { cost = mode_cost[t]; best_type = t; }
    int type = best_type;
    for (int i = 0; i < 4; i++) { pmv[i] = part_mv[best_type * 4 + i]; pmvd[i] = part_mvd[best_type * 4 + i]; }
    if (cost > ic[IC_SAD_SKIP])     /* P16x16 at the skip vector is cheaper (H:5512) */
    {
        type = 0;
        cost = ic[IC_SAD_SKIP] + mv_cost(mv_skip, mvp16, fp->lambda_mv_q4);
        pmv[0] = mv_skip;
        pmvd[0] = mv_sub2(mv_skip, mvp16);
        *use_skip_pred = 1;      /* same samples the reference re-interpolates (H:5520) */
    }
    *ptype = type; *pcost = cost;
    return best_type;
}
HD int inter_decide(const FrameParams *fp, const MBWork *w, int *ptype, int *pcost, int32_t pmv[4], int32_t pmvd[4], int *use_skip_pred)
{
    return inter_decide_p(fp, w->ic, w->mode_cost, &w->part_mv[0][0], &w->part_mvd[0][0], ptype, pcost, pmv, pmvd, use_skip_pred);
}

/* Does the speculative motion-estimation record of the macroblock (h264_wave.h, me_prepass_mb) belong to exactly the
 * inputs the macroblock really has -- the 13 context vectors and the two cluster candidates?  Then running the
 * estimation again would reproduce the record word for word.  Every warp evaluates it (same answer in all). */
HD int me_record_matches(const MBState &s, const uint32_t *mr, const int32_t cl[2])
{
    const MBWork *w = s.w;
    int ok = 1;
    FOR_LANES(i, 16)
    {
        int32_t have;
        if (i < 4) have = w->mvp0_left[i];
        else if (i < 8) have = w->mvp0_tl[i - 4];
        else if (i < 13) have = w->mvp0_top[i - 8];
        else if (i < 15) have = cl[i - 13];
        else have = 1;
        if ((int32_t)mr[ME_KEY + i] != have) ok = 0;
    }
#if H264_DEVICE
    ok = __all_sync(0xffffffffu, ok);
#endif
    return ok;
}

/* ------------------------------------------------------------------------------
 * a17: encode one macroblock (mb_encode H:5724 + the pixel/coefficient half of
 * mb_write H:4378) with the whole CTA.  cl[] = rounded mv_clusters candidates.
 * Writes the MB's record, quantised levels, unfiltered reconstruction and *spec_out
 * (identical in every thread).
 *
 *   all      : load inputs, neighbours, MV context; load the search window
 *   warp 0   : skip test + start candidates, then the 16x16 search
 *   warp 1   : 16x8 and 8x16 searches      } start when warp 0 has published the
 *   warp 2   : 8x8 search                  } candidate stage (named barrier 1)
 *   warp 3   : Intra16x16 + Intra4x4 (stops early when warp 0 decides "skip")
 *   all      : mode decision (same strict comparisons and order as the reference)
 *   warps 0,1: luma halves;  warps 2,3: chroma planes (prediction + transform)
 *   all      : coded block pattern, skip rollback, record
 * ---------------------------------------------------------------------------- */
HDF_encode_mb void encode_mb(const FrameParams *fp, MBWork *w, int mbx, int mby, const int32_t cl[2], MBSpec *spec_out, int with_intra = 1)
{
    MBState s;
    s.fp = fp; s.w = w; s.mbx = mbx; s.mby = mby;
    s.avail = mb_avail(mbx, mby, fp->nmbx);
    s.type = 0; s.cost = 0x7FFFFFFF; s.i16_mode = 2; s.mv_skip_pred = 0;
    s.pbest = w->skip_pred; s.ss = &w->ss[SS_SLOT(WARP_ID < 3 ? WARP_ID : 0)];
    s.win_ok = 0; s.win_x0 = s.win_y0 = 0;
    s.map = 0; s.lut = 0;
    const int is_p = fp->slice_type == SLICE_P;
    MBInfo *mi = fp->mbi + mby * fp->nmbx + mbx;
    int32_t pmv[4] = {0, 0, 0, 0}, pmvd[4] = {0, 0, 0, 0};
    PROF_INIT(s);

    mb_load(s);
    PROF_MARK(s, 0);
    if (is_p)
    {
        int mvp16 = mvp_get(w->mvp0_left, w->mvp0_tl, w->mvp0_top, s.avail, 0, 0, 4, 4);
        /* searches by look-up (h264_sadmap.h) when the macroblock's record covers where its predictor points; else the
         * pixel flavour with its shared-memory window.  Same decisions either way. */
        s.lut = lut_decide(s, mvp16);
        LUT_STAT(0); if (s.lut) LUT_STAT(1);
        if (!s.lut) win_load(s, mbx * 16 + ((mv_x(mvp16) + 1) >> 2), mby * 16 + ((mv_y(mvp16) + 1) >> 2));
    }
    PROF_MARK(s, 1);

    /* ---- motion estimation: already done ahead of the wavefront?  (h264_wave.h, me_prepass_mb) ---- */
    int me_hit = 0;
    if (is_p && s.lut && fp->use_me)
    {
        const uint32_t *mr = s.map + SM_ME_OFF;
        me_hit = me_record_matches(s, mr, cl);
        if (me_hit)
        {
            FOR_THREADS(i, 52)
            {
                if (i < 16) w->ic[i] = (int32_t)mr[ME_IC + i];
                else if (i < 20) w->mode_cost[i - 16] = (int32_t)mr[ME_COST + i - 16];
                else if (i < 36) (&w->part_mv[0][0])[i - 20] = (int32_t)mr[ME_MV + i - 20];
                else (&w->part_mvd[0][0])[i - 36] = (int32_t)mr[ME_MVD + i - 36];
            }
            CTA_SYNC();
        }
        LUT_STAT(4); if (me_hit) LUT_STAT(5);
    }
    /* Intra modes inside this sweep?  Always, unless the sweep speculates that none wins (with_intra == 0, sweep 0 of a P
     * frame); even then when the inter cost is already known (record) and so far above the usual that an intra mode
     * probably does win -- evaluating it here saves the repair -- and when the estimation has to run in place anyway. */
    int do_intra = with_intra;
    if (!do_intra)
    {
        do_intra = 1;
        if (me_hit)
        {
            int t_, c_, u_; int32_t m_[4], d_[4];
            inter_decide(fp, w, &t_, &c_, m_, d_, &u_);
            do_intra = t_ != MBT_SKIP && c_ >= (fp->have_cost_stat ? fp->cost_stat[2 + mby] : 0);
        }
    }
    /* ---- concurrent tasks ---- */
    if (is_p && !me_hit) me_phase(s, cl);
    if (!do_intra)
    {   /* sweep 0 of a P frame speculates "no intra mode wins" (h264_wave.h, wave_mb_intra_check verifies afterwards) */
        ON_WARP(3) { IF_LANE0 { w->intra_res[0] = 0x7FFFFFFF; w->intra_res[1] = 2; w->intra_res[2] = 0x7FFFFFFF; w->intra_res[3] = 0; } }
    } else
    ON_WARP(3)
    {
        /* Intra16x16: heuristic mode, one prediction, SAD cost (intra_choose_16x16 H:4876) */
        const pix_t *left = (s.avail & AVAIL_L) ? w->left_y : 0;
        const pix_t *top = (s.avail & AVAIL_T) ? w->top_y : 0;
        int m16 = intra16_estimate(w->inp_y, s.avail, fp->qp);
        intra16_pred(w->i16pred, left, top, m16);
        WSYNC();
        int cost16 = sad_sm_wh(w->inp_y, w->i16pred, 16, 16)
                   + ((bitsize_ue(m16 + 1) * fp->lambda_q4) >> 4) + fp->lambda_i16_q4;
        int cost4 = 0x7FFFFFFF, nz4 = 0;
        if (fp->speed < 2 || !is_p) cost4 = intra4_choose(s, &nz4, cost16);
        IF_LANE0 { w->intra_res[0] = cost16; w->intra_res[1] = m16; w->intra_res[2] = cost4; w->intra_res[3] = nz4; }
#if H264_DEVICE
        /* the intra decision is usually over early (pruning): help with the partition-mode searches */
        if (is_p)
        {
            int st;
            while ((st = *(volatile int32_t *)&w->ic[IC_STATE]) == 0) { }
            __threadfence_block();
            if (st == 2)
            {
                if (!me_hit) partition_tasks(s, 3);
                /* P16x16 wins most of the time: when this warp has nothing else to do it prepares that
                 * mode's chroma prediction, taking the chroma reference fetch off the critical tail */
                if (*(volatile int32_t *)&w->ic[IC_COST0] >= 0)      /* only when the 16x16 search is already over: never wait for it */
                {
                    __threadfence_block();
                    /* both planes in one pass, every load in flight together (same arithmetic as
                     * interp_chroma_block for one 8x8 partition) */
                    const int32_t mv0 = *(volatile int32_t *)&w->part_mv[0][0];
                    const int scs = fp->stride[1];
                    const int ax = mv_x(mv0) + mbx * 64, ay = mv_y(mv0) + mby * 64, dx = ax & 7, dy = ay & 7;
                    const int ca = (8 - dx) * (8 - dy), cb = dx * (8 - dy), cc = (8 - dx) * dy, cd = dx * dy;
                    const long o = (long)(ay >> 3) * scs + (ax >> 3);
                    for (int i = LANE_ID; i < 128; i += 32)
                    {
                        const int pl = i >> 6, k = i & 63, r = k >> 3, x = k & 7;
                        const pix_t *p = fp->ref[1 + pl] + o + r * scs + x;
                        int v;
                        if (dx | dy) v = (ca * ldpx(p) + cb * ldpx(p + 1) + cc * ldpx(p + scs) + cd * ldpx(p + scs + 1) + 32) >> 6;
                        else v = ldpx(p);
                        w->predc[r * 16 + pl * 8 + x] = (pix_t)v;
                    }
                    __syncwarp();
                    IF_LANE0 { w->predc_mv = mv0; w->predc_tag = 1; }
                }
            }
        }
#endif
    }
    PROF_WARP(s, fp, mby * fp->nmbx + mbx, 12 + WARP_ID);
    CTA_SYNC();
    PROF_MARK(s, 5);

    /* ---- mode decision (every thread, same values) ---- */
    int nz_mask = 0, used_cl = 0;
    int32_t cand_sig[4] = {0, 0, 0, 0};
    spec_out->inter_best = 0;
    for (int k = 0; k < 4; k++) spec_out->mode_cost[k] = 0x7FFFFFFF;
    if (is_p)
    {
        s.mv_skip_pred = w->ic[IC_MV_SKIP];
        int use_skip_pred;
        const int best_type = inter_decide(fp, w, &s.type, &s.cost, pmv, pmvd, &use_skip_pred);
        if (w->ic[IC_STATE] != 1)
        {
            used_cl = 1;
            for (int k = 0; k < 4; k++) cand_sig[k] = w->ic[IC_SIG + k];
            for (int t = 0; t < 4; t++) if ((w->ic[IC_PREF] >> t) & 1) spec_out->mode_cost[t] = w->mode_cost[t];
            spec_out->inter_best = best_type;
        }
        /* pixel flavour: the prediction block the search (or the skip test) left behind; look-up flavour: made below */
        s.pbest = use_skip_pred ? w->skip_pred : (s.lut ? w->mode_store[0] : (pix_t *)w + w->mode_pred[best_type]);
    }
    spec_out->pad[0] = s.cost;           /* cost of the inter decision: what an intra mode has to beat (wave_mb_intra_check) */
    if (s.type >= 0)
    {
        s.i16_mode = w->intra_res[1];
        if (w->intra_res[0] < s.cost) { s.cost = w->intra_res[0]; s.type = MBT_I16; s.pbest = w->i16pred; }
        if (w->intra_res[2] < s.cost) { s.cost = w->intra_res[2]; s.type = MBT_I4; nz_mask = w->intra_res[3]; }
    }

    spec_out->mv0 = pmv[0];
    spec_out->flags = ((is_p && s.type < 5) ? SPEC_UPDATES : 0) | (used_cl ? SPEC_USED_CL : 0) | (do_intra ? 0 : SPEC_NO_INTRA);
    spec_out->cl_used[0] = mv_round_fullpel(cl[0]); spec_out->cl_used[1] = mv_round_fullpel(cl[1]);
    for (int k = 0; k < 4; k++) spec_out->cand_sig[k] = cand_sig[k];

#if H264_DEVICE
    /* the searches are over: the window buffer is free for the next macroblock of the row, whose
     * MV predictor will most likely be this macroblock's vector */
    if (is_p && mbx + 1 < fp->nmbx && !s.lut)
    {
        const int mvn = s.type >= 5 ? 0 : (s.type <= 1 ? pmv[0] : pmv[1]);
        win_prefetch(fp, w, mbx + 1, mby, (mbx + 1) * 16 + ((mv_x(mvn) + 1) >> 2), mby * 16 + ((mv_y(mvn) + 1) >> 2));
    }
#endif
    /* ---- prediction of chroma, transform, quantisation, reconstruction ---- */
    PROF_MARK(s, 6);
    if (s.lut && s.type <= 3)
    {   /* look-up flavour: the luma prediction of the chosen inter mode (or of the skip vector) is made now, from the
         * final vectors; each of the two luma warps makes the half it transforms */
        ON_WARP(0) { luma_pred_half(s, 0, s.type, pmv, w->mode_store[0]); }
        ON_WARP(1) { luma_pred_half(s, 1, s.type, pmv, w->mode_store[0]); }
        if (me_hit && s.type == MBT_SKIP)
        {   /* early skip taken from the record: the chroma prediction the skip test would have left in predc */
            ON_WARP(2) { mc_chroma_plane(s, 0, MBT_SKIP, pmv); }
            ON_WARP(3) { mc_chroma_plane(s, 1, MBT_SKIP, pmv); }
        }
        s.pbest = w->mode_store[0];
        if (s.type == MBT_SKIP) CTA_SYNC();
    }
    int cbpl = 0, cbpc = 0;
    const int sy = fp->stride[0], sc = fp->stride[1];
    pix_t *decy = fp->dec[0] + (mby * 16) * sy + mbx * 16;
    pix_t *du = fp->dec[1] + (mby * 8) * sc + mbx * 8, *dv = fp->dec[2] + (mby * 8) * sc + mbx * 8;
    if (s.type != MBT_SKIP)
    {
        const int i16 = s.type == MBT_I16;
#if H264_DEVICE
        if (s.type != MBT_I4)
        {
            if (WARP_ID < 2) luma_tq_fast(s, WARP_ID, i16);
        } else
        {
            ON_WARP(0)
            {
                FOR_LANES(i, 64) { int r = i >> 2, c = (i & 3) * 4; *(uint32_t *)(decy + r * sy + c) = ld4_sm(w->i4rec + r * 16 + c); }
            }
        }
        const int have_predc = s.type == 0 && w->predc_tag && w->predc_mv == pmv[0];
        ON_WARP(2)
        {
            if (s.type >= 5) intra_chroma_plane(s, 0); else if (!have_predc) mc_chroma_plane(s, 0, s.type, pmv);
            chroma_tq_fast(s, 0);
        }
        ON_WARP(3)
        {
            if (s.type >= 5) intra_chroma_plane(s, 1); else if (!have_predc) mc_chroma_plane(s, 1, s.type, pmv);
            chroma_tq_fast(s, 1);
        }
#else
        if (s.type != MBT_I4)
        {
            ON_WARP(0) { luma_tq_half(s, 0, i16, 0); }
            ON_WARP(1) { luma_tq_half(s, 1, i16, 0); }
            if (i16)
            {   /* the DC transform needs all 16 blocks: warps 0 and 1 meet at named barriers */
                if (WARP_ID < 2) bar_sync(2, 64);
                ON_WARP(0) { luma_tq_half(s, 0, 1, 1); }
                if (WARP_ID < 2) bar_sync(3, 64);
            }
            ON_WARP(0) { luma_tq_half(s, 0, i16, 2); }
            ON_WARP(1) { luma_tq_half(s, 1, i16, 2); }
        } else
        {
            ON_WARP(0)
            {
                FOR_LANES(i, 64) { int r = i >> 2, c = (i & 3) * 4; *(uint32_t *)(decy + r * sy + c) = ld4_sm(w->i4rec + r * 16 + c); }
            }
        }
        ON_WARP(2)
        {
            if (s.type >= 5) intra_chroma_plane(s, 0); else mc_chroma_plane(s, 0, s.type, pmv);
            chroma_tq_plane(s, 0);
        }
        ON_WARP(3)
        {
            if (s.type >= 5) intra_chroma_plane(s, 1); else mc_chroma_plane(s, 1, s.type, pmv);
            chroma_tq_plane(s, 1);
        }
#endif
        if (WARP_ID >= 2) PROF_WARP(s, fp, mby * fp->nmbx + mbx, 15 + WARP_ID);
        PROF_MARK(s, 7);
        CTA_SYNC();
        if (s.type != MBT_I4) nz_mask = w->tq_res[0] | w->tq_res[1];
        if (nz_mask & 0xCC00) cbpl |= 1;
        if (nz_mask & 0x3300) cbpl |= 2;
        if (nz_mask & 0x00CC) cbpl |= 4;
        if (nz_mask & 0x0033) cbpl |= 8;
        for (int pl = 0; pl < 2; pl++)
        {
            if (w->tq_res[2 + pl] & 0xFF) cbpc = 2;
            cbpc |= w->tq_res[2 + pl] >> 8;
        }
        cbpc = imin(cbpc, 2);
        if (!(s.type | cbpl | cbpc) && pmv[0] == s.mv_skip_pred) s.type = MBT_SKIP;   /* rollback H:4494 */
    }
    PROF_MARK(s, 10);
    if (s.type == MBT_SKIP)
    {
        /* reconstruction = prediction (H:4417-4420); for the early skip predc holds the chroma
         * prediction of the skip test, for a rollback the chroma prediction just computed */
        FOR_THREADS(i, 96)
        {
            if (i < 64) { int r = i >> 2, c = (i & 3) * 4; *(uint32_t *)(decy + r * sy + c) = ld4_sm(s.pbest + r * 16 + c); }
            else
            {
                int k = i - 64, r = k >> 2, q = k & 3;
                pix_t *d = (q < 2 ? du : dv) + r * sc + (q & 1) * 4;
                *(uint32_t *)d = ld4_sm(w->predc + r * 16 + q * 4);
            }
        }
        nz_mask = 0; cbpl = cbpc = 0;
    }

    /* ---- macroblock record ---- */
    const int type = s.type;
    if (type == MBT_I16 && cbpl) cbpl = 15;
    FOR_THREADS(i, 32)
    {
        if (i < 16)
        {
            int v;
            if (type >= 5) v = MV_NA;
            else if (type <= 0) v = pmv[0];
            else
            {
                int bx = i & 3, by = i >> 2;
                int part = type == 1 ? (by >> 1) : (type == 2 ? (bx >> 1) : (by >> 1) * 2 + (bx >> 1));
                v = pmv[part];
            }
            mi->mv[i] = v;
            mi->i4_mode[i] = type == MBT_I4 ? w->i4_mode[i] : 2;
            mi->i4_code[i] = type == MBT_I4 ? w->i4_code[i] : 0;
            /* total_coeff as the neighbours' CAVLC context will see it (H:4630-4647) */
            int grp = (i >> 3) * 2 + ((i & 3) >> 1);
            int n = 0;
            if (type != MBT_SKIP && (cbpl & (1 << grp))) n = count_nz(w->qv_y[i], type == MBT_I16);
            mi->nnz[i] = (uint8_t)n;
        } else if (i < 24)
        {
            int b = i - 16;
            mi->nnz[i] = (uint8_t)((type != MBT_SKIP && cbpc == 2) ? count_nz(w->qv_c[b], 1) : 0);
        } else if (i < 28)
        {
            mi->mvd[i - 24] = pmvd[i - 24];
        } else if (i == 28)
        {
            mi->type = (int8_t)type;
            mi->i16_mode = (int8_t)s.i16_mode;
            mi->cbp = (uint8_t)(cbpl | (cbpc << 4));
            mi->flags = 0;
            mi->nz_mask = (uint16_t)(type == MBT_SKIP ? 0 : nz_mask);
            mi->pad0 = 0;
        }
    }
    /* the quantised levels are only read by the CAVLC pass: mb_store_coefs() writes them out after
     * the macroblock has been published to the wavefront */
    IF_THREAD0 { w->scal[8] = type; w->scal[9] = 1 + mby * fp->nmbx + mbx; }
    CTA_SYNC();
    PROF_MARK(s, 11);
    PROF_STORE(s, fp, mby * fp->nmbx + mbx, type);
}

/* Second half of the macroblock record: quantised levels -> HBM (for the CAVLC pass).  Called by
 * the whole CTA after encode_mb(), once the wavefront counters have been advanced; a no-op when
 * no macroblock has been encoded since the last call. */
HDF_mb_store_coefs void mb_store_coefs(const FrameParams *fp, MBWork *w)
{
    const int tag = w->scal[9];
    if (!tag) return;
    const int type = w->scal[8];
    int16_t *coef = fp->coef + (size_t)(tag - 1) * COEF_PER_MB;
    if (type != MBT_SKIP)
    {
        FOR_THREADS(i, COEF_PER_MB / 2)      /* two int16 per item */
        {
            int k = i * 2;
            uint32_t v;
            if (k < COEF_YDC) v = *(const uint32_t *)(&w->qv_y[0][0] + k);
            else if (k < COEF_C) v = *(const uint32_t *)(w->qdc_y + (k - COEF_YDC));
            else if (k < COEF_CDC) v = *(const uint32_t *)(&w->qv_c[0][0] + (k - COEF_C));
            else v = *(const uint32_t *)(w->qdc_c + (k - COEF_CDC));
            *(uint32_t *)(coef + k) = v;
        }
    }
    CTA_SYNC();
    IF_THREAD0 { w->scal[9] = 0; }
    CTA_SYNC();
}
