/*
 * h264_deblock.h -- in-loop deblocking of one macroblock by one warp
 * (df_strength H:5535, mb_deblock H:5642,
 * h264e_deblock_luma H:1505, h264e_deblock_chroma H:1469), then border extension
 * (a16, h264e_copy_borders H:2232).
 *
 * Lanes 0-15 own the 16 luma lines crossing the edges being filtered, lanes 16-23
 * the 8 U lines and lanes 24-31 the 8 V lines.  The four vertical edges are filtered
 * left to right on each line held in registers, then the four horizontal edges top
 * to bottom, exactly the per-macroblock order of the reference; macroblocks follow
 * an x+2y wavefront so that every neighbour a filter touches is final.
 */
#pragma once
#include "h264_common.h"

HD int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* one line across a luma edge: p[3..0] = p3..p0 at v[o-4..o-1], q0..q3 at v[o..o+3] */
HD void df_luma_line(int *v, int o, int bs, int alpha, int beta, int tc0)
{
    int p2 = v[o - 3], p1 = v[o - 2], p0 = v[o - 1], q0 = v[o], q1 = v[o + 1], q2 = v[o + 2];
    if (!bs) return;
    if (!(iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta)) return;
    int ap = iabs(p2 - p0), aq = iabs(q2 - q0);
    if (bs < 4)
    {
        int delta = (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3;
        int tc = tc0;
        if (ap < beta) { v[o - 2] = p1 + clip3(-tc0, tc0, ((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1); tc++; }
        if (aq < beta) { v[o + 1] = q1 + clip3(-tc0, tc0, ((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1); tc++; }
        delta = clip3(-tc, tc, delta);
        v[o - 1] = clip_u8(p0 + delta);
        v[o] = clip_u8(q0 - delta);
    } else
    {
        int small = iabs(p0 - q0) < ((alpha >> 2) + 2);
        if (ap < beta && small)
        {
            int p3 = v[o - 4];
            v[o - 1] = (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3;
            v[o - 2] = (p2 + p1 + p0 + q0 + 2) >> 2;
            v[o - 3] = (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3;
        } else v[o - 1] = (2 * p1 + p0 + q1 + 2) >> 2;
        if (aq < beta && small)
        {
            int q3 = v[o + 3];
            v[o] = (q2 + 2 * q1 + 2 * q0 + 2 * p0 + p1 + 4) >> 3;
            v[o + 1] = (q2 + q1 + p0 + q0 + 2) >> 2;
            v[o + 2] = (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3;
        } else v[o] = (2 * q1 + q0 + p1 + 2) >> 2;
    }
}

/* one line across a chroma edge (deblock_chroma H:1217) */
HD void df_chroma_line(int *v, int o, int bs, int alpha, int beta, int tc0)
{
    int p1 = v[o - 2], p0 = v[o - 1], q0 = v[o], q1 = v[o + 1];
    if (!bs) return;
    if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
    if (bs < 4)
    {
        int tc = tc0 + 1;
        int delta = clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
        v[o - 1] = clip_u8(p0 + delta);
        v[o] = clip_u8(q0 - delta);
    } else
    {
        v[o - 1] = (2 * p1 + p0 + q1 + 2) >> 2;
        v[o] = (2 * q1 + q0 + p1 + 2) >> 2;
    }
}

HD int mv_far(int a, int b) { return iabs(mv_x(a) - mv_x(b)) > 3 || iabs(mv_y(a) - mv_y(b)) > 3; }   /* H:3466 */

/* boundary strength between 4x4 block p (in MB mp, index ip) and q (in MB mq, index iq),
 * both inter (df_strength H:5594-5611) */
HD int bs_inter(const MBInfo *mp, int ip, const MBInfo *mq, int iq)
{
    if (((mp->nz_mask | 0u) & (0x8000u >> ip)) || (mq->nz_mask & (0x8000u >> iq))) return 2;
    return mv_far(mp->mv[ip], mq->mv[iq]) ? 1 : 0;
}

/* working tile of one macroblock being deblocked (shared memory) */
struct DeblockTile
{
    uint32_t y[20 * 6];      /* luma: rows -4..15, cols -4..19 (6 words per row, sample x at col x + 4) */
    uint32_t c[2][12 * 3];   /* chroma: rows -4..7, cols -4..7                                          */
    uint8_t bs[32];          /* [4*e + seg] vertical edges, [16 + 4*e + seg] horizontal edges          */
};

/* One macroblock, one warp, luma (part 0) or both chroma planes (part 1) -- the two are
 * independent and run as separate wavefronts.  The samples the filters can touch (the macroblock,
 * 4 columns of the left and 4 rows of the upper neighbour) are staged as aligned words in shared
 * memory, filtered there (one lane per line crossing the edges, held in registers; vertical
 * edges left to right, then horizontal edges top to bottom -- the reference's per-macroblock
 * order) and written back as words. */
/* boundary strength (df_strength H:5535) of item j of macroblock (mbx, mby): j -> edge e, 4-sample segment seg.
 * Everything an item may need is fetched up front with independent loads (one memory round trip instead of a
 * chain of dependent ones), the decision is taken afterwards. */
HD int deblock_bs_item(const FrameParams *fp, int mbx, int mby, int j)
{
    const MBInfo *mi = fp->mbi + mby * fp->nmbx + mbx;
    const MBInfo *ml = mi - 1, *mt = mi - fp->nmbx;
    const int horiz = j >> 4, e = (j >> 2) & 3, seg = j & 3;
    const int at_border = (horiz ? mby : mbx) == 0;
    /* p side: the neighbouring macroblock for edge 0 (this macroblock again at a picture border: unused) */
    const MBInfo *mp = e ? mi : (at_border ? mi : (horiz ? mt : ml));
    const int ip = e ? (horiz ? (e - 1) * 4 + seg : seg * 4 + e - 1) : (horiz ? 12 + seg : seg * 4 + 3);
    const int iq = horiz ? e * 4 + seg : seg * 4 + e;
    const int type_q = mi->type, type_p = mp->type;
    const unsigned nz_p = mp->nz_mask, nz_q = mi->nz_mask;
    const int mv_p = mp->mv[ip], mv_q = mi->mv[iq];
    const int intra = type_q >= 5;
    int bs;
    if (e == 0)
    {
        if (at_border) bs = 0;
        else if (intra || type_p >= 5) bs = 4;
        else bs = ((nz_p & (0x8000u >> ip)) || (nz_q & (0x8000u >> iq))) ? 2 : (mv_far(mv_p, mv_q) ? 1 : 0);
    } else if (intra) bs = 3;
    else bs = ((nz_p & (0x8000u >> ip)) || (nz_q & (0x8000u >> iq))) ? 2 : (mv_far(mv_p, mv_q) ? 1 : 0);
    return bs;
}
HD void deblock_bs(const FrameParams *fp, DeblockTile *t, int mbx, int mby)
{
    FOR_LANES(j, 32) { t->bs[j] = (uint8_t)deblock_bs_item(fp, mbx, mby, j); }
}

/* phase 0: everything that does not depend on the row above -- the macroblock's own rows (with the 4
 * columns to their left, final since the previous macroblock of this row) and the boundary strengths;
 * phase 1 (once the row above has got far enough): the 4 rows above, the filters, the write-back. */
#if H264_DEVICE
HD void df_cp_async4(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
HD void df_cp_async_wait()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
}
/* GPU software pipeline (k_deblock_rows): while macroblock x is being filtered, the samples of macroblock x + 1
 * that nothing can touch before its own filtering (its 16x16 / 8x8 samples, not the 4 columns to its left) are
 * already on their way into the other tile. */
HD void deblock_prefetch(const FrameParams *fp, DeblockTile *t, int mbx, int mby, int part)
{
    const int sy = fp->stride[0], sc = fp->stride[1];
    if (part == 0)
    {
        const pix_t *py = fp->dec[0] + (mby * 16) * sy + mbx * 16;
        FOR_LANES(i, 64) { int r = i >> 2, c = 1 + (i & 3); df_cp_async4(&t->y[(r + 4) * 6 + c], py + r * sy + c * 4 - 4); }
    } else
    {
        FOR_LANES(k, 32)
        {
            int pl = k >> 4, j = k & 15, r = j >> 1, c = 1 + (j & 1);
            const pix_t *pc = fp->dec[1 + pl] + (mby * 8) * sc + mbx * 8;
            df_cp_async4(&t->c[pl][(r + 4) * 3 + c], pc + r * sc + c * 4 - 4);
        }
    }
    /* its boundary strengths: k_deblock_rows (deblock_bs_item before, the store after the current macroblock's filters) */
}
/* the 4 columns to the left of macroblock x + 1 = the last 4 columns of macroblock x, final in tile `cur` */
HD void deblock_handover(const DeblockTile *cur, DeblockTile *nxt, int part)
{
    if (part == 0) { FOR_LANES(r, 16) nxt->y[(r + 4) * 6] = cur->y[(r + 4) * 6 + 4]; }
    else { FOR_LANES(k, 16) { int pl = k >> 3, r = k & 7; nxt->c[pl][(r + 4) * 3] = cur->c[pl][(r + 4) * 3 + 2]; } }
    df_cp_async_wait();
    WSYNC();
}
#endif

HD void deblock_mb(const FrameParams *fp, DeblockTile *t, int mbx, int mby, int part, int phase)
{
    const int sy = fp->stride[0], sc = fp->stride[1];
    if (part == 0)
    {
        pix_t *py = fp->dec[0] + (mby * 16) * sy + mbx * 16;
        if (phase == 0)
        {
            uint32_t v[3];
#pragma unroll
            for (int k = 0; k < 3; k++) { int i = LANE_ID + 32 * k; if (i < 80) { int r = i / 5, c = i - r * 5; v[k] = *(const uint32_t *)(py + r * sy + c * 4 - 4); } }
#pragma unroll
            for (int k = 0; k < 3; k++) { int i = LANE_ID + 32 * k; if (i < 80) { int r = i / 5, c = i - r * 5; t->y[(r + 4) * 6 + c] = v[k]; } }
#if !H264_DEVICE
            for (int i = 0; i < 80; i++) { int r = i / 5, c = i - r * 5; t->y[(r + 4) * 6 + c] = *(const uint32_t *)(py + r * sy + c * 4 - 4); }
#endif
            deblock_bs(fp, t, mbx, mby);
            return;
        }
        const int alpha = fp->df_alpha[0], beta = fp->df_beta[0];
        int tc0[4];
        for (int k = 0; k < 4; k++) tc0[k] = fp->df_tc0[0][k];
#if H264_DEVICE
        /* the 4 rows above are only needed by the horizontal edges: their loads fly while the vertical edges are filtered */
        uint32_t above = 0;
        { const int i = LANE_ID; if (i < 20) { int r = i / 5, c = i - r * 5; above = *(const uint32_t *)(py + (r - 4) * sy + c * 4 - 4); } }
#else
        FOR_LANES(i, 20) { int r = i / 5, c = i - r * 5; t->y[r * 6 + c] = *(const uint32_t *)(py + (r - 4) * sy + c * 4 - 4); }
#endif
        WSYNC();
        FOR_LANES(ln, 16)      /* vertical edges: lane owns one row */
        {
            uint32_t *row = t->y + (ln + 4) * 6;
            int v[20];
#pragma unroll
            for (int k = 0; k < 5; k++) unpack4(row[k], v + 4 * k);
#pragma unroll
            for (int e = 0; e < 4; e++) { const int bs = t->bs[4 * e + (ln >> 2)]; df_luma_line(v, 4 + 4 * e, bs, alpha, beta, tc0[bs & 3]); }
#pragma unroll
            for (int k = 0; k < 5; k++) row[k] = pack4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
#if H264_DEVICE
        { const int i = LANE_ID; if (i < 20) { int r = i / 5, c = i - r * 5; t->y[r * 6 + c] = above; } }
#endif
        WSYNC();
        FOR_LANES(ln, 16)      /* horizontal edges: lane owns one column */
        {
            pix_t *col = (pix_t *)t->y + ln + 4;
            int v[20];
#pragma unroll
            for (int i = 0; i < 20; i++) v[i] = col[i * 24];
#pragma unroll
            for (int e = 0; e < 4; e++) { const int bs = t->bs[16 + 4 * e + (ln >> 2)]; df_luma_line(v, 4 + 4 * e, bs, alpha, beta, tc0[bs & 3]); }
#pragma unroll
            for (int i = 1; i < 19; i++) col[i * 24] = (pix_t)v[i];
        }
        WSYNC();
        FOR_LANES(i, 100)
        {
            int r = i / 5, c = i - r * 5;
            if ((r >= 4 || (mby > 0 && c > 0)) && (c > 0 || (mbx > 0 && r >= 4)))
                *(uint32_t *)(py + (r - 4) * sy + c * 4 - 4) = t->y[r * 6 + c];
        }
    } else
    {
        pix_t *pc[2];
        pc[0] = fp->dec[1] + (mby * 8) * sc + mbx * 8;
        pc[1] = fp->dec[2] + (mby * 8) * sc + mbx * 8;
        if (phase == 0)
        {
            FOR_LANES(k, 48) { int pl = k / 24, j = k - pl * 24, r = j / 3, c = j - r * 3; t->c[pl][(r + 4) * 3 + c] = *(const uint32_t *)(pc[pl] + r * sc + c * 4 - 4); }
            deblock_bs(fp, t, mbx, mby);
            return;
        }
        const int alpha = fp->df_alpha[1], beta = fp->df_beta[1];
        int tc0[4];
        for (int k = 0; k < 4; k++) tc0[k] = fp->df_tc0[1][k];
        FOR_LANES(k, 24) { int pl = k / 12, j = k - pl * 12, r = j / 3, c = j - r * 3; t->c[pl][r * 3 + c] = *(const uint32_t *)(pc[pl] + (r - 4) * sc + c * 4 - 4); }
        WSYNC();
        FOR_LANES(ln, 16)      /* lanes 0-7: U lines, 8-15: V lines */
        {
            const int pl = ln >> 3, line = ln & 7;
            uint32_t *row = t->c[pl] + (line + 4) * 3;
            int v[12];
#pragma unroll
            for (int k = 0; k < 3; k++) unpack4(row[k], v + 4 * k);
#pragma unroll
            for (int e = 0; e < 4; e += 2) { const int bs = t->bs[4 * e + (line >> 1)]; df_chroma_line(v, 4 + 2 * e, bs, alpha, beta, tc0[bs & 3]); }
#pragma unroll
            for (int k = 0; k < 3; k++) row[k] = pack4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
        WSYNC();
        FOR_LANES(ln, 16)
        {
            const int pl = ln >> 3, line = ln & 7;
            pix_t *col = (pix_t *)t->c[pl] + line + 4;
            int v[12];
#pragma unroll
            for (int i = 0; i < 12; i++) v[i] = col[i * 12];
#pragma unroll
            for (int e = 0; e < 4; e += 2) { const int bs = t->bs[16 + 4 * e + (line >> 1)]; df_chroma_line(v, 4 + 2 * e, bs, alpha, beta, tc0[bs & 3]); }
#pragma unroll
            for (int i = 2; i < 10; i++) col[i * 12] = (pix_t)v[i];
        }
        WSYNC();
        FOR_LANES(k, 72)
        {
            int pl = k / 36, j = k - pl * 36, r = j / 3, c = j - r * 3;
            if ((r >= 4 || (mby > 0 && c > 0)) && (c > 0 || (mbx > 0 && r >= 4)))
                *(uint32_t *)(pc[pl] + (r - 4) * sc + c * 4 - 4) = t->c[pl][r * 3 + c];
        }
    }
    WSYNC();
}



/* a16: replicate the picture edges into the 16 (luma) / 8 (chroma) sample guard band.
 * One work item per guard sample; idx enumerates the guard samples of plane pl. */
HD void extend_border_sample(const FrameParams *fp, int pl, long idx)
{
    const int cr = pl != 0;
    const int g = cr ? 8 : 16;
    const int w = fp->nmbx * (cr ? 8 : 16), h = fp->nmby * (cr ? 8 : 16);
    const int stride = fp->stride[cr];
    pix_t *pic = fp->dec[pl];
    const long side = (long)2 * g * h;        /* left+right strips of the picture rows */
    int x, y;
    if (idx < side)
    {
        y = (int)(idx / (2 * g));
        int k = (int)(idx - (long)y * 2 * g);
        x = k < g ? k - g : w + (k - g);
    } else
    {
        long r = idx - side;
        int rw = w + 2 * g;
        int yy = (int)(r / rw);
        x = (int)(r - (long)yy * rw) - g;
        y = yy < g ? yy - g : h + (yy - g);
    }
    int sx = x < 0 ? 0 : (x >= w ? w - 1 : x), sy = y < 0 ? 0 : (y >= h ? h - 1 : y);
    pic[(long)y * stride + x] = pic[(long)sy * stride + sx];
}
HD long border_samples(const FrameParams *fp, int pl)
{
    const int cr = pl != 0;
    const int g = cr ? 8 : 16;
    const int w = fp->nmbx * (cr ? 8 : 16), h = fp->nmby * (cr ? 8 : 16);
    return (long)2 * g * h + (long)2 * g * (w + 2 * g);
}


/* half-sample planes of the finished picture (h264_pixel.h hpel_word): word wi of the padded luma plane */
HD void hpel_plane_word(const FrameParams *fp, long wi)
{
    uint32_t h, v, d;
    const long nwords = fp->luma_bytes >> 2;
    hpel_word(fp->dec_base, nwords, fp->stride[0], wi, &h, &v, &d);
    uint32_t *o = (uint32_t *)fp->hp_out;
    o[wi] = h; o[nwords + wi] = v; o[2 * nwords + wi] = d;
}
