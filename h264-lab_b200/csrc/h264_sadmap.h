/*
 * h264_sadmap.h -- SAD maps: the pixel work of motion estimation taken OFF the macroblock wavefront.
 *
 * The reference's motion search (me_search_diamond H:4973, the candidate stage of inter_choose_mode H:5360-5412, the
 * skip test H:5306-5350) is a serial, data-dependent walk, but every number it consumes is a pure function of
 * (input macroblock, reference picture, position):  SAD(block at motion vector v).  The walk depends on the
 * neighbours' final vectors (MV predictor, candidates) and so has to stay on the x+2y wavefront; the SADs do not.
 * A dependency-free pre-pass (k_sadmap, one CTA per macroblock, one THREAD per position -- no shuffles, no
 * reductions, every lane busy) therefore tabulates, for every macroblock of every P frame of the submission,
 *
 *   the integer map   the four 8x8-quadrant SADs of the 16x16 block at every full-sample offset within +-SM_R of a
 *                     centre (the co-located vector of the previous frame, rounded) -- 16x16, 16x8, 8x16 and 8x8
 *                     partition SADs are sums of quadrants (h264e_sad_mb_unlaign_8x8 H:2178);
 *   the quarter map   the same four numbers at every quarter-sample position within +-SM_QR quarter samples of the
 *                     integer position with the smallest 16x16 SAD: every sub-sample probe of the reference's
 *                     refinement (H:5083-5174) is a standard quarter-sample position, i.e. one sample of one of the
 *                     planes {G, b, h, j} or the rounded average of two (interp_luma_planes, h264_pixel.h).
 *                     SM_QR = 2 covers the seven probes around that position; measured on the bench content, 93 % of all
 *                     sub-sample look-ups fall there, 0.6 % between 3 and 6 quarter samples away (a 13 x 13 map would
 *                     catch those at seven times the work) and the rest further away than any small map reaches.
 *
 * The wavefront kernel then replays the reference's search with table look-ups (a few instructions per probe,
 * maps staged into shared memory by one bulk-copy (TMA) instruction per macroblock) instead of warp-wide pixel loops.
 * A position outside the maps is computed from the pictures exactly as before -- the maps are a cache of exact
 * values, never an approximation, so the decisions are the reference's by construction; a macroblock whose
 * predictor points away from its map skips the look-ups altogether and runs the pixel path with its search window.
 *
 * Record of one macroblock, SM_WORDS 32-bit words (2336 bytes, a multiple of 16 for the bulk copy):
 *   [0] centre of the integer map, macroblock-relative full samples (x | y << 16)      [1] centre of the quarter map
 *   [2] 1 when the record is valid for the frame                                       [3] reserved
 *   [SM_INT_OFF + 2 k]     q0 | q1 << 16     k = (dy + SM_R) * SM_N + dx + SM_R        (q0 TL, q1 TR, q2 BL, q3 BR)
 *   [SM_INT_OFF + 2 k + 1] q2 | q3 << 16     0xFFFFFFFF in both: not tabulated (block not inside the padded picture)
 *   [SM_Q_OFF + ...]       same for k = (qy + SM_QR) * SM_QN + qx + SM_QR
 *   [SM_ME_OFF + ...]      ME_WORDS words: the speculative motion-estimation record (h264_wave.h me_prepass_mb)
 */
#pragma once
#include "h264_common.h"
#include "h264_pixel.h"

/* sum of the quadrants a bw x bh partition at (ppx, ppy) covers */
HD int quads_part(uint32_t lo, uint32_t hi, int ppx, int ppy, int bw, int bh)
{
    const int q0 = (int)(lo & 0xFFFF), q1 = (int)(lo >> 16), q2 = (int)(hi & 0xFFFF), q3 = (int)(hi >> 16);
    if (bw == 16) { if (bh == 16) return q0 + q1 + q2 + q3; return ppy ? q2 + q3 : q0 + q1; }
    if (bh == 16) return ppx ? q1 + q3 : q0 + q2;
    return ppy ? (ppx ? q3 : q2) : (ppx ? q1 : q0);
}

/* look-up of macroblock-relative quarter-sample vector (x, y) in a record: 1 = hit */
HD int sadmap_lookup(const uint32_t *map, int cx, int cy, int qcx, int qcy, int x, int y, uint32_t *lo, uint32_t *hi)
{
    if (!((x | y) & 3))
    {
        const int ix = (x >> 2) - cx + SM_R, iy = (y >> 2) - cy + SM_R;
#if !H264_DEVICE && defined(SADMAP_STATS)
        { extern long g_emu_ihist[16]; int m = imax(iabs((x >> 2) - cx), iabs((y >> 2) - cy)); g_emu_ihist[m > 15 ? 15 : m]++; }
#endif
        if ((unsigned)ix < (unsigned)SM_N && (unsigned)iy < (unsigned)SM_N)
        {
            const uint32_t *e = map + SM_INT_OFF + 2 * (iy * SM_N + ix);
            *lo = e[0]; *hi = e[1];
            if (*lo != SM_INVALID) return 1;
        }
    }
    const int qx = x - 4 * qcx + SM_QR, qy = y - 4 * qcy + SM_QR;
#if !H264_DEVICE && defined(SADMAP_STATS)
    { extern long g_emu_qhist[16]; int m = imax(iabs(x - 4 * qcx), iabs(y - 4 * qcy)); g_emu_qhist[m > 15 ? 15 : m]++; }
#endif
    if ((unsigned)qx < (unsigned)SM_QN && (unsigned)qy < (unsigned)SM_QN)
    {
        const uint32_t *e = map + SM_Q_OFF + 2 * (qy * SM_QN + qx);
        *lo = e[0]; *hi = e[1];
        return *lo != SM_INVALID;
    }
    return 0;
}

/* four samples of the prediction at ABSOLUTE quarter-sample position (ax, ay) (first sample of the word): the position
 * table of h264e_qpel_interpolate_luma (H:2079-2130) on the half-sample planes, per word -- same values as
 * interp_luma_planes() */
HD void interp_luma_sel(const pix_t *pg, const pix_t *pb0, const pix_t *ph0, const pix_t *pj0, int st, int ax, int ay,
                        const pix_t **pa, const pix_t **pb)
{
    /* branch-free: the lanes of a warp ask for different positions (partitions, probes), and one instruction stream
     * keeps all their loads in flight together.  Positions made of a single sample average it with itself.
     * The four planes are given by the address of their sample (0, 0) and a common stride (the pictures, or a staged
     * window of them). */
    const int dx = ax & 3, dy = ay & 3;
    const long o = (long)(ay >> 2) * st + (ax >> 2);
    const int pos = 1 << (dx + 4 * dy);
    const int m0 = (pos & 0xe0ee) != 0, m1 = (pos & 0xbbb0) != 0, m2 = (pos & 0x4e40) != 0, m3 = (pos & 0xfafa) != 0;
    const pix_t *g = pg + o;
    const pix_t *p0 = pb0 + o + ((pos & 0xe000) ? st : 0);
    const pix_t *p1 = ph0 + o + ((pos & 0x8880) ? 1 : 0);
    const pix_t *p2 = pj0 + o;
    const pix_t *p3 = g + ((dx + 1) >> 2) + ((dy + 1) >> 2) * st;
    const pix_t *a = m0 ? p0 : (m1 ? p1 : (m2 ? p2 : g));
    const pix_t *b = m0 ? (m1 ? p1 : (m2 ? p2 : (const pix_t *)0)) : ((m1 && m2) ? p2 : (const pix_t *)0);
    *pa = a;
    *pb = b ? b : (m3 ? p3 : a);
}
HD void interp_luma_ptrs(const FrameParams *fp, int ax, int ay, const pix_t **pa, const pix_t **pb)
{
    interp_luma_sel(fp->ref[0], fp->hp[0], fp->hp[1], fp->hp[2], fp->stride[0], ax, ay, pa, pb);
}
HD uint32_t interp_luma_word(const FrameParams *fp, int ax, int ay)
{
    const pix_t *a, *b;
    interp_luma_ptrs(fp, ax, ay, &a, &b);
    return avg4(ld4u(a), ld4u(b));
}

/* word c4 of row r of the input macroblock (mbx, mby); samples beyond the visible picture replicate the last column / row
 * (pix_copy_cropped_mb H:3536) */
HD uint32_t sadmap_inp_word(const FrameParams *fp, int mbx, int mby, int r, int c4)
{
    const int wv = fp->width, hv = fp->height, x = mbx * 16 + 4 * c4, y = mby * 16 + r;
    if (x + 4 <= wv && y < hv) return ld4u(fp->inp[0] + (long)y * fp->inp_stride[0] + x);
    const pix_t *row = fp->inp[0] + (long)imin(y, hv - 1) * fp->inp_stride[0];
    uint32_t v = 0;
    for (int q = 0; q < 4; q++) v |= (uint32_t)row[imin(x + q, wv - 1)] << (8 * q);
    return v;
}

/* centre of the integer map of macroblock n: where the co-located macroblock of the previous frame pointed (the record
 * array still holds the previous frame when the pre-pass runs), else (0, 0).  Any choice is exact; this one is cheap. */
HD void sadmap_center(const FrameParams *fp, int n, int *cx, int *cy)
{
    *cx = 0; *cy = 0;
    if (fp->spec_from_prev)        /* the previous frame of this session was a P frame */
    {
        const int mv = fp->mbi[n].mv[0];
        if (fp->mbi[n].type < 5 && mv_x(mv) != MV_NA) { *cx = (mv_x(mv) + 2) >> 2; *cy = (mv_y(mv) + 2) >> 2; }
    }
    /* keep the centre inside what a vector may be at all (H:6322-6325), so that the window stays near the picture */
    const int mbx = n % fp->nmbx, mby = n / fp->nmbx;
    *cx = imin(imax(*cx, -14 - mbx * 16), fp->nmbx * 16 - 2 - mbx * 16);
    *cy = imin(imax(*cy, -14 - mby * 16), fp->nmby * 16 - 2 - mby * 16);
}

/* a 16x16 block at absolute full-sample position (ax, ay), read with one extra column / row, lies inside the padded picture */
HD int sadmap_block_inside(const FrameParams *fp, int ax, int ay)
{
    return ax >= -16 && ay >= -16 && ax + 17 <= fp->nmbx * 16 + 16 && ay + 17 <= fp->nmby * 16 + 16;
}

#if !H264_DEVICE
/* Host emulation of the pre-pass for one macroblock: the definition of the record's contents in plain loops. */
static void sadmap_build_mb(const FrameParams *fp, int mbx, int mby)
{
    const int n = mby * fp->nmbx + mbx, st = fp->stride[0];
    uint32_t *rec = fp->sadmap + (size_t)n * SM_WORDS;
    uint32_t inp[64];
    for (int r = 0; r < 16; r++) for (int c = 0; c < 4; c++) inp[r * 4 + c] = sadmap_inp_word(fp, mbx, mby, r, c);
    int cx, cy;
    sadmap_center(fp, n, &cx, &cy);
    int best = 0x7FFFFFFF, bx = cx, by = cy;
    for (int dy = -SM_R; dy <= SM_R; dy++)
        for (int dx = -SM_R; dx <= SM_R; dx++)
        {
            const int ax = mbx * 16 + cx + dx, ay = mby * 16 + cy + dy;
            uint32_t lo = SM_INVALID, hi = SM_INVALID;
            if (sadmap_block_inside(fp, ax, ay))
            {
                int q[4] = {0, 0, 0, 0};
                for (int r = 0; r < 16; r++)
                    for (int c = 0; c < 4; c++)
                        q[(r >> 3) * 2 + (c >> 1)] += sad4(ld4u(fp->ref[0] + (long)(ay + r) * st + ax + 4 * c), inp[r * 4 + c]);
                lo = (uint32_t)q[0] | ((uint32_t)q[1] << 16); hi = (uint32_t)q[2] | ((uint32_t)q[3] << 16);
                const int tot = q[0] + q[1] + q[2] + q[3];
                if (tot < best) { best = tot; bx = cx + dx; by = cy + dy; }
            }
            uint32_t *e = rec + SM_INT_OFF + 2 * ((dy + SM_R) * SM_N + dx + SM_R);
            e[0] = lo; e[1] = hi;
        }
    for (int qy = -SM_QR; qy <= SM_QR; qy++)
        for (int qx = -SM_QR; qx <= SM_QR; qx++)
        {
            const int ax = (mbx * 16 + bx) * 4 + qx, ay = (mby * 16 + by) * 4 + qy;
            uint32_t lo = SM_INVALID, hi = SM_INVALID;
            if (sadmap_block_inside(fp, ax >> 2, ay >> 2))
            {
                int q[4] = {0, 0, 0, 0};
                for (int r = 0; r < 16; r++)
                    for (int c = 0; c < 4; c++)
                        q[(r >> 3) * 2 + (c >> 1)] += sad4(interp_luma_word(fp, ax + 16 * c, ay + 4 * r), inp[r * 4 + c]);
                lo = (uint32_t)q[0] | ((uint32_t)q[1] << 16); hi = (uint32_t)q[2] | ((uint32_t)q[3] << 16);
            }
            uint32_t *e = rec + SM_Q_OFF + 2 * ((qy + SM_QR) * SM_QN + qx + SM_QR);
            e[0] = lo; e[1] = hi;
        }
    rec[0] = (uint32_t)mv_pack(cx, cy); rec[1] = (uint32_t)mv_pack(bx, by); rec[2] = 1; rec[3] = 0;
    rec[SM_ME_OFF + ME_KEY + 15] = 0;       /* no motion-estimation record for this frame yet */
}
#endif
