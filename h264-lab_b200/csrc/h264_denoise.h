/*
 * h264_denoise.h -- temporal noise suppressor applied to the input picture before it is encoded
 * (h264e_denoise_run H:1547-1620, hooked into H264E_encode at H:6686-6696).
 *
 * The reference filters in place with a row-shuffling trick (every output row is parked one row up in
 * the previous-picture buffer and moved back at the end, H:1600-1619); read in raster order that is a
 * pure function of two pictures: out = cur on the picture border, and inside
 *     d  = |cur - prev|,  nb = |sum over the 4-neighbours of (cur - prev)| >> 2
 *     g  = (255 - gain[d]) * (255 - min(4 gain[nb], 255))                       (Q16 weight of prev)
 *     out = (prev * g + (0xffff - g) * cur + 0x8000) >> 16
 * with prev = the previous OUTPUT of the filter (all zero before the first picture, H:6345-6349).
 * Here the filter writes a second buffer instead (ping-pong), one thread per 4 samples: HBM-bound,
 * 3 bytes of traffic per sample.  The filtered picture is what the macroblock path then reads as input.
 */
#pragma once
#include "h264_common.h"
#include "h264_denoise_tab.h"

/* one output sample; c* = current picture (centre, left, right, up, down), p* = previous output */
HD int denoise_sample(int cc, int cl, int cr, int cu, int cd, int pc, int pl, int pr, int pu, int pd)
{
    int d = cc - pc;
    int nb = (cl - pl) + (cr - pr) + (cu - pu) + (cd - pd);
    if (d < 0) d = -d;
    if (nb < 0) nb = -nb;
    nb >>= 2;
    const unsigned g = (unsigned)denoise_weight[d][0] * (unsigned)denoise_weight[nb][1];
    return (int)(((unsigned)pc * g + (0xffffu - g) * (unsigned)cc + (1u << 15)) >> 16);
}

/* samples x0 .. x0+3 of row y of a w x h plane -> out; planes smaller than 3x3 are left alone (H:1550) */
HD void denoise_word(const pix_t *cur, int cur_stride, const pix_t *prev, pix_t *out, int dn_stride, int w, int h, int x0, int y)
{
    for (int k = 0; k < 4; k++)
    {
        const int x = x0 + k;
        if (x >= w) break;
        const pix_t *c = cur + y * cur_stride + x;
        const pix_t *p = prev + y * dn_stride + x;
        int v;
        if (x == 0 || y == 0 || x == w - 1 || y == h - 1) v = c[0];
        else v = denoise_sample(c[0], c[-1], c[1], c[-cur_stride], c[cur_stride], p[0], p[-1], p[1], p[-dn_stride], p[dn_stride]);
        out[y * dn_stride + x] = (pix_t)v;
    }
}
