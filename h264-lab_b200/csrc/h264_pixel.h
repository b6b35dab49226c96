/*
 * h264_pixel.h -- leaf pixel / coefficient operations of the macroblock path
 * (SURVEY.md 8(a) rows a1-a3, a7-a11, a16), written from the algorithm, not from
 * the reference's code structure: every output sample is computed independently
 * from its input neighbourhood so that a warp can stride over samples without
 * intermediate buffers or barriers.
 */
#pragma once
#include "h264_common.h"

/* ------------------------------------------------------------------------------
 * unaligned 4-byte load through a generic pointer: the reference samples come either
 * from the shared-memory search window of the macroblock (MBWork.win, the fast case) or
 * directly from the reference frame in global memory (any position outside the window),
 * so that both cases run the same code.
 * ---------------------------------------------------------------------------- */
HD uint32_t ld4u(const pix_t *p)
{
#if H264_DEVICE
    uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    unsigned sh = (unsigned)(a & 3) * 8;
    return __funnelshift_r(q[0], q[1], sh);
#else
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
#endif
}
HD int ldpx(const pix_t *p) { return (int)*p; }
HD uint32_t ld4_sm(const pix_t *p) { return *(const uint32_t *)p; }   /* 4-aligned shared/local */

HD int sad4(uint32_t a, uint32_t b)      /* sum of |a_i - b_i| over four packed bytes */
{
#if H264_DEVICE
    return (int)__vsadu4(a, b);
#else
    int s = 0;
    for (int k = 0; k < 4; k++) { int d = (int)((a >> (8 * k)) & 255) - (int)((b >> (8 * k)) & 255); s += d < 0 ? -d : d; }
    return s;
#endif
}
HD uint32_t avg4(uint32_t a, uint32_t b)  /* per-byte (a + b + 1) >> 1 */
{
#if H264_DEVICE
    return __vavgu4(a, b);
#else
    uint32_t o = 0;
    for (int k = 0; k < 4; k++) o |= ((((a >> (8 * k)) & 255) + ((b >> (8 * k)) & 255) + 1) >> 1) << (8 * k);
    return o;
#endif
}
HD void unpack4(uint32_t v, int *o) { o[0] = v & 255; o[1] = (v >> 8) & 255; o[2] = (v >> 16) & 255; o[3] = v >> 24; }
HD uint32_t pack4(int a, int b, int c, int d) { return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24); }

/* ------------------------------------------------------------------------------
 * a1: SAD of a w x h block (window or frame, any alignment) against the cached input
 * MB (stride 16).  (sad_block H:2162, h264e_sad_mb_unlaign_wh H:2189). w, h in {8,16}.
 * Returns the warp-uniform total.
 * ---------------------------------------------------------------------------- */
HDF_sad_frame_wh int sad_frame_wh(const pix_t *a, int a_stride, const pix_t *b16, int w, int h)
{
    const int sh = w == 16 ? 2 : 1;      /* words per row: 4 or 2 */
    int acc = 0;
    FOR_LANES(i, h << sh)
    {
        int r = i >> sh, c = (i & ((1 << sh) - 1)) * 4;
        acc += sad4(ld4u(a + r * a_stride + c), ld4_sm(b16 + r * 16 + c));
    }
    return wsum(acc);
}

/* SAD of two stride-16 blocks in the working set */
HDF_sad_sm_wh int sad_sm_wh(const pix_t *a16, const pix_t *b16, int w, int h)
{
    const int sh = w == 16 ? 2 : 1;
    int acc = 0;
    FOR_LANES(i, h << sh)
    {
        int r = i >> sh, c = (i & ((1 << sh) - 1)) * 4;
        acc += sad4(ld4_sm(a16 + r * 16 + c), ld4_sm(b16 + r * 16 + c));
    }
    return wsum(acc);
}

/* a1: four 8x8 quadrant SADs (TL, TR, BL, BR) of a 16x16 block + their sum
 * (h264e_sad_mb_unlaign_8x8 H:2178). */
HDF_sad_mb_quad int sad_mb_quad(const pix_t *a, int a_stride, const pix_t *b16, int sad4out[4])
{
    int q01 = 0, q23 = 0;    /* two 16-bit lanes each: an 8x8 SAD is <= 16320 */
    FOR_LANES(i, 64)
    {
        int r = i >> 2, c = (i & 3) * 4;
        int s = sad4(ld4u(a + r * a_stride + c), ld4_sm(b16 + r * 16 + c));
        s <<= (c & 8) ? 16 : 0;
        if (r < 8) q01 += s; else q23 += s;
    }
    q01 = wsum(q01);
    q23 = wsum(q23);
    sad4out[0] = q01 & 0xFFFF; sad4out[1] = (int)((uint32_t)q01 >> 16);
    sad4out[2] = q23 & 0xFFFF; sad4out[3] = (int)((uint32_t)q23 >> 16);
    return sad4out[0] + sad4out[1] + sad4out[2] + sad4out[3];
}

/* SADs of the eight integer neighbours of block position p in ONE pass (speculative
 * evaluation for the greedy diamond of me_search: the search consumes them in the
 * reference's order, so evaluating more positions than the reference changes nothing).
 * out[0..3] = (+1,0) (-1,0) (0,+1) (0,-1); out[4..7] = (+1,+1) (-1,+1) (+1,-1) (-1,-1).
 * A 16x16 SAD is <= 65280, so two fit one 32-bit accumulator. */
HDF_sad_nb8 void sad_nb8(const pix_t *p, int ps, const pix_t *b16, int w, int h, int out[8])
{
    const int sh = w == 16 ? 2 : 1;
    uint32_t a01 = 0, a23 = 0, a45 = 0, a67 = 0;
    FOR_LANES(i, h << sh)
    {
        int r = i >> sh, c = (i & ((1 << sh) - 1)) * 4;
        const pix_t *q = p + r * ps + c;
        uint32_t in = ld4_sm(b16 + r * 16 + c);
        uint32_t up_l = ld4u(q - ps - 1), up_c = ld4u(q - ps), up_r = ld4u(q - ps + 1);
        uint32_t mi_l = ld4u(q - 1), mi_r = ld4u(q + 1);
        uint32_t dn_l = ld4u(q + ps - 1), dn_c = ld4u(q + ps), dn_r = ld4u(q + ps + 1);
        a01 += (uint32_t)sad4(mi_r, in) | ((uint32_t)sad4(mi_l, in) << 16);
        a23 += (uint32_t)sad4(dn_c, in) | ((uint32_t)sad4(up_c, in) << 16);
        a45 += (uint32_t)sad4(dn_r, in) | ((uint32_t)sad4(dn_l, in) << 16);
        a67 += (uint32_t)sad4(up_r, in) | ((uint32_t)sad4(up_l, in) << 16);
    }
    a01 = (uint32_t)wsum((int)a01); a23 = (uint32_t)wsum((int)a23);
    a45 = (uint32_t)wsum((int)a45); a67 = (uint32_t)wsum((int)a67);
    out[0] = a01 & 0xFFFF; out[1] = a01 >> 16; out[2] = a23 & 0xFFFF; out[3] = a23 >> 16;
    out[4] = a45 & 0xFFFF; out[5] = a45 >> 16; out[6] = a67 & 0xFFFF; out[7] = a67 >> 16;
}

/* SADs of the seven sub-pel probes of me_search in one pass, from the integer prediction I
 * and the three half-sample blocks H1 (primary), H2 (secondary), C (diagonal):
 * out = { H1, avg(I,H1), H2, avg(I,H2), avg(H1,H2), C, avg(C,H1) }  (H:5119-5161) */
HDF_sad_qpel7 void sad_qpel7(const pix_t *I, const pix_t *H1, const pix_t *H2, const pix_t *C, const pix_t *b16, int w, int h, int out[7])
{
    const int sh = w == 16 ? 2 : 1;
    uint32_t a01 = 0, a23 = 0, a45 = 0, a6 = 0;
    FOR_LANES(i, h << sh)
    {
        int r = i >> sh, o = r * 16 + (i & ((1 << sh) - 1)) * 4;
        uint32_t in = ld4_sm(b16 + o), vi = ld4_sm(I + o), v1 = ld4_sm(H1 + o), v2 = ld4_sm(H2 + o), vc = ld4_sm(C + o);
        a01 += (uint32_t)sad4(v1, in) | ((uint32_t)sad4(avg4(vi, v1), in) << 16);
        a23 += (uint32_t)sad4(v2, in) | ((uint32_t)sad4(avg4(vi, v2), in) << 16);
        a45 += (uint32_t)sad4(avg4(v1, v2), in) | ((uint32_t)sad4(vc, in) << 16);
        a6 += (uint32_t)sad4(avg4(vc, v1), in);
    }
    a01 = (uint32_t)wsum((int)a01); a23 = (uint32_t)wsum((int)a23);
    a45 = (uint32_t)wsum((int)a45); a6 = (uint32_t)wsum((int)a6);
    out[0] = a01 & 0xFFFF; out[1] = a01 >> 16; out[2] = a23 & 0xFFFF; out[3] = a23 >> 16;
    out[4] = a45 & 0xFFFF; out[5] = a45 >> 16; out[6] = a6;
}

/* ------------------------------------------------------------------------------
 * a2: luma sub-pel interpolation.  Six-tap (1,-5,20,20,-5,1) half-pel, centre half-pel
 * from 16-bit horizontal intermediates, quarter positions as rounded averages of the two
 * nearest integer / half samples (H:1971-2130, ITU-T H.264 8.4.2.2.1).  Block functions:
 * each lane produces 8 output samples from packed words, dst stride 16.
 * ---------------------------------------------------------------------------- */
HD int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }

HDF_copy_block void copy_block(const pix_t *src, int ss, pix_t *dst, int w, int h)
{
    const int sh = w == 16 ? 2 : 1;
    FOR_LANES(i, h << sh)
    {
        int r = i >> sh, c = (i & ((1 << sh) - 1)) * 4;
        *(uint32_t *)(dst + r * 16 + c) = ld4u(src + r * ss + c);
    }
}

/* ------------------------------------------------------------------------------
 * Half-sample planes.  The reference filters every probed block on the fly (hpel_lpf_hor /
 * _ver / _diag, H:1971-2051); the filters depend on nothing but the reference picture, so
 * here they are applied ONCE per finished picture, to the whole padded luma plane, by a
 * trivially parallel pass (k_hpel): planes b (horizontal), h (vertical), j (centre, from the
 * 16-bit horizontal intermediates), stored with the luma plane's own linear indexing so that a
 * probe is a copy.  One call produces the four samples of word `wi` of each plane; taps
 * address the padded buffer linearly, exactly like a filter run at that position would
 * (indices clamped to the buffer, which no position the search may use ever needs).
 * ---------------------------------------------------------------------------- */
HD uint32_t hp_ldw(const pix_t *buf, long nwords, long wi)
{
    wi = wi < 0 ? 0 : (wi >= nwords ? nwords - 1 : wi);
    return *(const uint32_t *)(buf + 4 * wi);
}
HD void hpel_word(const pix_t *buf, long nwords, int stride, long wi, uint32_t *ph, uint32_t *pv, uint32_t *pd)
{
    const int sw = stride >> 2;
    int t[6][4], c[6][4];
#pragma unroll
    for (int k = 0; k < 6; k++)
    {
        const long wr = wi + (long)(k - 2) * sw;
        int b[12];
        unpack4(hp_ldw(buf, nwords, wr - 1), b); unpack4(hp_ldw(buf, nwords, wr), b + 4); unpack4(hp_ldw(buf, nwords, wr + 1), b + 8);
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            t[k][j] = (int16_t)tap6(b[j + 2], b[j + 3], b[j + 4], b[j + 5], b[j + 6], b[j + 7]);
            c[k][j] = b[j + 4];
        }
    }
    int oh[4], ov[4], od[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        oh[j] = clip_u8((t[2][j] + 16) >> 5);
        ov[j] = clip_u8((tap6(c[0][j], c[1][j], c[2][j], c[3][j], c[4][j], c[5][j]) + 16) >> 5);
        od[j] = clip_u8((tap6(t[0][j], t[1][j], t[2][j], t[3][j], t[4][j], t[5][j]) + 512) >> 10);
    }
    *ph = pack4(oh[0], oh[1], oh[2], oh[3]);
    *pv = pack4(ov[0], ov[1], ov[2], ov[3]);
    *pd = pack4(od[0], od[1], od[2], od[3]);
}

/* rounded average of two stride-16 blocks (h264e_qpel_average_wh_align H:2065);
 * dst may alias either source (element-wise) */
HDF_average_block void average_block(const pix_t *s0, const pix_t *s1, pix_t *dst, int w, int h)
{
    const int sh = w == 16 ? 2 : 1;
    FOR_LANES(i, h << sh)
    {
        int r = i >> sh, c = (i & ((1 << sh) - 1)) * 4;
        *(uint32_t *)(dst + r * 16 + c) = avg4(ld4_sm(s0 + r * 16 + c), ld4_sm(s1 + r * 16 + c));
    }
}

/* Prediction block at quarter-sample offset (dx,dy) from integer position `g` (a pointer into
 * the integer plane, any alignment) -> dst (stride 16): the position table of
 * h264e_qpel_interpolate_luma (H:2079-2130) with the filters replaced by the half-sample
 * planes.  hb/hv/hd address the same sample as g in planes b/h/j; all four share `stride`. */
HDF_interp_luma_planes void interp_luma_planes(const pix_t *g, const pix_t *hb, const pix_t *hv, const pix_t *hd, int stride,
                            int dx, int dy, int w, int h, pix_t *dst)
{
    const int pos = 1 << (dx + 4 * dy);
    const pix_t *a = g, *b = 0;
    if (pos != 1)
    {
        a = 0;
        if (pos & 0xe0ee) a = hb + ((pos & 0xe000) ? stride : 0);
        if (pos & 0xbbb0) { const pix_t *q = hv + ((pos & 0x8880) ? 1 : 0); if (a) b = q; else a = q; }
        if (pos & 0x4e40) { if (a) b = hd; else a = hd; }
        if ((pos & 0xfafa) && !b) b = g + ((dx + 1) >> 2) + ((dy + 1) >> 2) * stride;
    }
    const int sh = w == 16 ? 2 : 1;
    FOR_LANES(i, h << sh)
    {
        int r = i >> sh, c = (i & ((1 << sh) - 1)) * 4;
        uint32_t v = ld4u(a + r * stride + c);
        if (b) v = avg4(v, ld4u(b + r * stride + c));
        *(uint32_t *)(dst + r * 16 + c) = v;
    }
}

/* a3: chroma 1/8-pel bilinear block (h264e_qpel_interpolate_chroma H:2133).
 * src addresses the integer sample of the block's top-left, (dx,dy) in 0..7. */
HDF_interp_chroma_block void interp_chroma_block(const pix_t *src, int stride, int dx, int dy, int w, int h, pix_t *dst)
{
    int a = (8 - dx) * (8 - dy), b = dx * (8 - dy), c = (8 - dx) * dy, d = dx * dy;
    FOR_LANES(i, w * h)
    {
        int r = i / w, x = i - r * w;
        const pix_t *p = src + r * stride + x;
        int v;
        if (dx | dy) v = (a * ldpx(p) + b * ldpx(p + 1) + c * ldpx(p + stride) + d * ldpx(p + stride + 1) + 32) >> 6;
        else v = ldpx(p);
        dst[r * 16 + x] = (pix_t)v;
    }
}

/* ------------------------------------------------------------------------------
 * a9: forward 4x4 core transform of (inp - pred); result stored transposed,
 * out[v + 4u] = coefficient with vertical frequency v and horizontal frequency u
 * (FwdTransformResidual4x42 H:2385, TRANSPOSE_BLOCK 1).
 * ---------------------------------------------------------------------------- */
HDF_fwd4x4 void fwd4x4(const pix_t *inp, int inp_stride, const pix_t *pred, int16_t *out)
{
    int t[16];
#pragma unroll
    for (int x = 0; x < 4; x++)          /* vertical pass on column x */
    {
        int f0 = (int)inp[x] - pred[x];
        int f1 = (int)inp[x + inp_stride] - pred[x + 16];
        int f2 = (int)inp[x + 2 * inp_stride] - pred[x + 32];
        int f3 = (int)inp[x + 3 * inp_stride] - pred[x + 48];
        int s03 = f0 + f3, d03 = f0 - f3, s12 = f1 + f2, d12 = f1 - f2;
        t[x * 4 + 0] = s03 + s12;
        t[x * 4 + 1] = 2 * d03 + d12;
        t[x * 4 + 2] = s03 - s12;
        t[x * 4 + 3] = d03 - 2 * d12;
    }
#pragma unroll
    for (int v = 0; v < 4; v++)          /* horizontal pass on vertical frequency v */
    {
        int d0 = t[v], d1 = t[v + 4], d2 = t[v + 8], d3 = t[v + 12];
        int s03 = d0 + d3, d03 = d0 - d3, s12 = d1 + d2, d12 = d1 - d2;
        out[v + 0] = (int16_t)(s03 + s12);
        out[v + 4] = (int16_t)(2 * d03 + d12);
        out[v + 8] = (int16_t)(s03 - s12);
        out[v + 12] = (int16_t)(d03 - 2 * d12);
    }
}

/* a11: inverse core transform of dq[] (same transposed layout) added to pred and
 * clipped (TransformResidual4x4 H:2436 + h264e_transform_add H:2638). */
HDF_inv4x4_add void inv4x4_add(const int16_t *dq, const pix_t *pred, pix_t *out, int out_stride)
{
    int t[16];
#pragma unroll
    for (int v = 0; v < 4; v++)          /* horizontal inverse for vertical frequency v */
    {
        int d0 = dq[v], d1 = dq[v + 4], d2 = dq[v + 8], d3 = dq[v + 12];
        int e0 = d0 + d2, e1 = d0 - d2, e2 = (d1 >> 1) - d3, e3 = d1 + (d3 >> 1);
        t[v * 4 + 0] = (int16_t)(e0 + e3);
        t[v * 4 + 1] = (int16_t)(e1 + e2);
        t[v * 4 + 2] = (int16_t)(e1 - e2);
        t[v * 4 + 3] = (int16_t)(e0 - e3);
    }
#pragma unroll
    for (int x = 0; x < 4; x++)          /* vertical inverse on column x */
    {
        int f0 = t[x], f1 = t[x + 4], f2 = t[x + 8], f3 = t[x + 12];
        int g0 = f0 + f2, g1 = f0 - f2, g2 = (f1 >> 1) - f3, g3 = f1 + (f3 >> 1);
        int r0 = (int16_t)((g0 + g3 + 32) >> 6), r1 = (int16_t)((g1 + g2 + 32) >> 6);
        int r2 = (int16_t)((g1 - g2 + 32) >> 6), r3 = (int16_t)((g0 - g3 + 32) >> 6);
        out[x] = (pix_t)clip_u8(r0 + pred[x]);
        out[x + out_stride] = (pix_t)clip_u8(r1 + pred[x + 16]);
        out[x + 2 * out_stride] = (pix_t)clip_u8(r2 + pred[x + 32]);
        out[x + 3 * out_stride] = (pix_t)clip_u8(r3 + pred[x + 48]);
    }
}

HD void copy4x4(const pix_t *pred, pix_t *out, int out_stride)
{
#pragma unroll
    for (int r = 0; r < 4; r++) *(uint32_t *)(out + r * out_stride) = *(const uint32_t *)(pred + r * 16);
}

/* position class of coefficient i inside qdat: 0 -> (0,0)-type, 2 -> mixed, 4 -> (1,1)-type
 * (g_idx2quant H:2366) */
HD int quant_class(int i) { return ((i & 1) + ((i >> 2) & 1)) * 2; }

/* "all coefficients from i0 on are small" test against 8 thresholds (is_zero H:2491) */
HDF_coefs_small int coefs_small(const int16_t *c, int i0, const uint16_t *thr)
{
    for (int i = i0; i < 16; i++)
    {
        unsigned t = thr[i & 7];
        if ((unsigned)((int)c[i] + (int)t) > 2u * t) return 0;
    }
    return 1;
}

/* a9: dead-zone quantiser + dequantiser of one 4x4 block, coefficients i0..15
 * (inner loop of quantize(), H:2567-2585).  Returns 1 when any level is non-zero. */
HDF_quant4x4 int quant4x4(int16_t *dq, int16_t *qv, int i0, const uint16_t *qdat)
{
    int nz = 0;
    int rnd = qdat[6];
    for (int i = i0; i < 16; i++)
    {
        int cl = quant_class(i);
        int c = dq[i];
        int r = c < 0 ? 0xFFFF - rnd : rnd;
        int v = (c * (int)qdat[cl] + r) >> 16;
        nz |= v;
        qv[i] = (int16_t)v;
        dq[i] = (int16_t)(v * (int)qdat[cl + 1]);
    }
    return nz != 0;
}

/* 4-point butterfly used by both passes of the 4x4 Hadamard (hadamar4_2d H:2269) */
HD void had4(int a, int b, int c, int d, int *o)
{
    int s = a + c, t = a - c, u = b + d, w = b - d;
    o[0] = s + u; o[1] = t + w; o[2] = t - w; o[3] = s - u;
}

/* a10: luma DC path of an Intra16x16 MB (h264e_quant_luma_dc H:2344):
 * dc[16] = DC transform coefficients of the 16 blocks (raster);
 * out: qdc[16] quantised levels, dq0[16] dequantised DC per block.  Serial (one lane). */
HDN void luma_dc_quant(const int16_t *dc, int16_t *qdc, int16_t *dq0, const uint16_t *qdat)
{
    int t[16], x[16], o[4];
    for (int i = 0; i < 4; i++) { had4(dc[i], dc[i + 4], dc[i + 8], dc[i + 12], o); for (int k = 0; k < 4; k++) t[4 * i + k] = (int16_t)o[k]; }
    for (int k = 0; k < 4; k++) { had4(t[k], t[k + 4], t[k + 8], t[k + 12], o); for (int m = 0; m < 4; m++) x[k + 4 * m] = (int16_t)o[m]; }
    int q = (int16_t)qdat[0];
    for (int i = 0; i < 16; i++)
    {
        int v = x[i];
        int r = v < 0 ? (1 << 18) - 0x20000 : 0x20000;
        v = (v * q + r) >> 18;
        qdc[i] = (int16_t)v;
        x[i] = (int16_t)v;
    }
    for (int i = 0; i < 4; i++) { had4(x[i], x[i + 4], x[i + 8], x[i + 12], o); for (int k = 0; k < 4; k++) t[4 * i + k] = (int16_t)o[k]; }
    for (int k = 0; k < 4; k++) { had4(t[k], t[k + 4], t[k + 8], t[k + 12], o); for (int m = 0; m < 4; m++) x[k + 4 * m] = (int16_t)o[m]; }
    int d = (int16_t)(qdat[1] >> 2);
    for (int i = 0; i < 16; i++) dq0[i] = (int16_t)(x[i] * d);
}

/* a10: chroma DC path of one plane (h264e_quant_chroma_dc H:2355). Returns 1 when any
 * quantised DC level is non-zero. */
HDN int chroma_dc_quant(const int16_t *dc, int16_t *qdc, int16_t *dq0, const uint16_t *qdat)
{
    int a = dc[0], b = dc[1], c = dc[2], d = dc[3], x[4], y[4];
    x[0] = (int16_t)(a + b + c + d); x[1] = (int16_t)(a - b + c - d);
    x[2] = (int16_t)(a + b - c - d); x[3] = (int16_t)(a - b - c + d);
    int q = (int16_t)(qdat[0] << 1);
    for (int i = 0; i < 4; i++)
    {
        int v = x[i];
        int r = v < 0 ? (1 << 18) - 0xAAAA : 0xAAAA;
        v = (v * q + r) >> 18;
        qdc[i] = (int16_t)v;
        x[i] = (int16_t)v;
    }
    y[0] = (int16_t)(x[0] + x[1] + x[2] + x[3]); y[1] = (int16_t)(x[0] - x[1] + x[2] - x[3]);
    y[2] = (int16_t)(x[0] + x[1] - x[2] - x[3]); y[3] = (int16_t)(x[0] - x[1] - x[2] + x[3]);
    int dqm = (int16_t)(qdat[1] >> 1);
    for (int i = 0; i < 4; i++) dq0[i] = (int16_t)(y[i] * dqm);
    return (y[0] | y[1] | y[2] | y[3]) != 0;
}

/* ------------------------------------------------------------------------------
 * a7: DC value from the available neighbours (intra_predict_dc H:1625)
 * ---------------------------------------------------------------------------- */
HD int dc_pred(const pix_t *left, const pix_t *top, int n, int log2n)
{
    int s = 0, k = 0;
    if (left) { for (int i = 0; i < n; i++) s += left[i]; k++; }
    if (top) { for (int i = 0; i < n; i++) s += top[i]; k++; }
    if (!k) return 128;
    if (k == 2) return (s + n) >> (log2n + 1);
    return (s + (n >> 1)) >> log2n;
}

/* a7: 16x16 luma prediction, mode 0=V 1=H 2=DC (h264e_intra_predict_16x16 H:1677).
 * left/top are NULL when unavailable. */
HDF_intra16_pred void intra16_pred(pix_t *dst, const pix_t *left, const pix_t *top, int mode)
{
    int dc = 0;
    if (mode == 2) dc = dc_pred(left, top, 16, 4) * 0x01010101u;
    FOR_LANES(i, 64)
    {
        int r = i >> 2, c = (i & 3) * 4;
        uint32_t v;
        if (mode == 0) v = ld4_sm(top + c);
        else if (mode == 1) v = left[r] * 0x01010101u;
        else v = (uint32_t)dc;
        *(uint32_t *)(dst + r * 16 + c) = v;
    }
}

/* a7: 8x8 chroma prediction for both planes (h264e_intra_predict_chroma H:1716).
 * dst: U at +0, V at +8, stride 16.  left/top: U 0..7, V 8..15, NULL when unavailable.
 * mode uses the LUMA numbering 0=V 1=H 2=DC, as the reference calls it (H:5784). */
HDN void intra_chroma_pred(pix_t *dst, const pix_t *left, const pix_t *top, int mode)
{
    FOR_LANES(i, 32)
    {
        int r = i >> 2, q = i & 3;          /* q: 0,1 = U left/right 4 samples; 2,3 = V */
        int pl = q >> 1, xh = q & 1, yh = r >> 2;
        uint32_t v;
        if (mode == 0) v = ld4_sm(top + pl * 8 + xh * 4);
        else if (mode == 1) v = left[pl * 8 + r] * 0x01010101u;
        else
        {
            const pix_t *l = left ? left + pl * 8 + yh * 4 : 0;
            const pix_t *t = top ? top + pl * 8 + xh * 4 : 0;
            int dc;
            if (xh == yh) dc = dc_pred(l, t, 4, 2);            /* corner blocks: both edges   */
            else if (xh) dc = t ? dc_pred(0, t, 4, 2) : dc_pred(l, 0, 4, 2);   /* top-right: top preferred  */
            else dc = l ? dc_pred(l, 0, 4, 2) : dc_pred(0, t, 4, 2);           /* bottom-left: left preferred */
            v = (uint32_t)dc * 0x01010101u;
        }
        *(uint32_t *)(dst + r * 16 + q * 4) = v;
    }
}

/* a7: gradient heuristic that picks the Intra16x16 mode from eight input samples
 * (intra_estimate_16x16 H:4838). p = input MB, stride 16. */
HD int intra16_estimate(const pix_t *p, int avail, int qp)
{
    int p00 = p[0], p01 = p[15], p10 = p[15 * 16], p11 = p[15 * 16 + 15];
    int dx = iabs(p00 - p01) + iabs(p10 - p11) + iabs((int)p[8 * 16] - p[8 * 16 + 15]);
    int dy = iabs(p00 - p10) + iabs(p01 - p11) + iabs((int)p[8] - p[15 * 16 + 8]);
    if (dx > 30 + 3 * dy && dy < 150 - qp && (avail & AVAIL_T)) return 0;
    if (dy > 30 + 3 * dx && dx < 150 - qp && (avail & AVAIL_L)) return 1;
    return 2;
}

/* ------------------------------------------------------------------------------
 * a8: one of the nine Intra4x4 predictions (ITU-T H.264 8.3.1.2, as implemented by
 * h264e_intra_choose_4x4 H:1810-1960).  e[] holds the 13 neighbours:
 *   e[0..3] = L3,L2,L1,L0   e[4] = UL   e[5..12] = U0..U7
 * i.e. e[4+k] is the sample k steps along the edge from the corner (k<0: left).
 * ---------------------------------------------------------------------------- */
HD void intra4_predict(int mode, const int *e, int avail, pix_t *out /* 16, stride 4 */)
{
#define PT(j) e[5 + (j)]      /* p[j,-1], j = -1..7 */
#define PL(j) e[3 - (j)]      /* p[-1,j], j = -1..3 */
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++)
        {
            int v;
            switch (mode)
            {
            case 0: v = PT(x); break;
            case 1: v = PL(y); break;
            case 2:
            {
                int s = 0, k = 0;
                if (avail & AVAIL_L) { s += PL(0) + PL(1) + PL(2) + PL(3); k++; }
                if (avail & AVAIL_T) { s += PT(0) + PT(1) + PT(2) + PT(3); k++; }
                v = k == 0 ? 128 : (k == 2 ? (s + 4) >> 3 : (s + 2) >> 2);
                break;
            }
            case 3:     /* diagonal down-left */
                v = (x == 3 && y == 3) ? (PT(6) + 3 * PT(7) + 2) >> 2 : (PT(x + y) + 2 * PT(x + y + 1) + PT(x + y + 2) + 2) >> 2;
                break;
            case 4:     /* diagonal down-right */
                if (x > y) v = (PT(x - y - 2) + 2 * PT(x - y - 1) + PT(x - y) + 2) >> 2;
                else if (x < y) v = (PL(y - x - 2) + 2 * PL(y - x - 1) + PL(y - x) + 2) >> 2;
                else v = (PT(0) + 2 * PT(-1) + PL(0) + 2) >> 2;
                break;
            case 5:     /* vertical-right */
            {
                int z = 2 * x - y, k = x - (y >> 1);
                if (z >= 0 && !(z & 1)) v = (PT(k - 1) + PT(k) + 1) >> 1;
                else if (z >= 0) v = (PT(k - 2) + 2 * PT(k - 1) + PT(k) + 2) >> 2;
                else if (z == -1) v = (PL(0) + 2 * PT(-1) + PT(0) + 2) >> 2;
                else v = (PL(y - 1) + 2 * PL(y - 2) + PL(y - 3) + 2) >> 2;
                break;
            }
            case 6:     /* horizontal-down */
            {
                int z = 2 * y - x, k = y - (x >> 1);
                if (z >= 0 && !(z & 1)) v = (PL(k - 1) + PL(k) + 1) >> 1;
                else if (z >= 0) v = (PL(k - 2) + 2 * PL(k - 1) + PL(k) + 2) >> 2;
                else if (z == -1) v = (PL(0) + 2 * PT(-1) + PT(0) + 2) >> 2;
                else v = (PT(x - 1) + 2 * PT(x - 2) + PT(x - 3) + 2) >> 2;
                break;
            }
            case 7:     /* vertical-left */
                if (!(y & 1)) v = (PT(x + (y >> 1)) + PT(x + (y >> 1) + 1) + 1) >> 1;
                else v = (PT(x + (y >> 1)) + 2 * PT(x + (y >> 1) + 1) + PT(x + (y >> 1) + 2) + 2) >> 2;
                break;
            default:    /* horizontal-up */
            {
                int z = x + 2 * y, k = y + (x >> 1);
                if (z > 5) v = PL(3);
                else if (z == 5) v = (PL(2) + 3 * PL(3) + 2) >> 2;
                else if (!(z & 1)) v = (PL(k) + PL(k + 1) + 1) >> 1;
                else v = (PL(k) + 2 * PL(k + 1) + PL(k + 2) + 2) >> 2;
                break;
            }
            }
            out[y * 4 + x] = (pix_t)v;
        }
#undef PT
#undef PL
}
