/*
 * h264_cavlc.h -- macroblock-layer syntax + CAVLC residual coding for one
 * macroblock, executed by ONE thread per macroblock (every macroblock of the frame
 * in parallel), followed by a prefix sum of the bit lengths and a scatter-pack
 * into the slice payload (SURVEY.md 8(a) rows a12 syntax part, a13).
 *
 * Bit strings are kept as 32-bit words, most significant bit first.
 *
 * The reference codes coefficients in plain array order without zig-zag
 * (h264e_vlc_encode H:2786-2798); that order is reproduced.
 */
#pragma once
#include "h264_common.h"
#include "h264_cavlc_tables.h"

struct BitW
{
    uint32_t *buf;
    uint64_t acc;
    int nacc;     /* valid bits in acc (< 32 between calls) */
    int pos;      /* complete words written */
    int cap;
};

HD void bw_init(BitW &b, uint32_t *buf, int cap) { b.buf = buf; b.acc = 0; b.nacc = 0; b.pos = 0; b.cap = cap; }
HD void bw_put(BitW &b, int n, uint32_t val)
{
    b.acc = (b.acc << n) | val;
    b.nacc += n;
    if (b.nacc >= 32)
    {
        b.nacc -= 32;
        if (b.pos < b.cap) b.buf[b.pos] = (uint32_t)(b.acc >> b.nacc);
        b.pos++;
    }
}
HD int bw_bits(const BitW &b) { return b.pos * 32 + b.nacc; }
HD void bw_flush(BitW &b)       /* left-align the partial word */
{
    if (b.nacc && b.pos < b.cap) b.buf[b.pos] = (uint32_t)(b.acc << (32 - b.nacc));
}
HD void bw_ue(BitW &b, uint32_t v)         /* h264e_bs_put_golomb H:2738 */
{
    uint32_t t = v + 1;
    int size = 0;
    for (uint32_t u = t; u; u >>= 1) size++;
    bw_put(b, 2 * size - 1, t);
}
HD void bw_se(BitW &b, int v)              /* h264e_bs_put_sgolomb H:2760 */
{
    v = 2 * v - 1;
    v ^= v >> 31;
    bw_ue(b, (uint32_t)v);
}

/* coded_block_pattern -> codeNum (ITU-T H.264 Table 9-4, inverted).  Row 0: intra 4x4,
 * row 1: inter.  The forward table (codeNum -> cbp) is the standard's. */
H264_TAB uint8_t cbp_intra_of_code[48] = {
    47, 31, 15, 0, 23, 27, 29, 30, 7, 11, 13, 14, 39, 43, 45, 46, 16, 3, 5, 10, 12, 19, 21, 26,
    28, 35, 37, 42, 44, 1, 2, 4, 8, 17, 18, 20, 24, 6, 9, 22, 25, 32, 33, 34, 36, 40, 38, 41};
H264_TAB uint8_t cbp_inter_of_code[48] = {
    0, 16, 1, 2, 4, 8, 32, 3, 5, 10, 12, 15, 47, 7, 11, 13, 14, 6, 9, 31, 35, 37, 42, 44,
    33, 34, 36, 40, 39, 43, 45, 46, 17, 18, 20, 24, 19, 21, 26, 28, 23, 27, 29, 30, 22, 25, 38, 41};
HD int cbp_code(int inter, int cbp)
{
    const uint8_t *t = inter ? cbp_inter_of_code : cbp_intra_of_code;
    for (int i = 0; i < 48; i++) if (t[i] == cbp) return i;
    return 0;
}

/* a13: one residual block.  c = first coded coefficient, n = maxNumCoeff (4, 15, 16),
 * nA/nB = neighbouring total_coeff (NNZ_NA when unavailable, 17/17 selects the
 * chroma DC table).  (h264e_vlc_encode H:2775-2949) */
HD void cavlc_block(BitW &b, const int16_t *c, int n, int nA, int nB)
{
    int lv[16], idx[16];
    int nnz = 0;
    for (int i = n - 1; i >= 0; i--)
        if (c[i]) { lv[nnz] = c[i]; idx[nnz] = i; nnz++; }
    int t1 = 0;
    while (t1 < 3 && t1 < nnz && (lv[t1] == 1 || lv[t1] == -1)) t1++;

    int ctx = nA + nB;
    if (ctx <= 34) ctx = (ctx + 1) >> 1;
    ctx &= 31;
    int tab = ctx < 2 ? 0 : (ctx < 4 ? 1 : (ctx < 8 ? 2 : (ctx < 17 ? 3 : 4)));
    bw_put(b, cavlc_coeff_token_len[tab][nnz * 4 + t1], cavlc_coeff_token_code[tab][nnz * 4 + t1]);
    if (!nnz) return;

    if (t1)
    {
        uint32_t signs = 0;
        for (int k = 0; k < t1; k++) signs = (signs << 1) | (lv[k] < 0);
        bw_put(b, t1, signs);
    }
    int sl = (nnz > 10 && t1 < 3) ? 1 : 0;
    for (int k = t1; k < nnz; k++)
    {
        int level = lv[k];
        int mag = level < 0 ? -level : level;
        int code = level > 0 ? 2 * level - 2 : -2 * level - 1;
        if (k == t1 && t1 < 3) code -= 2;
        int prefix, sbits;
        uint32_t suffix;
        if (sl == 0)
        {
            if (code < 14) { prefix = code; sbits = 0; suffix = 0; }
            else if (code < 30) { prefix = 14; sbits = 4; suffix = (uint32_t)(code - 14); }
            else { prefix = 15; sbits = 12; suffix = (uint32_t)(code - 30); }
        } else
        {
            prefix = code >> sl;
            if (prefix < 15) { sbits = sl; suffix = (uint32_t)(code - (prefix << sl)); }
            else { prefix = 15; sbits = 12; suffix = (uint32_t)(code - (15 << sl)); }
        }
        bw_put(b, prefix + 1 + sbits, (1u << sbits) | suffix);
        if (sl == 0) sl = 1;
        if (mag > (3 << (sl - 1)) && sl < 6) sl++;
    }
    if (nnz < n)
    {
        int tz = idx[0] + 1 - nnz;
        if (n == 4) bw_put(b, cavlc_total_zeros_dc_len[nnz - 1][tz], cavlc_total_zeros_dc_code[nnz - 1][tz]);
        else bw_put(b, cavlc_total_zeros_len[nnz - 1][tz], cavlc_total_zeros_code[nnz - 1][tz]);
        int zl = tz;
        for (int k = 0; k < nnz - 1 && zl > 0; k++)
        {
            int run = idx[k] - idx[k + 1] - 1;
            int t = (zl < 7 ? zl : 7) - 1;
            bw_put(b, cavlc_run_before_len[t][run], cavlc_run_before_code[t][run]);
            zl -= run;
        }
    }
}

/* total_coeff of a neighbouring block for the coeff_token context */
HD int nnz_left(const FrameParams *fp, const MBInfo *mi, int mbx, int idx_in_left, int idx_in_cur, int at_edge)
{
    if (!at_edge) return mi->nnz[idx_in_cur];
    if (mbx == 0) return NNZ_NA;
    return mi[-1].nnz[idx_in_left];
}
HD int nnz_top(const FrameParams *fp, const MBInfo *mi, int mby, int idx_in_top, int idx_in_cur, int at_edge)
{
    if (!at_edge) return mi->nnz[idx_in_cur];
    if (mby == 0) return NNZ_NA;
    return mi[-fp->nmbx].nnz[idx_in_top];
}

/* a12 (syntax half of mb_write, H:4501-4690): macroblock n of the frame, or the
 * trailing skip run when n == nmb.  Writes the bits to slot n and returns the count. */
HD int cavlc_mb(const FrameParams *fp, int n)
{
    const int nmb = fp->nmbx * fp->nmby;
    BitW b;
    bw_init(b, fp->mb_bits + (size_t)n * MB_BITS_WORDS, MB_BITS_WORDS);
    const int is_p = fp->slice_type == SLICE_P;
    int run = 0;
    if (is_p) { for (int k = n - 1; k >= 0 && fp->mbi[k].type == MBT_SKIP; k--) run++; }
    if (n == nmb)
    {
        if (run) bw_ue(b, (uint32_t)run);
        bw_flush(b);
        return bw_bits(b);
    }
    const MBInfo *mi = fp->mbi + n;
    const int type = mi->type;
    if (type == MBT_SKIP) return 0;
    const int mbx = n % fp->nmbx, mby = n / fp->nmbx;
    const int16_t *coef = fp->coef + (size_t)n * COEF_PER_MB;
    const int cbp = mi->cbp, cbpl = cbp & 15, cbpc = cbp >> 4;
    const int i16 = type == MBT_I16;
    const uint8_t scan8[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};   /* 4x4 block coding order */

    if (is_p) bw_ue(b, (uint32_t)run);
    int mbt = type;
    if (i16) mbt += mi->i16_mode + cbpc * 4 + (cbpl ? 12 : 0);
    if (mbt >= 5 && !is_p) mbt -= 5;
    bw_ue(b, (uint32_t)mbt);
    if (type == 3) bw_put(b, 4, 0xF);      /* four sub_mb_type = 8x8 -> ue(0) each */
    if (type >= 5)
    {
        if (type == MBT_I4)
            for (int i = 0; i < 16; i++)
            {
                int m = mi->i4_code[scan8[i]];
                if (m < 0) bw_put(b, 1, 1); else bw_put(b, 4, (uint32_t)m);
            }
        int cm = mi->i16_mode;
        if (!(cm & 1)) cm ^= 2;
        bw_ue(b, (uint32_t)cm);
    } else
    {
        int nparts = type == 0 ? 1 : (type == 3 ? 4 : 2);
        for (int p = 0; p < nparts; p++) { bw_se(b, mv_x(mi->mvd[p])); bw_se(b, mv_y(mi->mvd[p])); }
    }
    if (!i16) bw_ue(b, (uint32_t)cbp_code(type < 5, cbp));
    if (cbp || i16) bw_put(b, 1, 1);       /* mb_qp_delta = 0: QP is constant inside a frame */

    if (i16)
        cavlc_block(b, coef + COEF_YDC, 16, nnz_left(fp, mi, mbx, 3, 0, 1), nnz_top(fp, mi, mby, 12, 0, 1));
    if (cbpl)
        for (int i = 0; i < 16; i++)
        {
            int j = scan8[i];
            if (!(cbp & (1 << (i >> 2)))) continue;
            int x = j & 3, y = j >> 2;
            int nA = nnz_left(fp, mi, mbx, y * 4 + 3, j - 1, x == 0);
            int nB = nnz_top(fp, mi, mby, 12 + x, j - 4, y == 0);
            cavlc_block(b, coef + COEF_Y + j * 16 + i16, 16 - i16, nA, nB);
        }
    if (cbpc)
    {
        cavlc_block(b, coef + COEF_CDC, 4, 17, 17);
        cavlc_block(b, coef + COEF_CDC + 4, 4, 17, 17);
        if (cbpc > 1)
            for (int pl = 0; pl < 2; pl++)
                for (int k = 0; k < 4; k++)
                {
                    int x = k & 1, y = k >> 1, base = 16 + pl * 4;
                    int nA = nnz_left(fp, mi, mbx, base + y * 2 + 1, base + k - 1, x == 0);
                    int nB = nnz_top(fp, mi, mby, base + 2 + x, base + k - 2, y == 0);
                    cavlc_block(b, coef + COEF_C + (pl * 4 + k) * 16 + 1, 15, nA, nB);
                }
    }
    bw_flush(b);
    return bw_bits(b);
}

/* scatter the bit string of slot n to bit offset `bo` of the slice payload.  The payload
 * must have been zeroed; neighbouring strings share boundary words, hence the atomic OR. */
HD void pack_mb(const FrameParams *fp, int n, int nbits, int bo)
{
    const uint32_t *src = fp->mb_bits + (size_t)n * MB_BITS_WORDS;
    int nw = (nbits + 31) >> 5;
    int sh = bo & 31;
    uint32_t *dst = fp->out_words + (bo >> 5);
    for (int k = 0; k < nw; k++)
    {
        uint32_t wv = src[k];
        int rem = nbits - 32 * k;
        if (rem < 32) wv &= ~(0xffffffffu >> rem);
        uint32_t hi = wv >> sh, lo = sh ? wv << (32 - sh) : 0;
#if H264_DEVICE
        if (hi) atomicOr(dst + k, hi);
        if (lo) atomicOr(dst + k + 1, lo);
#else
        dst[k] |= hi;
        if (lo) dst[k + 1] |= lo;
#endif
    }
}
