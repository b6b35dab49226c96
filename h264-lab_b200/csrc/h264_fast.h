/*
 * h264_fast.h -- fast path of sweep 0 of a P frame: DECIDE serially, WORK in parallel.
 *
 * After the two pre-passes (h264_sadmap.h, h264_wave.h me_prepass_mb) the only thing a P macroblock still needs from
 * its neighbours is their final vectors -- to check that its motion-estimation record was computed from the right
 * context -- and with the intra modes verified after the sweep (wave_mb_intra_check) nothing it computes is needed by
 * a neighbour except its own vectors.  A row therefore advances in two kinds of steps:
 *
 *   decide   ONE warp walks up to FAST_BATCH consecutive macroblocks: builds the 13-vector context (left: the previous
 *            macroblock of the row, kept in shared memory; top: the records of the row above), compares it with the
 *            record's key, takes the inter decision from the record (inter_decide_p: a few dozen scalar instructions),
 *            writes the vectors / types / speculation record and publishes "decided" progress -- the row below can go on.
 *            The walk stops at a macroblock that cannot be decided this way (no record, different key, or an inter
 *            cost so high that an intra mode may win): that one goes through the complete encode_mb by the whole CTA.
 *   work     the warps of the CTA take one decided macroblock each and do everything else for it: input samples,
 *            luma / chroma prediction from the vectors, transform / quantisation / reconstruction, coded block
 *            pattern, skip rollback, total_coeff, levels.  No dependency on any other macroblock.
 *
 * The wavefront's serial chain per macroblock shrinks to the decide step; the pixel work runs four macroblocks wide.
 * Results are the same as encode_mb's for the same inputs: fast_decide + fast_work are encode_mb restricted to
 * "record hit, inter mode, no intra evaluation" (H:5724-5812 with inter_choose_mode's outcome given).
 */
#pragma once
#include "h264_common.h"
#include "h264_mbenc.h"
#include "h264_wave.h"

#if MB_WARPS != 1
/* one warp: input samples of macroblock (x, y) into t->inp_y / t->inp_c, with the edge replication of cropped pictures
 * (pix_copy_cropped_mb H:3536) */
HD void fast_load_input(const FrameParams *fp, TQBuf *t, int x, int y)
{
    const int wv = fp->width, hv = fp->height;
    FOR_LANES(i, 96)
    {
        if (i < 64) *(uint32_t *)(t->inp_y + (i >> 2) * 16 + (i & 3) * 4) = sadmap_inp_word(fp, x, y, i >> 2, i & 3);
        else
        {
            const int k2 = i - 64, r = k2 >> 2, q = k2 & 3, pl = q >> 1, c = (q & 1) * 4;
            const int cx = x * 8 + c, cy = y * 8 + r;
            uint32_t v;
            if (cx + 4 <= wv / 2 && cy < hv / 2) v = ld4u(fp->inp[1 + pl] + (long)cy * fp->inp_stride[1 + pl] + cx);
            else
            {
                const pix_t *row = fp->inp[1 + pl] + (long)imin(cy, hv / 2 - 1) * fp->inp_stride[1 + pl];
                v = 0;
                for (int q2 = 0; q2 < 4; q2++) v |= (uint32_t)row[imin(cx + q2, wv / 2 - 1)] << (8 * q2);
            }
            *(uint32_t *)(t->inp_c + (k2 >> 2) * 16 + (k2 & 3) * 4) = v;
        }
    }
}

/* one warp.  mr = the macroblock's motion-estimation record (shared-memory copy or where it lies).  1 = decided: vectors,
 * speculation record and w->fd[slot] written (the caller publishes the progress); 0 = take the complete path. */
HDN int fast_decide(const FrameParams *fp, MBWork *w, int x, int y, const uint32_t *mr, int slot, int thr)
{
    const int nmbx = fp->nmbx, n = y * nmbx + x, av = mb_avail(x, y, nmbx);
#if !H264_DEVICE
    extern int g_emu_reason;            /* emulation statistics: why the macroblock leaves the fast path (0: it does not) */
    g_emu_reason = 0;
    if (mr[ME_KEY + 15] != 1u) g_emu_reason = 1;
#endif
    if (mr[ME_KEY + 15] != 1u) return 0;
    const MBInfo *mbi = fp->mbi + n;
    int ok = 1;
    FOR_LANES(i, 16)
    {
        /* where the lane's context word lies (no memory access in the cases), then ONE load: the lanes' loads of the row
         * above and of the cluster candidates are in flight together */
        int32_t have = MV_NA;
        const int32_t *src = 0;
        if (i < 4) { if (av & AVAIL_L) have = w->last_mv[4 * i + 3]; }
        else if (i == 4) { if (av & AVAIL_TL) src = &mbi[-nmbx - 1].mv[15]; }
        else if (i < 8) { if (av & AVAIL_L) have = w->last_mv[4 * (i - 5) + 3]; }
        else if (i < 12) { if (av & AVAIL_T) src = &mbi[-nmbx].mv[12 + (i - 8)]; }
        else if (i == 12) { if (av & AVAIL_TR) src = &mbi[-nmbx + 1].mv[12]; }
        else if (i < 15) { if (fp->spec_from_prev) src = &fp->cl_true[2 * n + (i - 13)]; else have = mv_round_fullpel(fp->clusters[i - 13]); }
        else have = 1;
        if (src) have = *src;
        if ((int32_t)mr[ME_KEY + i] != have)
        {
#if !H264_DEVICE
            { extern long g_emu_miss[20]; if (ok) { g_emu_miss[i]++; g_emu_reason = 2 + i + (((int32_t)mr[ME_KEY + i] == MV_NA) ? 100 : (have == MV_NA ? 200 : 0)); if ((int32_t)mr[ME_KEY + i] == MV_NA || have == MV_NA) g_emu_miss[16]++; } }
#endif
            ok = 0;
        }
    }
#if H264_DEVICE
    ok = __all_sync(0xffffffffu, ok);
#endif
    if (!ok) return 0;
    const int32_t *ic = (const int32_t *)(mr + ME_IC), *mc = (const int32_t *)(mr + ME_COST);
    int type, cost, usp;
    int32_t pmv[4] = {0, 0, 0, 0}, pmvd[4] = {0, 0, 0, 0};
    const int best_type = inter_decide_p(fp, ic, mc, (const int32_t *)(mr + ME_MV), (const int32_t *)(mr + ME_MVD), &type, &cost, pmv, pmvd, &usp);
    const int searched = ic[IC_STATE] != 1;
    /* an inter cost this far above the usual: an intra mode probably wins, look at it now (encode_mb) rather than repair later */
#if !H264_DEVICE
    if (searched && cost >= thr) g_emu_reason = 20;
#endif
    if (searched && cost >= thr) return 0;           /* thr: FrameParams::cost_stat of the row (the caller reads it once per row) */
    MBInfo *mi = fp->mbi + n;
    uint32_t *sp = (uint32_t *)(fp->spec + n);
    FOR_LANES(i, 32)
    {
        /* what the neighbours and the later passes read of a decided macroblock: the record's vector part ... */
        if (i < 16)
        {
            int v;
            if (type <= 0) v = pmv[0];
            else { const int bx = i & 3, by = i >> 2; v = pmv[type == 1 ? (by >> 1) : (type == 2 ? (bx >> 1) : (by >> 1) * 2 + (bx >> 1))]; }
            mi->mv[i] = v;
            w->last_mv[i] = v;
        } else if (i < 20) ((uint32_t *)mi->i4_mode)[i - 16] = 0x02020202u;
        else if (i < 24) ((uint32_t *)mi->i4_code)[i - 20] = 0;
        else if (i < 28) mi->mvd[i - 24] = pmvd[i - 24];
        else if (i == 28) { mi->type = (int8_t)type; mi->i16_mode = 2; mi->cbp = 0; mi->flags = 0; mi->nz_mask = 0; mi->pad0 = 0; }
        else if (i == 29) { fp->changed_pass[n] = 0; fp->need_reenc[n] = 0; }
        else if (i == 30) { w->fd[slot].type = type; w->fd[slot].mv_skip = ic[IC_MV_SKIP]; w->fd[slot].x = x; for (int k = 0; k < 4; k++) w->fd[slot].pmv[k] = pmv[k]; }
        /* ... and the speculation record (MBSpec, 16 words; same contents as encode_mb's spec_out) */
        if (i < 16)
        {
            uint32_t v = 0;
            if (i == 0) v = (uint32_t)pmv[0];
            else if (i == 1) v = SPEC_UPDATES | (searched ? SPEC_USED_CL : 0) | SPEC_NO_INTRA;
            else if (i < 4) v = (uint32_t)mv_round_fullpel((int32_t)mr[ME_KEY + 13 + (i - 2)]);
            else if (i < 8) v = searched ? (uint32_t)ic[IC_SIG + (i - 4)] : 0;
            else if (i < 12) v = (searched && ((ic[IC_PREF] >> (i - 8)) & 1)) ? (uint32_t)mc[i - 8] : 0x7FFFFFFFu;
            else if (i == 12) v = searched ? (uint32_t)best_type : 0;
            else if (i == 13) v = (uint32_t)cost;
            sp[i] = v;
        }
    }
    WSYNC();
    return 1;
}

/* one warp, private buffers w->wb[wid] / w->wpred[wid]: everything else for decided macroblock (x, y) */
HDN void fast_work(const FrameParams *fp, MBWork *w, int x, int y, int slot, int wid)
{
    TQBuf *t = &w->wb[wid];
    pix_t *pred = w->wpred[wid];
    const int n = y * fp->nmbx + x;
    int type = w->fd[slot].type;
    const int type0 = type, mv_skip = w->fd[slot].mv_skip;
    int32_t pmv[4];
    for (int k = 0; k < 4; k++) pmv[k] = w->fd[slot].pmv[k];
    MBState s;
    s.fp = fp; s.w = (MBWork *)t;          /* the functions called below only touch the TQBuf part (h264_common.h) */
    s.mbx = x; s.mby = y; s.avail = mb_avail(x, y, fp->nmbx);
    s.type = type; s.cost = 0; s.i16_mode = 2; s.mv_skip_pred = mv_skip;
    s.pbest = pred; s.ss = 0; s.win_ok = 0; s.win_x0 = s.win_y0 = 0; s.map = 0; s.lut = 0;
    fast_load_input(fp, t, x, y);
    luma_pred_half(s, 0, type, pmv, pred);
    luma_pred_half(s, 1, type, pmv, pred);
    mc_chroma_plane(s, 0, type, pmv);
    mc_chroma_plane(s, 1, type, pmv);
    WSYNC();
    int cbpl = 0, cbpc = 0, nz_mask = 0;
    const int sy = fp->stride[0], sc = fp->stride[1];
    if (type != MBT_SKIP)
    {
#if H264_DEVICE
        luma_tq_fast(s, 0, 0);
        luma_tq_fast(s, 1, 0);
        chroma_tq_fast(s, 0);
        chroma_tq_fast(s, 1);
#else
        luma_tq_half(s, 0, 0, 0); luma_tq_half(s, 1, 0, 0);
        luma_tq_half(s, 0, 0, 2); luma_tq_half(s, 1, 0, 2);
        chroma_tq_plane(s, 0); chroma_tq_plane(s, 1);
#endif
        WSYNC();
        nz_mask = t->tq_res[0] | t->tq_res[1];
        if (nz_mask & 0xCC00) cbpl |= 1;
        if (nz_mask & 0x3300) cbpl |= 2;
        if (nz_mask & 0x00CC) cbpl |= 4;
        if (nz_mask & 0x0033) cbpl |= 8;
        for (int pl = 0; pl < 2; pl++)
        {
            if (t->tq_res[2 + pl] & 0xFF) cbpc = 2;
            cbpc |= t->tq_res[2 + pl] >> 8;
        }
        cbpc = imin(cbpc, 2);
        if (!(type | cbpl | cbpc) && pmv[0] == mv_skip) type = MBT_SKIP;       /* rollback H:4494: the reconstruction already is the prediction */
    }
    if (type0 == MBT_SKIP)
    {   /* early skip: reconstruction = prediction (H:4417-4420) */
        pix_t *decy = fp->dec[0] + (y * 16) * sy + x * 16;
        pix_t *du = fp->dec[1] + (y * 8) * sc + x * 8, *dv = fp->dec[2] + (y * 8) * sc + x * 8;
        FOR_LANES(i, 96)
        {
            if (i < 64) { const int r = i >> 2, c = (i & 3) * 4; *(uint32_t *)(decy + r * sy + c) = ld4_sm(pred + r * 16 + c); }
            else
            {
                const int k = i - 64, r = k >> 2, q = k & 3;
                pix_t *d = (q < 2 ? du : dv) + r * sc + (q & 1) * 4;
                *(uint32_t *)d = ld4_sm(t->predc + r * 16 + q * 4);
            }
        }
    }
    if (type == MBT_SKIP) { nz_mask = 0; cbpl = cbpc = 0; }
    MBInfo *mi = fp->mbi + n;
    FOR_LANES(i, 32)
    {
        if (i < 16)
        {
            const int grp = (i >> 3) * 2 + ((i & 3) >> 1);
            mi->nnz[i] = (uint8_t)((type != MBT_SKIP && (cbpl & (1 << grp))) ? count_nz(t->qv_y[i], 0) : 0);
        } else if (i < 24) mi->nnz[i] = (uint8_t)((type != MBT_SKIP && cbpc == 2) ? count_nz(t->qv_c[i - 16], 1) : 0);
        else if (i == 28)
        {
            mi->type = (int8_t)type; mi->i16_mode = 2; mi->cbp = (uint8_t)(cbpl | (cbpc << 4)); mi->flags = 0;
            mi->nz_mask = (uint16_t)nz_mask; mi->pad0 = 0;
        }
    }
    if (type != MBT_SKIP)
    {   /* quantised levels for the entropy-coding pass (mb_store_coefs) */
        int16_t *coef = fp->coef + (size_t)n * COEF_PER_MB;
        FOR_LANES(i, COEF_PER_MB / 2)
        {
            const int k = i * 2;
            uint32_t v;
            if (k < COEF_YDC) v = *(const uint32_t *)(&t->qv_y[0][0] + k);
            else if (k < COEF_C) v = 0;                                          /* luma DC levels: Intra16x16 only */
            else if (k < COEF_CDC) v = *(const uint32_t *)(&t->qv_c[0][0] + (k - COEF_C));
            else v = *(const uint32_t *)(t->qdc_c + (k - COEF_CDC));
            *(uint32_t *)(coef + k) = v;
        }
    }
    WSYNC();
}
#endif
