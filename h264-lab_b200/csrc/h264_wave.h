/*
 * h264_wave.h -- exact wavefront encoding of P frames despite the reference's
 * raster-order "mv_clusters" chain (SURVEY.md section 0 finding 2, section 7 hard part 1).
 *
 * The reference updates two running MV averages after EVERY macroblock in raster order
 * (mv_clusters_update H:5263, called at H:5776-5779) and offers their rounded values as
 * the last two motion-search start candidates of the NEXT macroblock (H:5382-5383).  A
 * wavefront cannot know them.  Scheme used here (all on the GPU):
 *
 *   predict  before the sweep, one warp replays the cluster update over the PREVIOUS frame's
 *            motion field starting from this frame's true start state: the predicted
 *            trajectory (motion fields are temporally coherent);
 *   pass 0   every macroblock is decided in x+2y wavefront order with the predicted cluster
 *            candidates (frame-start value after an intra frame);
 *   replay   the follower (sweep 0) / one warp (later sweeps) replays the true cluster
 *            trajectory in raster order from the macroblocks' final (type, mv[0]) and counts
 *            the macroblocks whose speculated candidates differ from the true ones ("dirty");
 *   pass k   another wavefront sweep: a macroblock is revisited only if it is dirty or a
 *            causal neighbour (L, T, TL, TR) changed in this sweep.  A dirty macroblock
 *            first re-runs just the candidate stage: if (mv_best, sad_best, cost_best,
 *            partition hints) are unchanged, so is everything that follows; otherwise the
 *            macroblock is re-encoded and compared with its previous record to decide
 *            whether neighbours must follow;
 *   ...      until a replay finds no dirty macroblock.  Then every macroblock has been
 *            decided with exactly the candidates the raster-order reference would have
 *            used, by induction over raster order; each sweep extends the exact prefix,
 *            so the loop terminates.
 */
#pragma once
#include "h264_common.h"
#include "h264_mbenc.h"

/* Repair of pass p: macroblocks whose re-check failed are re-encoded in REPAIR_ROUNDS (3) fully
 * parallel rounds (round r re-encodes everything tagged REPAIR_TAG(p, r), with whatever
 * neighbour data is current, and tags the causal successors of every macroblock whose
 * neighbour-visible result changed for round r + 1 -- a fixpoint iteration that ends when a round
 * changes nothing); what is still tagged after the last round goes through one wavefront
 * sweep, which follows cascades of any length. */
#ifndef REPAIR_ROUNDS
#define REPAIR_ROUNDS 3
#endif
#define REPAIR_TAG(pass, r) ((pass) * 8 + (r))
/* rounds of the speculative motion-estimation pre-pass (me_prepass_mb): round 0 + refinements */
#ifndef ME_ROUNDS
#define ME_ROUNDS 3
#endif

HD void spec_store(const FrameParams *fp, int n, const MBSpec &sp)
{
    IF_THREAD0 { fp->spec[n] = sp; }
}


/* ------------------------------------------------------------------------------
 * Speculative motion estimation AHEAD of the wavefront.
 *
 * What the motion estimation of a macroblock (candidate stage + searches, me_phase) depends on, besides the pictures, is
 * small: the 13 vectors of its MV-predictor context and the two cluster candidates.  With the SAD maps in place
 * the estimation is pure table work, so it can be run for EVERY macroblock of the frame at once, before the wavefront
 * starts, on PREDICTED inputs: round 0 takes the context from the previous frame's motion field (the record array
 * still holds it), every further round takes it from the field the previous round predicted for this frame (a Jacobi
 * iteration that settles after a round or two where motion is coherent) and only re-runs macroblocks whose
 * context has changed.  Each result is stored next to the macroblock's SAD maps together with the inputs it was
 * computed from.  The wavefront (encode_mb) compares those inputs with the real ones: equal -- identical inputs,
 * identical function -- means the record IS what the estimation would produce now, and the macroblock skips it;
 * different means the estimation runs in place, as before.  Nothing is ever trusted without that comparison, so the
 * decisions stay the reference's; only the critical path of the wavefront gets shorter.
 * ---------------------------------------------------------------------------- */
/* one thread: does the record of macroblock n still belong to the context the predicted field gives it?  (refinement rounds) */
HD int me_record_stale(const FrameParams *fp, int n)
{
    const int nmbx = fp->nmbx, y = n / nmbx, x = n - y * nmbx, av = mb_avail(x, y, nmbx);
    const uint32_t *mr = fp->sadmap + (size_t)n * SM_WORDS + SM_ME_OFF;
    if (mr[ME_KEY + 15] != 1u) return 1;
    const int32_t *f = fp->me_field;
    for (int i = 0; i < 15; i++)
    {
        int32_t have;
        if (i < 4) have = (av & AVAIL_L) ? f[(n - 1) * 16 + 4 * i + 3] : MV_NA;
        else if (i == 4) have = (av & AVAIL_TL) ? f[(n - nmbx - 1) * 16 + 15] : MV_NA;
        else if (i < 8) have = (av & AVAIL_L) ? f[(n - 1) * 16 + 4 * (i - 5) + 3] : MV_NA;
        else if (i < 12) have = (av & AVAIL_T) ? f[(n - nmbx) * 16 + 12 + (i - 8)] : MV_NA;
        else if (i == 12) have = (av & AVAIL_TR) ? f[(n - nmbx + 1) * 16 + 12] : MV_NA;
        else have = fp->spec_from_prev ? fp->cl_true[2 * n + (i - 13)] : mv_round_fullpel(fp->clusters[i - 13]);
        if ((int32_t)mr[ME_KEY + i] != have) return 1;
    }
    return 0;
}

HDN void me_prepass_mb(const FrameParams *fp, MBWork *w, int x, int y, int round)
{
    const int nmbx = fp->nmbx, n = y * nmbx + x;
    uint32_t *mr = fp->sadmap + (size_t)n * SM_WORDS + SM_ME_OFF;
    if (round > 0 && mr[ME_KEY + 15] == 1u)
    {   /* refinement rounds: most records already belong to the context the field now predicts -- check that first, it
         * takes 15 loads and no macroblock set-up */
        const int av = mb_avail(x, y, nmbx);
        const int32_t *f = fp->me_field;
        int same = 1;
        FOR_LANES(i, 15)
        {
            int32_t have;
            if (i < 4) have = (av & AVAIL_L) ? f[(n - 1) * 16 + 4 * i + 3] : MV_NA;
            else if (i == 4) have = (av & AVAIL_TL) ? f[(n - nmbx - 1) * 16 + 15] : MV_NA;
            else if (i < 8) have = (av & AVAIL_L) ? f[(n - 1) * 16 + 4 * (i - 5) + 3] : MV_NA;
            else if (i < 12) have = (av & AVAIL_T) ? f[(n - nmbx) * 16 + 12 + (i - 8)] : MV_NA;
            else if (i == 12) have = (av & AVAIL_TR) ? f[(n - nmbx + 1) * 16 + 12] : MV_NA;
            else have = fp->spec_from_prev ? fp->cl_true[2 * n + (i - 13)] : mv_round_fullpel(fp->clusters[i - 13]);
            if ((int32_t)mr[ME_KEY + i] != have) same = 0;
        }
#if H264_DEVICE
        same = __all_sync(0xffffffffu, same);          /* the pre-pass runs one warp per macroblock */
#endif
        if (same) return;
    }
    MBState s;
    s.fp = fp; s.w = w; s.mbx = x; s.mby = y;
    s.avail = mb_avail(x, y, nmbx);
    s.type = 0; s.cost = 0x7FFFFFFF; s.i16_mode = 2; s.mv_skip_pred = 0;
    s.pbest = w->skip_pred; s.ss = &w->ss[0];
    s.win_ok = 0; s.win_x0 = s.win_y0 = 0;
    s.map = 0; s.lut = 0;
    mb_load(s, 1);                               /* input samples, SAD maps, MV context from the record array = the PREVIOUS frame's field */
    if (round > 0)
    {   /* context from the field predicted for this frame (same availability rules as mb_load) */
        const int av = s.avail;
        const int32_t *f = fp->me_field;
        FOR_THREADS(i, 13)
        {
            if (i < 4) w->mvp0_left[i] = (av & AVAIL_L) ? f[(n - 1) * 16 + 4 * i + 3] : MV_NA;
            else if (i == 4) w->mvp0_tl[0] = (av & AVAIL_TL) ? f[(n - nmbx - 1) * 16 + 15] : MV_NA;
            else if (i < 8) w->mvp0_tl[i - 4] = (av & AVAIL_L) ? f[(n - 1) * 16 + 4 * (i - 5) + 3] : MV_NA;
            else if (i < 12) w->mvp0_top[i - 8] = (av & AVAIL_T) ? f[(n - nmbx) * 16 + 12 + (i - 8)] : MV_NA;
            else w->mvp0_top[4] = (av & AVAIL_TR) ? f[(n - nmbx + 1) * 16 + 12] : MV_NA;
        }
        CTA_SYNC();
    }
    int32_t cl[2];
    if (fp->spec_from_prev) { cl[0] = fp->cl_true[2 * n]; cl[1] = fp->cl_true[2 * n + 1]; }         /* what sweep 0 will use (wave_mb_first) */
    else { cl[0] = mv_round_fullpel(fp->clusters[0]); cl[1] = mv_round_fullpel(fp->clusters[1]); }
    if (round > 0 && me_record_matches(s, mr, cl)) return;      /* the record already belongs to this context */
    const int mvp16 = mvp_get(w->mvp0_left, w->mvp0_tl, w->mvp0_top, s.avail, 0, 0, 4, 4);
    s.lut = lut_decide(s, mvp16);
    int32_t pmv[4] = {0, 0, 0, 0}, pmvd[4] = {0, 0, 0, 0};
    int type = 0;
    if (s.lut)
    {
        me_phase(s, cl);
        CTA_SYNC();
        FOR_THREADS(i, 68)
        {
            int32_t v;
            if (i < 4) v = w->mvp0_left[i];
            else if (i < 8) v = w->mvp0_tl[i - 4];
            else if (i < 13) v = w->mvp0_top[i - 8];
            else if (i < 15) v = cl[i - 13];
            else if (i == 15) v = 1;
            else if (i < 32) v = w->ic[i - 16];
            else if (i < 36) v = w->mode_cost[i - 32];
            else if (i < 52) v = (&w->part_mv[0][0])[i - 36];
            else v = (&w->part_mvd[0][0])[i - 52];
            mr[i] = (uint32_t)v;
        }
        int cost, usp;
        inter_decide(fp, w, &type, &cost, pmv, pmvd, &usp);
        /* an inter cost nearly twice the usual: the macroblock will most likely end up intra, i.e. without vectors
         * for its neighbours' context (only a prediction -- the wavefront checks every context) */
        if (type != MBT_SKIP && fp->have_cost_stat && cost >= imax(fp->cost_stat[1], fp->cost_stat[2 + y] * 15 / fp->thr_eighths)) type = -3;
    } else
    {   /* no estimation by look-up here: no record; the field keeps the previous frame's vectors for the neighbours' context */
        IF_THREAD0 { mr[ME_KEY + 15] = 0; }
        type = -2;
    }
    /* the motion field this macroblock is predicted to end up with (an intra outcome cannot be known here) */
    FOR_THREADS(i, 16)
    {
        int v;
        if (type == -2) v = fp->mbi[n].mv[i];
        else if (type == -3) v = MV_NA;
        else if (type <= 0) v = pmv[0];
        else
        {
            const int bx = i & 3, by = i >> 2;
            v = pmv[type == 1 ? (by >> 1) : (type == 2 ? (bx >> 1) : (by >> 1) * 2 + (bx >> 1))];
        }
        fp->me_field[n * 16 + i] = v;
    }
    CTA_SYNC();
}

/* ------------------------------------------------------------------------------
 * Intra modes verified AFTER sweep 0 (P frames, fp->spec_no_intra).
 *
 * The Intra16x16 / Intra4x4 costs of a macroblock (H:4876, H:4723) need the unfiltered reconstruction of its left and
 * upper neighbours, and Intra4x4 is a chain of 16 dependent block decisions: on the wavefront that chain was the
 * longest thing a P macroblock did, although an intra mode wins for about one macroblock in a hundred.  Sweep 0
 * therefore decides every macroblock among the inter modes only and records the cost to beat; once the sweep is
 * over, the neighbours' reconstruction exists everywhere and this check -- independent per macroblock, so fully
 * parallel -- evaluates the intra costs exactly as encode_mb would (same exact pruning against the inter cost).  Where
 * an intra mode wins, the macroblock is queued for a full re-encode in repair pass 1; the repair machinery above
 * re-encodes it with all modes and follows every consequence (successors whose context changed, the cluster
 * trajectory).  A macroblock whose neighbours change later is re-encoded in full anyway (it is their causal
 * successor), so every final decision has seen the intra modes with the final neighbours.
 * ---------------------------------------------------------------------------- */
HDN void wave_mb_intra_check(const FrameParams *fp, MBWork *w, int x, int y)
{
    const int n = y * fp->nmbx + x;
    const int flags = fp->spec[n].flags;
    /* early skips never look at intra modes (H:5767); only macroblocks decided without them need the check */
    if (!(flags & SPEC_NO_INTRA) || !(flags & SPEC_USED_CL)) return;
    const int inter_cost = fp->spec[n].pad[0];
    MBState s;
    s.fp = fp; s.w = w; s.mbx = x; s.mby = y;
    s.avail = mb_avail(x, y, fp->nmbx);
    s.type = 0; s.cost = inter_cost; s.i16_mode = 2; s.mv_skip_pred = 0;
    s.pbest = w->skip_pred; s.ss = &w->ss[0];
    s.win_ok = 0; s.win_x0 = s.win_y0 = 0;
    s.map = 0; s.lut = 0;
    mb_load(s);
    int win = 0;
    ON_WARP(0)
    {
        const pix_t *left = (s.avail & AVAIL_L) ? w->left_y : 0;
        const pix_t *top = (s.avail & AVAIL_T) ? w->top_y : 0;
        const int m16 = intra16_estimate(w->inp_y, s.avail, fp->qp);
        intra16_pred(w->i16pred, left, top, m16);
        WSYNC();
        const int cost16 = sad_sm_wh(w->inp_y, w->i16pred, 16, 16) + ((bitsize_ue(m16 + 1) * fp->lambda_q4) >> 4) + fp->lambda_i16_q4;
        win = cost16 < inter_cost;
        if (!win && fp->speed < 2)
        {
            int nz4 = 0;
            const int cost4 = intra4_choose(s, &nz4, cost16, inter_cost);
            win = cost4 < inter_cost;
        }
        IF_LANE0 { w->scal[4] = win; }
    }
    CTA_SYNC();
    win = w->scal[4];
    IF_THREAD0
    {
        if (win)
        {
            fp->need_reenc[n] = REPAIR_TAG(1, 0);
            atomic_add_stat(fp->fsync + FS_NFAIL);
        }
    }
    CTA_SYNC();
}

/* pass 0 / I frames */
HDF_wave_mb_first void wave_mb_first(const FrameParams *fp, MBWork *w, int x, int y)
{
    const int n = y * fp->nmbx + x;
    int32_t cl[2] = {0, 0};
    MBSpec sp;
    if (fp->slice_type == SLICE_P)
    {
        /* speculation: the co-located value of the previous P frame's trajectory when there is
         * one (motion fields are temporally coherent), else the frame-start value */
        /* speculation: when the previous frame was a P frame, the trajectory PREDICTED for this
         * frame by replaying the cluster update over the previous frame's motion field from
         * this frame's true start state (wave_replay(predict), motion is temporally coherent);
         * otherwise the frame-start value */
        if (fp->spec_from_prev) { cl[0] = fp->cl_true[2 * n]; cl[1] = fp->cl_true[2 * n + 1]; }
        else { cl[0] = mv_round_fullpel(fp->clusters[0]); cl[1] = mv_round_fullpel(fp->clusters[1]); }
    }
    encode_mb(fp, w, x, y, cl, &sp, !(fp->slice_type == SLICE_P && fp->spec_no_intra));
    if (fp->slice_type == SLICE_P)
    {
        spec_store(fp, n, sp);
        IF_THREAD0 { fp->changed_pass[n] = 0; fp->need_reenc[n] = 0; }
    }
}

/* candidate stage only: returns 1 when its outcome equals the recorded one */
HDN int wave_cand_check(const FrameParams *fp, MBWork *w, int x, int y, const int32_t cl[2], const MBSpec &old)
{
    MBState s;
    s.fp = fp; s.w = w; s.mbx = x; s.mby = y;
    s.avail = mb_avail(x, y, fp->nmbx);
    s.type = 0; s.cost = 0x7FFFFFFF; s.i16_mode = 2; s.mv_skip_pred = 0;
    s.pbest = w->skip_pred; s.ss = &w->ss[0];
    s.win_ok = 0; s.win_x0 = s.win_y0 = 0;
    s.map = 0; s.lut = 0;
    mb_load(s);
    int mvp16 = mvp_get(w->mvp0_left, w->mvp0_tl, w->mvp0_top, s.avail, 0, 0, 4, 4);
    s.lut = lut_decide(s, mvp16);
#if H264_DEVICE && MB_WARPS == 1 && !defined(CHECK_WITH_WINDOW)
    /* single-warp re-check: the candidate stage reads a handful of 16x16 blocks; fetching them straight
     * from the reference picture (ref_at() without a window) moves less data than staging a 64x48 window */
    (void)mvp16;
#else
    if (!s.lut) win_load(s, x * 16 + ((mv_x(mvp16) + 1) >> 2), y * 16 + ((mv_y(mvp16) + 1) >> 2));
#endif
    ON_WARP(0) { inter_stage_a(s, cl); }
    CTA_SYNC();
    int same = 0;
    if (w->ic[IC_STATE] == 2 && w->ic[IC_SIG] == old.cand_sig[0] && w->ic[IC_SIG + 1] == old.cand_sig[1] &&
        w->ic[IC_SIG + 2] == old.cand_sig[2])
    {
        const int newp = w->ic[IC_SIG + 3], oldp = old.cand_sig[3];
        if (newp == oldp) same = 1;
        else
        {
            /* Same start point, different partition hints.  Searches of a mode do not depend on the
             * candidates, so the modes searched before keep their recorded cost; hinted modes that
             * were not searched yet are searched now (same tasks / warps as in encode_mb), and the
             * inter decision (H:5500) is re-run over the hinted set.  Unchanged winner == unchanged
             * macroblock. */
            const int extra = newp & ~oldp;
            if (extra)
            {
                ON_WARP(1)
                {
                    s.ss = &w->ss[SS_SLOT(1)];
                    if (extra & 1) inter_mode_search(s, 1);
                    if (extra & 2) inter_mode_search(s, 2);
                }
                ON_WARP(2)
                {
                    s.ss = &w->ss[SS_SLOT(2)];
                    if (extra & 4) inter_mode_search(s, 3);
                }
                CTA_SYNC();
            }
            int cost = 0xffffff, best = 0;
            for (int t = 0; t < 4; t++)
                if (t == 0 || ((newp >> (t - 1)) & 1))
                {
                    const int ct = (t > 0 && ((extra >> (t - 1)) & 1)) ? w->mode_cost[t] : old.mode_cost[t];
                    if (ct < cost) { cost = ct; best = t; }
                }
            if (best == old.inter_best) same = extra ? 3 : 2;
        }
    }
#if !H264_DEVICE
    {
        extern int g_emu_dbg[8];
        if (!same)
        {
            if (w->ic[IC_STATE] != 2) g_emu_dbg[0]++;
            else if (w->ic[IC_SIG] != old.cand_sig[0] || w->ic[IC_SIG + 1] != old.cand_sig[1] || w->ic[IC_SIG + 2] != old.cand_sig[2]) g_emu_dbg[1]++;
            else { g_emu_dbg[2]++; if ((w->ic[IC_SIG + 3] & old.cand_sig[3]) == old.cand_sig[3]) g_emu_dbg[3]++; }
        } else g_emu_dbg[4]++;
    }
#endif
    CTA_SYNC();
    return same;
}

/* Parallel re-check of one macroblock before repair sweep `pass`: independent of every other
 * macroblock (it only reads sweep results), so all dirty macroblocks of all frames are checked
 * at once.  A dirty macroblock re-runs just the candidate stage with its true candidates; if
 * the outcome is the recorded one nothing downstream can change and the record simply adopts
 * the true candidates, else the macroblock is queued for a re-encode in the sweep. */
HDN void wave_mb_check(const FrameParams *fp, MBWork *w, int x, int y, int pass)
{
    const int n = y * fp->nmbx + x;
    const MBSpec old = fp->spec[n];
    int32_t ct[2];
    ct[0] = fp->cl_true[2 * n]; ct[1] = fp->cl_true[2 * n + 1];
    const int need_cl = (old.flags & SPEC_USED_CL) && (ct[0] != old.cl_used[0] || ct[1] != old.cl_used[1]);
    if (!need_cl) return;
    int chk = wave_cand_check(fp, w, x, y, ct, old);
    IF_THREAD0
    {
        atomic_add_stat(fp->fsync + FS_CHECKS);
        if (chk)
        {
            fp->spec[n].cl_used[0] = ct[0]; fp->spec[n].cl_used[1] = ct[1];
            if (chk >= 2)
            {
                const int newp = w->ic[IC_SIG + 3], extra = newp & ~old.cand_sig[3];
                for (int t = 1; t < 4; t++) if ((extra >> (t - 1)) & 1) fp->spec[n].mode_cost[t] = w->mode_cost[t];
                fp->spec[n].cand_sig[3] = newp;
            }
        } else
        {
            fp->need_reenc[n] = REPAIR_TAG(pass, 0);
            atomic_add_stat(fp->fsync + FS_NFAIL);
#if !H264_DEVICE
            { extern int g_emu_dbg[8]; g_emu_dbg[5]++; }
#endif
        }
    }
    CTA_SYNC();
}

/* repair sweep `pass` (>= 1): re-encode what the re-check queued and whatever depends on a
 * macroblock that changed in this sweep */
/* full re-encode of one macroblock with its true candidates; returns 1 when what its causal
 * successors consume has changed */
HDF_wave_mb_reencode int wave_mb_reencode(const FrameParams *fp, MBWork *w, int x, int y)
{
    const int nmbx = fp->nmbx, n = y * nmbx + x;
    const MBSpec old = fp->spec[n];
    int32_t ct[2];
    ct[0] = fp->cl_true[2 * n]; ct[1] = fp->cl_true[2 * n + 1];
    /* full re-encode; keep the previous record and reconstruction for comparison */
    MBInfo *mi = fp->mbi + n;
    const int sy = fp->stride[0], sc = fp->stride[1];
    const pix_t *dy = fp->dec[0] + (y * 16) * sy + x * 16;
    const pix_t *du = fp->dec[1] + (y * 8) * sc + x * 8, *dv = fp->dec[2] + (y * 8) * sc + x * 8;
    FOR_THREADS(i, 36) { w->old_mbi[i] = ((const uint32_t *)mi)[i]; }
    FOR_THREADS(i, 96)
    {
        uint32_t v;
        if (i < 64) v = *(const uint32_t *)(dy + (i >> 2) * sy + (i & 3) * 4);
        else { int k = i - 64, r = k >> 2, q = k & 3; v = *(const uint32_t *)((q < 2 ? du : dv) + r * sc + (q & 1) * 4); }
        w->old_rec[i] = v;
    }
    CTA_SYNC();
    MBSpec sp;
    encode_mb(fp, w, x, y, ct, &sp);
    /* Only what a causal successor consumes can propagate: the MV grid and the I4x4 modes
     * (MV / mode predictors) and the unfiltered right column / bottom row of the
     * reconstruction (intra prediction).  mvd, cbp, levels only feed this MB's own bits. */
    ON_WARP(0)
    {
        int diff = 0;
        FOR_LANES(i, 36) { if (i < 16 || (i >= 22 && i < 26)) diff |= w->old_mbi[i] != ((const uint32_t *)mi)[i]; }
        FOR_LANES(i, 96)
        {
            uint32_t v, mask;
            if (i < 64) { v = *(const uint32_t *)(dy + (i >> 2) * sy + (i & 3) * 4); mask = ((i >> 2) == 15 ? 0xffffffffu : 0u) | ((i & 3) == 3 ? 0xff000000u : 0u); }
            else
            {
                int k = i - 64, r = k >> 2, q = k & 3;
                v = *(const uint32_t *)((q < 2 ? du : dv) + r * sc + (q & 1) * 4);
                mask = (r == 7 ? 0xffffffffu : 0u) | ((q & 1) ? 0xff000000u : 0u);
            }
            diff |= ((w->old_rec[i] ^ v) & mask) != 0;
        }
        diff = wor(diff);
        IF_LANE0 { w->scal[3] = diff; }
    }
    CTA_SYNC();
    const int diff = w->scal[3];
    spec_store(fp, n, sp);
    IF_THREAD0
    {
        atomic_add_stat(fp->fsync + FS_REENC);
#if !H264_DEVICE
        { extern int g_emu_dbg[8]; g_emu_dbg[6]++; if (diff) g_emu_dbg[7]++; }
#endif
        if (sp.mv0 != old.mv0 || ((sp.flags ^ old.flags) & SPEC_UPDATES))
        {
            atomic_add_stat(fp->fsync + FS_TRAJ_CHANGED);
#if H264_DEVICE
            atomicMax(fp->fsync + FS_TRAJ_FIRST, 0x3fffffff - n);
#else
            if (fp->fsync[FS_TRAJ_FIRST] < 0x3fffffff - n) fp->fsync[FS_TRAJ_FIRST] = 0x3fffffff - n;
#endif
        }
    }
    CTA_SYNC();
    return diff;
}

/* parallel repair round r of pass `pass`: the macroblock is tagged for this round */
HDN void wave_mb_round(const FrameParams *fp, MBWork *w, int x, int y, int pass, int r)
{
    const int nmbx = fp->nmbx, n = y * nmbx + x;
    if (wave_mb_reencode(fp, w, x, y))
    {
        IF_THREAD0
        {
            const int tag = REPAIR_TAG(pass, r + 1);
            if (r + 1 == REPAIR_ROUNDS) atomic_add_stat(fp->fsync + FS_WAVE_TAGS);     /* the wave has something to do */
            if (x < nmbx - 1) fp->need_reenc[n + 1] = tag;
            if (y < fp->nmby - 1)
            {
                if (x > 0) fp->need_reenc[n + nmbx - 1] = tag;
                fp->need_reenc[n + nmbx] = tag;
                if (x < nmbx - 1) fp->need_reenc[n + nmbx + 1] = tag;
            }
        }
    }
    CTA_SYNC();
}

/* wavefront repair sweep of pass `pass` (after the parallel rounds): re-encode what is still
 * tagged and whatever depends on a macroblock that changed in this sweep */
HDN void wave_mb_repair(const FrameParams *fp, MBWork *w, int x, int y, int pass)
{
    const int nmbx = fp->nmbx, n = y * nmbx + x;
    /* all flags with independent loads: one memory round trip on the common "nothing to do" path */
    const int has_l = x > 0, has_t = y > 0, has_tl = y > 0 && x > 0, has_tr = y > 0 && x < nmbx - 1;
    const int f0 = fp->need_reenc[n];
    const int f1 = fp->changed_pass[has_l ? n - 1 : n];
    const int f2 = fp->changed_pass[has_t ? n - nmbx : n];
    const int f3 = fp->changed_pass[has_tl ? n - nmbx - 1 : n];
    const int f4 = fp->changed_pass[has_tr ? n - nmbx + 1 : n];
    const int need = (f0 == REPAIR_TAG(pass, REPAIR_ROUNDS)) | (has_l & (f1 == pass)) | (has_t & (f2 == pass)) | (has_tl & (f3 == pass)) | (has_tr & (f4 == pass));
    if (!need) return;
    if (wave_mb_reencode(fp, w, x, y)) { IF_THREAD0 { fp->changed_pass[n] = pass; } }
    CTA_SYNC();
}

/* Sequential replay of the cluster trajectory by one warp (lane 0 walks, the warp stages
 * 32 records at a time).  Writes cl_true[], the end state, and returns the dirty count. */
HDN int wave_replay(const FrameParams *fp, MBWork *w, int predict, int first_block = 0)
{
    const int nmb = fp->nmbx * fp->nmby;
    int32_t c[2];
    c[0] = fp->clusters[0]; c[1] = fp->clusters[1];
    if (first_block > 0) { c[0] = fp->cl_ckpt[2 * first_block]; c[1] = fp->cl_ckpt[2 * first_block + 1]; }
    int ndirty = 0;
#if H264_DEVICE
    /* every lane replays the (cheap, strictly sequential) recurrence redundantly; the records of
     * 32 macroblocks at a time sit in registers and are broadcast with shuffles, the next 32
     * are prefetched meanwhile */
    {
        const int lane = LANE_ID;
        int mv0 = 0, flags = 0, u0 = 0, u1 = 0;
        long long csum = 0; int ccnt = 0;        /* predict: mean inter cost of the frame (see FrameParams::cost_stat) */
        if (32 * first_block + lane < nmb) { const MBSpec *sp = fp->spec + 32 * first_block + lane; mv0 = sp->mv0; flags = sp->flags; u0 = sp->cl_used[0]; u1 = sp->cl_used[1]; if (predict && (flags & SPEC_USED_CL)) { csum += sp->pad[0]; ccnt++; } }
        for (int base = 32 * first_block; base < nmb; base += 32)
        {
            if (!predict && lane < 2) fp->cl_ckpt[2 * (base >> 5) + lane] = c[lane];
            int nmv0 = 0, nflags = 0, nu0 = 0, nu1 = 0;
            if (base + 32 + lane < nmb)
            {
                const MBSpec *sp = fp->spec + base + 32 + lane;
                nmv0 = sp->mv0; nflags = sp->flags; nu0 = sp->cl_used[0]; nu1 = sp->cl_used[1];
                if (predict && (nflags & SPEC_USED_CL)) { csum += sp->pad[0]; ccnt++; }
            }
            const int cnt = imin(32, nmb - base);
            int t0 = 0, t1 = 0;
#pragma unroll 4
            for (int i = 0; i < cnt; i++)
            {
                const int f = __shfl_sync(0xffffffffu, flags, i), m = __shfl_sync(0xffffffffu, mv0, i);
                const int a0 = __shfl_sync(0xffffffffu, u0, i), a1 = __shfl_sync(0xffffffffu, u1, i);
                const int r0 = mv_round_fullpel(c[0]), r1 = mv_round_fullpel(c[1]);
                if (lane == i) { t0 = r0; t1 = r1; }
                if ((f & SPEC_USED_CL) && (r0 != a0 || r1 != a1)) ndirty++;
                if (f & SPEC_UPDATES) clusters_update(c, m);
            }
            if (lane < cnt) { fp->cl_true[2 * (base + lane)] = t0; fp->cl_true[2 * (base + lane) + 1] = t1; }
            mv0 = nmv0; flags = nflags; u0 = nu0; u1 = nu1;
        }
        if (predict && fp->cost_stat)
        {
#pragma unroll
            for (int o = 16; o; o >>= 1) { csum += __shfl_xor_sync(0xffffffffu, csum, o); ccnt += __shfl_xor_sync(0xffffffffu, ccnt, o); }
            const int thr = ccnt ? (int)((csum / ccnt) * fp->thr_eighths / 8) : 0;
            if (lane == 0) { fp->cost_stat[0] = thr; fp->cost_stat[1] = ccnt ? (int)((csum / ccnt) * 15 / 8) : 0x7FFFFFFF; }
            for (int y = 0; y < fp->nmby; y++)
            {   /* per-row thresholds */
                long long rs = 0; int rc = 0;
                for (int x = lane; x < fp->nmbx; x += 32) { const MBSpec *sp = fp->spec + y * fp->nmbx + x; if (sp->flags & SPEC_USED_CL) { rs += sp->pad[0]; rc++; } }
#pragma unroll
                for (int o = 16; o; o >>= 1) { rs += __shfl_xor_sync(0xffffffffu, rs, o); rc += __shfl_xor_sync(0xffffffffu, rc, o); }
                if (lane == 0) fp->cost_stat[2 + y] = imax(thr, rc ? (int)((rs / rc) * fp->thr_eighths / 8) : 0);
            }
        }
    }
#else
    if (predict && fp->cost_stat)
    {
        long long csum = 0; int ccnt = 0;
        for (int n = 0; n < nmb; n++) if (fp->spec[n].flags & SPEC_USED_CL) { csum += fp->spec[n].pad[0]; ccnt++; }
        fp->cost_stat[0] = ccnt ? (int)((csum / ccnt) * fp->thr_eighths / 8) : 0;
        fp->cost_stat[1] = ccnt ? (int)((csum / ccnt) * 15 / 8) : 0x7FFFFFFF;
        for (int y = 0; y < fp->nmby; y++)
        {
            long long rs = 0; int rc = 0;
            for (int x = 0; x < fp->nmbx; x++) { const MBSpec *sp = fp->spec + y * fp->nmbx + x; if (sp->flags & SPEC_USED_CL) { rs += sp->pad[0]; rc++; } }
            fp->cost_stat[2 + y] = imax(fp->cost_stat[0], rc ? (int)((rs / rc) * fp->thr_eighths / 8) : 0);
        }
    }
    for (int base = 32 * first_block; base < nmb; base += 32)
    {
        if (!predict) { fp->cl_ckpt[2 * (base >> 5)] = c[0]; fp->cl_ckpt[2 * (base >> 5) + 1] = c[1]; }
        FOR_LANES(i, 32)
        {
            int n = base + i;
            if (n < nmb)
            {
                const MBSpec *sp = fp->spec + n;
                w->rp_mv0[i] = sp->mv0; w->rp_flags[i] = sp->flags;
                w->rp_used0[i] = sp->cl_used[0]; w->rp_used1[i] = sp->cl_used[1];
            }
        }
        WSYNC();
        IF_LANE0
        {
            int cnt = imin(32, nmb - base);
            for (int i = 0; i < cnt; i++)
            {
                int r0 = mv_round_fullpel(c[0]), r1 = mv_round_fullpel(c[1]);
                w->rp_true0[i] = r0; w->rp_true1[i] = r1;
                int f = w->rp_flags[i];
                if ((f & SPEC_USED_CL) && (r0 != w->rp_used0[i] || r1 != w->rp_used1[i])) ndirty++;
                if (f & SPEC_UPDATES) clusters_update(c, w->rp_mv0[i]);
            }
        }
        WSYNC();
        FOR_LANES(i, 32)
        {
            int n = base + i;
            if (n < nmb) { fp->cl_true[2 * n] = w->rp_true0[i]; fp->cl_true[2 * n + 1] = w->rp_true1[i]; }
        }
        WSYNC();
    }
#endif
    if (predict) return 0;
    IF_LANE0
    {
        fp->fsync[FS_CL_END] = c[0]; fp->fsync[FS_CL_END + 1] = c[1];
        fp->fsync[FS_NDIRTY] = ndirty;
        w->scal[0] = ndirty;
    }
    WSYNC();
    return w->scal[0];
}

/* After the parallel re-check for sweep `pass`: when no macroblock needs a re-encode, every
 * record is consistent with the replayed trajectory and the frame is exact. Returns the state. */
HD int wave_after_check(const FrameParams *fp, int pass)
{
    if (fp->fsync[FS_STATE] != pass) return fp->fsync[FS_STATE];
    if (fp->fsync[FS_NFAIL] == 0)
    {
        fp->clusters[0] = fp->fsync[FS_CL_END]; fp->clusters[1] = fp->fsync[FS_CL_END + 1];
        fp->fsync[FS_PASSES] = pass;
        fp->fsync[FS_STATE] = FS_DONE;
        return FS_DONE;
    }
    fp->fsync[FS_NFAIL] = 0;
    return pass;
}

/* Executed once per pass by the last row to finish: returns the next pass number or FS_DONE.
 * On FS_DONE the cluster state is committed for the next frame. */
HDN int wave_end_of_pass(const FrameParams *fp, MBWork *w, int pass)   /* one warp */
{
    int next;
    if (fp->slice_type != SLICE_P) return FS_DONE;
    int need_replay = pass == 0 || fp->fsync[FS_TRAJ_CHANGED] != 0;
    WSYNC();
    if (need_replay)
    {
        /* the trajectory is unchanged up to the first macroblock that changed: resume at its block */
        const int tf = fp->fsync[FS_TRAJ_FIRST];
        const int first_block = (pass > 0 && tf > 0) ? (0x3fffffff - tf) >> 5 : 0;
        WSYNC();
        /* GPU: the follower of the repair wave may have replayed already (same resume point) */
        const int replayed = pass > 0 && fp->fsync[FS_REPLAYED] == pass && fp->fsync[FS_REPLAY_TF] == tf;
        const int nd_follower = fp->fsync[FS_NDIRTY];
        WSYNC();
        IF_LANE0 { fp->fsync[FS_TRAJ_CHANGED] = 0; fp->fsync[FS_TRAJ_FIRST] = 0; fp->fsync[FS_WAVE_TAGS] = 0; }
        int nd = replayed ? nd_follower : wave_replay(fp, w, 0, first_block);
        next = nd ? pass + 1 : FS_DONE;
        if (pass == 0 && fp->spec_no_intra) next = 1;       /* the intra verification may still queue re-encodes: wave_after_check(1) decides */
    } else next = FS_DONE;
    if (next == FS_DONE)
    {
        IF_LANE0 { fp->clusters[0] = fp->fsync[FS_CL_END]; fp->clusters[1] = fp->fsync[FS_CL_END + 1]; fp->fsync[FS_PASSES] = pass + 1; }
    }
    WSYNC();
    return next;
}
