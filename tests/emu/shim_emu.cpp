/*
 * tests/emu/shim_emu.cpp -- DEVELOPER / TEST TOOL, NOT PART OF THE PRODUCT.
 *
 * Host emulation of the C-ABI shim (include/h264b200_shim.h): compiles the very same
 * macroblock code that nvcc compiles for sm_100a (h264-lab_b200/csrc/h264_*.h) with
 * g++, running every "warp" as a sequential loop, macroblocks in raster order.  It
 * exists so that bit-exactness against the reference can be debugged in a container
 * without a GPU and so that the host C layer (rate control, headers, NAL) can be
 * tested on CPU.  The product library (libh264lab_b200.so) never links this file
 * and has no CPU fallback.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
int g_emu_dbg[8];
long g_emu_lut[8];
long g_emu_miss[20];
long g_emu_qhist[16];
long g_emu_ihist[16];
extern "C" long *emu_ihist(void) { return g_emu_ihist; }
extern "C" long *emu_qhist(void) { return g_emu_qhist; }
int g_emu_reason = 0;
extern "C" long *emu_miss_stats(void) { return g_emu_miss; }
extern "C" long *emu_lut_stats(void) { return g_emu_lut; }
extern "C" int *emu_dbg(void) { return g_emu_dbg; }
#include "../../h264-lab_b200/csrc/h264_common.h"
#include "../../h264-lab_b200/csrc/h264_pixel.h"
#include "../../h264-lab_b200/csrc/h264_mbenc.h"
#include "../../h264-lab_b200/csrc/h264_wave.h"
#include "../../h264-lab_b200/csrc/h264_fast.h"
#include "../../h264-lab_b200/csrc/h264_cavlc.h"
#include "../../h264-lab_b200/csrc/h264_deblock.h"
#include "../../h264-lab_b200/csrc/h264_denoise.h"
#include "../../include/h264b200_shim.h"

struct h264b200_ctx
{
    int width, height, nmbx, nmby;
    int stride[2];
    std::vector<pix_t> frame[2];     /* two padded frames: [cur] = dec, [cur^1] = ref */
    std::vector<pix_t> hpel;         /* half-sample planes b, h, j of the reference picture */
    size_t luma_bytes;
    size_t plane_off[3];
    int cur, last_dec;
    std::vector<MBInfo> mbi;
    std::vector<int16_t> coef;
    std::vector<uint32_t> mb_bits;
    std::vector<int> mb_nbits;
    std::vector<uint32_t> out_words;
    std::vector<pix_t> clip;
    std::vector<pix_t> dn[2]; int dn_cur;   /* temporal noise suppressor: previous / new filtered picture */
    std::vector<MBSpec> spec;
    std::vector<int32_t> cl_true, cl_ckpt;
    std::vector<int> changed_pass, need_reenc;
    std::vector<uint32_t> sadmap;
    std::vector<int32_t> me_field;
    int fsync[FS_WORDS];
    int stats[4];
    int32_t clusters[2];
    std::vector<int> cost_stat; int cost_stat_valid = 0;
    int have_traj;
};

static long g_launches = 0;

extern "C" int h264b200_ctx_create(h264b200_ctx **out, int width, int height, int device)
{
    (void)device;
    h264b200_ctx *c = new h264b200_ctx();
    c->width = width; c->height = height;
    c->nmbx = (width + 15) >> 4; c->nmby = (height + 15) >> 4;
    int w = c->nmbx * 16, h = c->nmby * 16;
    c->stride[0] = w + 32; c->stride[1] = (w + 32) / 2;
    size_t ysz = (size_t)c->stride[0] * (h + 32), csz = (size_t)c->stride[1] * (h / 2 + 16);
    c->plane_off[0] = (size_t)c->stride[0] * 16 + 16;
    c->plane_off[1] = ysz + (size_t)c->stride[1] * 8 + 8;
    c->plane_off[2] = ysz + csz + (size_t)c->stride[1] * 8 + 8;
    for (int i = 0; i < 2; i++) c->frame[i].assign(ysz + 2 * csz + 64, 0);
    c->luma_bytes = ysz;
    c->hpel.assign(3 * ysz + 64, 0);
    int nmb = c->nmbx * c->nmby;
    c->mbi.resize(nmb);
    c->coef.resize((size_t)nmb * COEF_PER_MB);
    c->mb_bits.resize((size_t)(nmb + 1) * MB_BITS_WORDS);
    c->mb_nbits.resize(nmb + 2);
    c->spec.resize(nmb); c->cl_true.resize(2 * nmb); c->cl_ckpt.resize(2 * (nmb / 32 + 2)); c->changed_pass.resize(nmb); c->need_reenc.resize(nmb);
    c->out_words.resize((size_t)nmb * 160 + 1024);
    c->sadmap.assign((size_t)nmb * SM_WORDS, 0);
    c->me_field.assign((size_t)nmb * 16, 0);
    c->cur = 0; c->last_dec = 0;
    c->clusters[0] = c->clusters[1] = 0;
    c->have_traj = 0;
    memset(c->stats, 0, sizeof(c->stats));
    *out = c;
    return 0;
}
extern "C" void h264b200_ctx_destroy(h264b200_ctx *c) { delete c; }
extern "C" void h264b200_ctx_reset(h264b200_ctx *c) { c->clusters[0] = c->clusters[1] = 0; c->cur = 0; c->last_dec = 0; c->have_traj = 0; c->cost_stat_valid = 0; c->dn[0].clear(); c->dn[1].clear(); c->dn_cur = 0; }

static void run_job(h264b200_job *job)
{
    h264b200_ctx *c = job->ctx;
    FrameParams fp;
    memset(&fp, 0, sizeof(fp));
    const h264b200_frame_params &p = job->p;
    fp.width = c->width; fp.height = c->height; fp.nmbx = c->nmbx; fp.nmby = c->nmby;
    fp.slice_type = p.slice_type; fp.qp = p.qp; fp.speed = p.speed; fp.disable_deblock = p.disable_deblock;
    fp.lambda_q4 = p.lambda_q4; fp.lambda_mv_q4 = p.lambda_mv_q4; fp.lambda_i4_q4 = p.lambda_i4_q4;
    fp.lambda_i16_q4 = p.lambda_i16_q4; fp.skip_thr_inter = p.skip_thr_inter; fp.skip_thr_i4x4 = p.skip_thr_i4x4;
    fp.mvlim_x0 = -14 * 4; fp.mvlim_y0 = -14 * 4;
    fp.mvlim_x1 = (c->nmbx * 16 - 2) * 4; fp.mvlim_y1 = (c->nmby * 16 - 2) * 4;
    for (int i = 0; i < 2; i++)
    {
        fp.df_alpha[i] = p.df_alpha[i]; fp.df_beta[i] = p.df_beta[i];
        for (int k = 0; k < 4; k++) fp.df_tc0[i][k] = p.df_tc0[i][k];
    }
    memcpy(fp.qdat, p.qdat, sizeof(fp.qdat));
    for (int i = 0; i < 3; i++)
    {
        if (job->preloaded_index >= 0)
        {
            size_t fs = (size_t)c->width * c->height * 3 / 2, ys = (size_t)c->width * c->height;
            const pix_t *b = c->clip.data() + fs * job->preloaded_index;
            fp.inp[i] = i == 0 ? b : (i == 1 ? b + ys : b + ys + ys / 4);
            fp.inp_stride[i] = i ? c->width / 2 : c->width;
        } else { fp.inp[i] = job->yuv[i]; fp.inp_stride[i] = job->stride[i]; }
        if (p.denoise)
        {   /* filter the submitted picture, encode the filter's output (same flow as shim_cuda.cu) */
            const int w = i ? c->width >> 1 : c->width, h = i ? c->height >> 1 : c->height;
            const size_t s0 = (size_t)c->width * c->height, s1 = s0 / 4, off = i == 0 ? 0 : (i == 1 ? s0 : s0 + s1);
            for (int k = 0; k < 2; k++) if (c->dn[k].empty()) c->dn[k].assign(s0 + 2 * s1 + 16, 0);
            const pix_t *prev = c->dn[c->dn_cur].data() + off;
            pix_t *out = c->dn[c->dn_cur ^ 1].data() + off;
            if (w > 2 && h > 2)
                for (int y = 0; y < h; y++)
                    for (int x0 = 0; x0 < w; x0 += 4) denoise_word(fp.inp[i], fp.inp_stride[i], prev, out, w, w, h, x0, y);
            fp.inp[i] = out; fp.inp_stride[i] = w;
        }
        fp.dec[i] = c->frame[c->cur].data() + c->plane_off[i];
        fp.ref[i] = c->frame[c->cur ^ 1].data() + c->plane_off[i];
    }
    if (p.denoise == 2)
    {   /* transparent frame of a session with the noise suppressor: only the filter state advances (see shim_cuda.cu) */
        c->dn_cur ^= 1;
        job->status = 0; job->out_words = NULL; job->out_bits = 0;
        return;
    }
    for (int i = 0; i < 3; i++) fp.hp[i] = c->hpel.data() + i * c->luma_bytes + c->plane_off[0];
    fp.hp_out = c->hpel.data(); fp.dec_base = c->frame[c->cur].data(); fp.luma_bytes = (int)c->luma_bytes;
    fp.update_ref = job->update_ref;
    fp.stride[0] = c->stride[0]; fp.stride[1] = c->stride[1];
    fp.mbi = c->mbi.data(); fp.coef = c->coef.data();
    fp.clusters = c->clusters;
    if (c->cost_stat.empty()) c->cost_stat.assign(2 + c->nmby, 0);
    fp.cost_stat = c->cost_stat.data();
    fp.thr_eighths = getenv("H264B200_THR") ? atoi(getenv("H264B200_THR")) : 13;
    fp.have_cost_stat = c->cost_stat_valid;
    fp.spec = c->spec.data(); fp.cl_true = c->cl_true.data(); fp.cl_ckpt = c->cl_ckpt.data(); fp.changed_pass = c->changed_pass.data(); fp.need_reenc = c->need_reenc.data();
    memset(c->fsync, 0, sizeof(c->fsync));
    fp.fsync = c->fsync;
    fp.max_passes = 1000;
    fp.spec_from_prev = (p.slice_type == SLICE_P && c->have_traj && !getenv("H264B200_NO_PREV_TRAJ"));
    fp.mb_bits = c->mb_bits.data(); fp.mb_nbits = c->mb_nbits.data();
    fp.out_words = c->out_words.data();
    job->out_words = c->out_words.data();
    const int out_cap_words = (int)c->out_words.size();
    fp.hdr_bits = p.hdr_bits;
    const int nmb = c->nmbx * c->nmby;

    /* SAD-map pre-pass (h264_sadmap.h): before the sweep, while the record array still holds the previous frame */
    fp.sadmap = c->sadmap.data();
    fp.use_sadmap = p.slice_type == SLICE_P && !getenv("H264B200_NO_SADMAP");
    if (fp.use_sadmap)
        for (int y = 0; y < c->nmby; y++)
            for (int x = 0; x < c->nmbx; x++) sadmap_build_mb(&fp, x, y);

    MBWork *w = new MBWork();
    if (fp.spec_from_prev) { wave_replay(&fp, w, 1); c->cost_stat_valid = fp.have_cost_stat = 1; }
    /* speculative motion estimation ahead of the wavefront (h264_wave.h): round 0 + refinement rounds */
    fp.me_field = c->me_field.data();
    fp.use_me = fp.use_sadmap && !getenv("H264B200_NO_ME_PREPASS");
    fp.spec_no_intra = p.slice_type == SLICE_P && !getenv("H264B200_NO_INTRA_SPEC");
    if (fp.use_me)
        for (int round = 0; round < (getenv("H264B200_ME_ROUNDS") ? atoi(getenv("H264B200_ME_ROUNDS")) : ME_ROUNDS); round++)
            for (int y = 0; y < c->nmby; y++)
                for (int x = 0; x < c->nmbx; x++) me_prepass_mb(&fp, w, x, y, round);
    const int use_fast = fp.use_me && fp.spec_no_intra && !getenv("H264B200_NO_FAST");
    for (int pass = 0;;)
    {
        if (pass > 0)
            for (int r = 0; r < REPAIR_ROUNDS; r++)
                for (int y = 0; y < c->nmby; y++)
                    for (int x = 0; x < c->nmbx; x++)
                        if (fp.need_reenc[y * c->nmbx + x] == REPAIR_TAG(pass, r)) { wave_mb_round(&fp, w, x, y, pass, r); mb_store_coefs(&fp, w); }
        for (int y = 0; y < c->nmby; y++)
            for (int x = 0; x < c->nmbx; x++)
            {
                if (pass == 0)
                {
                    /* fast path of P frames (h264_fast.h): decide from the motion-estimation record, then the pixel work */
                    const int n = y * c->nmbx + x;
                    if (use_fast && fast_decide(&fp, w, x, y, fp.sadmap + (size_t)n * SM_WORDS + SM_ME_OFF, 0, fp.have_cost_stat ? fp.cost_stat[2 + y] : 0)) { fast_work(&fp, w, x, y, 0, 0); g_emu_lut[6]++; continue; }
                    if (use_fast && getenv("H264B200_DUMP_SLOW")) fprintf(stderr, "slow y %d x %d reason %d\n", y, x, g_emu_reason);
                    wave_mb_first(&fp, w, x, y);
                    for (int i = 0; i < 16; i++) w->last_mv[i] = fp.mbi[n].mv[i];
                }
                else wave_mb_repair(&fp, w, x, y, pass);
                mb_store_coefs(&fp, w);
            }
        int next = wave_end_of_pass(&fp, w, pass);
        if (next == FS_DONE) break;
        /* parallel re-check of the dirty macroblocks before the repair sweep */
        c->fsync[FS_STATE] = next;
        if (pass == 0 && fp.spec_no_intra)      /* sweep 0 left the intra modes out: verify them now, everywhere */
            for (int y = 0; y < c->nmby; y++)
                for (int x = 0; x < c->nmbx; x++) wave_mb_intra_check(&fp, w, x, y);
        for (int y = 0; y < c->nmby; y++)
            for (int x = 0; x < c->nmbx; x++) wave_mb_check(&fp, w, x, y, next);
        if (wave_after_check(&fp, next) == FS_DONE) break;
        pass = next;
        if (pass > fp.max_passes) { job->status = -4; delete w; return; }
    }
    delete w;
    if (getenv("H264B200_DUMP") && p.slice_type == SLICE_P)
    {   /* developer statistic: inter cost vs final type of every macroblock */
        FILE *f = fopen(getenv("H264B200_DUMP"), "a");
        for (int n = 0; n < nmb; n++) fprintf(f, "%d %d %d %d\n", n, c->spec[n].pad[0], (int)c->mbi[n].type, c->spec[n].flags);
        fprintf(f, "-1 0 0 0\n");
        fclose(f);
    }
    c->have_traj = (p.slice_type == SLICE_P);
    c->stats[0] += c->fsync[FS_PASSES]; c->stats[1] += c->fsync[FS_REENC]; c->stats[2] += c->fsync[FS_CHECKS]; c->stats[3]++;
    g_launches++;

    memset(c->out_words.data(), 0, c->out_words.size() * 4);
    int bo = p.hdr_bits;
    for (int n = 0; n <= nmb; n++)
    {
        int nb = cavlc_mb(&fp, n);
        if (bo + nb + 64 > out_cap_words * 32) { job->status = -2; return; }
        pack_mb(&fp, n, nb, bo);
        bo += nb;
    }
    {
        int run = 0;
        if (p.slice_type == SLICE_P) for (int k = nmb - 1; k >= 0 && fp.mbi[k].type == MBT_SKIP; k--) run++;
        job->trailing_skip_run = run;
    }
    job->out_bits = bo;
    g_launches++;

    if (!p.disable_deblock)
    {
        DeblockTile tile;
        for (int y = 0; y < c->nmby; y++)
            for (int x = 0; x < c->nmbx; x++) { for (int part = 0; part < 2; part++) { deblock_mb(&fp, &tile, x, y, part, 0); deblock_mb(&fp, &tile, x, y, part, 1); } }
    }
    for (int pl = 0; pl < 3; pl++)
    {
        long ns = border_samples(&fp, pl);
        for (long i = 0; i < ns; i++) extend_border_sample(&fp, pl, i);
    }
    if (job->update_ref)
        for (long i = 0; i < (long)(c->luma_bytes >> 2); i++) hpel_plane_word(&fp, i);
    g_launches++;
    for (int pl = 0; pl < 3; pl++)
        if (job->recon[pl])
        {
            int ww = c->nmbx * (pl ? 8 : 16), hh = c->nmby * (pl ? 8 : 16);
            for (int r = 0; r < hh; r++) memcpy(job->recon[pl] + (size_t)r * job->recon_stride[pl], fp.dec[pl] + (size_t)r * fp.stride[pl != 0], ww);
        }
    c->last_dec = c->cur;
    if (job->update_ref) c->cur ^= 1;
    if (p.denoise) c->dn_cur ^= 1;
    job->status = 0;
}

extern "C" int h264b200_encode_frames(int n, h264b200_job *jobs)
{
    int rc = 0;
    for (int i = 0; i < n; i++) { run_job(jobs + i); if (jobs[i].status && !rc) rc = jobs[i].status; }
    return rc;
}
extern "C" int h264b200_get_recon(h264b200_ctx *c, unsigned char *const planes[3], const int strides[3])
{
    for (int pl = 0; pl < 3; pl++)
    {
        int ww = c->nmbx * (pl ? 8 : 16), hh = c->nmby * (pl ? 8 : 16);
        const pix_t *src = c->frame[c->last_dec].data() + c->plane_off[pl];
        for (int r = 0; r < hh; r++) memcpy(planes[pl] + (size_t)r * strides[pl], src + (size_t)r * c->stride[pl != 0], ww);
    }
    return 0;
}
/* the emulation reads host memory directly: nothing to stage */
extern "C" int h264b200_prefetch_input(h264b200_ctx *, const unsigned char *const *, const int *) { return 0; }
extern "C" long h264b200_prefetch_hits(void) { return 0; }
extern "C" int h264b200_preload(h264b200_ctx *c, int nframes, const unsigned char *frames)
{
    size_t fs = (size_t)c->width * c->height * 3 / 2;
    c->clip.assign(frames, frames + fs * nframes);
    return 0;
}
extern "C" int h264b200_debug_get_sadmap(h264b200_ctx *c, unsigned int *out, int max_words)
{
    const int words = SM_WORDS * c->nmbx * c->nmby;
    if (max_words < words) return -3;
    memcpy(out, c->sadmap.data(), sizeof(uint32_t) * (size_t)words);
    return SM_WORDS;
}
extern "C" void h264b200_note_transparent(h264b200_ctx *c) { if (c) c->last_dec = c->cur ^ 1; }
extern "C" int h264b200_last_timing_ex(float *out_ms, int n) { for (int i = 0; i < n; i++) out_ms[i] = 0; return 8; }
extern "C" void h264b200_last_timing(float out_ms[4]) { out_ms[0] = out_ms[1] = out_ms[2] = out_ms[3] = 0; }
extern "C" void h264b200_ctx_stats(h264b200_ctx *c, int out[4]) { for (int i = 0; i < 4; i++) out[i] = c->stats[i]; }
extern "C" int h264b200_ctx_stats_ex(h264b200_ctx *c, int *out, int n) { for (int i = 0; i < n; i++) out[i] = i < 4 ? c->stats[i] : 0; return 8; }
extern "C" long h264b200_launch_count(void) { return g_launches; }
extern "C" const char *h264b200_backend_name(void) { return "host-emulation (test only)"; }
extern "C" void emu_get_cl_true(h264b200_ctx *c, int32_t *out) { memcpy(out, c->cl_true.data(), c->cl_true.size() * 4); }
