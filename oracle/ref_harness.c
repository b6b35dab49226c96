/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C harness around the UNMODIFIED reference encoder.  The reference is a
 * single header (`/root/reference/src/h264-lab.h`) whose implementation part is
 * always compiled (H:320), so this file simply #includes it from where it lies
 * (the Makefile passes -I/root/reference/src; nothing is copied into the repo)
 * and exports
 *   (1) the reference's public API under ref_* names            (H:264-312)
 *   (2) whole-sequence helpers used for golden vectors / CPU baseline
 *   (3) direct entry points to the reference's static L1 functions
 *        for known-answer tests of the
 *       restatement in oracle/h264_oracle.c and of the CUDA kernels.
 *
 * The output of the build (oracle/_ref/libh264ref.so) is git-ignored and
 * travels to the GPU box with the snapshot.
 */
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef REF_MB_HOOK_ENABLED
/* Only defined for the developer-only "hooked" build (see Makefile target
 * `hooked`): a sed-patched temporary copy of the header calls this once per
 * macroblock right after mb_encode().  */
struct H264E_persist_tag;
static void ref_mb_hook(struct H264E_persist_tag *enc);
#define REF_MB_HOOK(enc) ref_mb_hook(enc)
#endif

#ifdef REF_OPCOUNT_ENABLED
/* Only defined for the measurement build (Makefile target _ref/libh264ref_count.so): opcount.sed adds one
 * REF_CNT(category, n) statement at the top of the reference's leaf functions in a temporary copy of the header. */
static long long g_ref_ops[16];
#define REF_CNT(k, n) (g_ref_ops[k] += (n))
#endif

#include "h264-lab.h"

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* (1) public API                                                      */
/* ------------------------------------------------------------------ */
EXPORT int ref_sizeof(const H264E_create_param_t *p, int *sp, int *ss) { return H264E_sizeof(p, sp, ss); }
EXPORT int ref_init(void *enc, const H264E_create_param_t *p) { return H264E_init((H264E_persist_t *)enc, p); }
EXPORT int ref_encode(void *enc, void *scratch, const H264E_run_param_t *rp, H264E_io_yuv_t *yuv,
                      unsigned char **coded, int *ncoded)
{
    return H264E_encode((H264E_persist_t *)enc, (H264E_scratch_t *)scratch, rp, yuv, coded, ncoded);
}
EXPORT void ref_set_vbv_state(void *enc, int a, int b) { H264E_set_vbv_state((H264E_persist_t *)enc, a, b); }
EXPORT int ref_sizeof_enc_struct(void) { return (int)sizeof(h264e_enc_t); }
EXPORT int ref_sizeof_scratch_struct(void) { return (int)sizeof(scratch_t); }

/* Reconstruction of the frame just encoded.  After H264E_encode the
 * reconstructed picture has been moved to enc->ref (H:3580-3596). Copies the
 * nmbx*16 x nmby*16 luma and the two chroma planes, tightly packed. */
EXPORT void ref_get_recon(void *venc, unsigned char *y, unsigned char *u, unsigned char *v)
{
    h264e_enc_t *enc = (h264e_enc_t *)venc;
    int w = enc->frame.w, h = enc->frame.h, r;
    for (r = 0; r < h; r++) memcpy(y + r * w, enc->ref.yuv[0] + r * enc->ref.stride[0], w);
    for (r = 0; r < h / 2; r++) memcpy(u + r * (w / 2), enc->ref.yuv[1] + r * enc->ref.stride[1], w / 2);
    for (r = 0; r < h / 2; r++) memcpy(v + r * (w / 2), enc->ref.yuv[2] + r * enc->ref.stride[2], w / 2);
}

/* Reconstruction including the replicated 16/8-pixel borders (H:2232). */
EXPORT void ref_get_recon_padded(void *venc, unsigned char *y, unsigned char *u, unsigned char *v)
{
    h264e_enc_t *enc = (h264e_enc_t *)venc;
    int w = enc->frame.w + 32, h = enc->frame.h + 32, r;
    for (r = 0; r < h; r++) memcpy(y + r * w, enc->ref.yuv[0] + (r - 16) * enc->ref.stride[0] - 16, w);
    w >>= 1; h >>= 1;
    for (r = 0; r < h; r++) memcpy(u + r * w, enc->ref.yuv[1] + (r - 8) * enc->ref.stride[1] - 8, w);
    for (r = 0; r < h; r++) memcpy(v + r * w, enc->ref.yuv[2] + (r - 8) * enc->ref.stride[2] - 8, w);
}

EXPORT void ref_get_state(void *venc, int *out /* [8] */)
{
    h264e_enc_t *enc = (h264e_enc_t *)venc;
    out[0] = enc->rc.qp;
    out[1] = enc->mv_clusters[0].u32;
    out[2] = enc->mv_clusters[1].u32;
    out[3] = enc->frame.num;
    out[4] = enc->rc.vbv_bits;
    out[5] = enc->rc.dqp_smooth;
    out[6] = enc->rc.prev_qp;
    out[7] = enc->next_idr_pic_id;
}

#ifdef REF_OPCOUNT_ENABLED
EXPORT void ref_reset_opcounts(void) { memset(g_ref_ops, 0, sizeof(g_ref_ops)); }
EXPORT void ref_get_opcounts(long long *out /* [16] */) { memcpy(out, g_ref_ops, sizeof(g_ref_ops)); }
#endif

/* ------------------------------------------------------------------ */
/* (2) whole-sequence helper                                           */
/* ------------------------------------------------------------------ */
/*
 * Encode `nframes` I420 frames (tightly packed, stride == width) exactly the
 * way the reference CLI does (T:507-526, T:588-604): fixed QP when kbps == 0,
 * otherwise desired_frame_bytes = kbps*1000/8/30 with QP 10..50.
 * One fresh encoder instance per call (== one closed-GOP segment / stream).
 *
 *  out        : concatenated access units
 *  out_sizes  : bytes per frame
 *  recon      : optional, nframes * (W16*H16*3/2) reconstructed frames
 *  returns total bytes or -(error code)
 *  *seconds   : wall time of the H264E_encode loop only (CLOCK_MONOTONIC)
 */
static long ref_encode_sequence_impl(int width, int height, int gop, int qp, int kbps, int speed, int flags,
                                int nframes, const unsigned char *yuv_in,
                                unsigned char *out, long out_cap, int *out_sizes,
                                unsigned char *recon, double *seconds);
EXPORT long ref_encode_sequence(int width, int height, int gop, int qp, int kbps, int speed,
                                int nframes, const unsigned char *yuv_in,
                                unsigned char *out, long out_cap, int *out_sizes,
                                unsigned char *recon, double *seconds)
{
    return ref_encode_sequence_impl(width, height, gop, qp, kbps, speed, 0, nframes, yuv_in, out, out_cap, out_sizes, recon, seconds);
}
/* same with create-time options: flags bit 0 = temporal_denoise_flag (H:122), bit 1 = vbv_overflow_empty_frame_flag
 * (H:96), bit 2 = vbv_underflow_stuffing_flag (H:101) */
EXPORT long ref_encode_sequence_ex(int width, int height, int gop, int qp, int kbps, int speed, int flags,
                                   int nframes, const unsigned char *yuv_in,
                                   unsigned char *out, long out_cap, int *out_sizes,
                                   unsigned char *recon, double *seconds)
{
    return ref_encode_sequence_impl(width, height, gop, qp, kbps, speed, flags, nframes, yuv_in, out, out_cap, out_sizes, recon, seconds);
}
/* the reference's temporal noise suppressor on caller buffers (known-answer tests) */
EXPORT void ref_denoise_run(unsigned char *frm, unsigned char *frmprev, int w, int h, int stride_frm, int stride_frmprev)
{
    h264e_denoise_run(frm, frmprev, w, h, stride_frm, stride_frmprev);
}
static long ref_encode_sequence_impl(int width, int height, int gop, int qp, int kbps, int speed, int flags,
                                int nframes, const unsigned char *yuv_in,
                                unsigned char *out, long out_cap, int *out_sizes,
                                unsigned char *recon, double *seconds)
{
    H264E_create_param_t cp;
    H264E_run_param_t rp;
    H264E_io_yuv_t yuv;
    int sp = 0, ss = 0, err, i;
    long pos = 0;
    size_t frame_size = (size_t)width * height * 3 / 2;
    unsigned char *frame_copy;
    void *enc, *scratch;
    struct timespec t0, t1;
    double acc = 0;

    memset(&cp, 0, sizeof(cp));
    memset(&rp, 0, sizeof(rp));
    cp.enableNEON = 1;
    cp.num_layers = 1;
    cp.gop = gop;
    cp.width = width;
    cp.height = height;
    cp.const_input_flag = 1;
    cp.vbv_size_bytes = 100000 / 8;
    cp.temporal_denoise_flag = flags & 1;
    cp.vbv_overflow_empty_frame_flag = (flags >> 1) & 1;
    cp.vbv_underflow_stuffing_flag = (flags >> 2) & 1;
    err = H264E_sizeof(&cp, &sp, &ss);
    if (err) return -err;
    enc = aligned_alloc(64, ((size_t)sp + 63) & ~(size_t)63);
    scratch = aligned_alloc(64, ((size_t)ss + 63) & ~(size_t)63);
    frame_copy = (unsigned char *)malloc(frame_size);
    H264E_init((H264E_persist_t *)enc, &cp);

    for (i = 0; i < nframes; i++)
    {
        unsigned char *coded = NULL;
        int ncoded = 0;
        memcpy(frame_copy, yuv_in + (size_t)i * frame_size, frame_size);
        yuv.yuv[0] = frame_copy;                          yuv.stride[0] = width;
        yuv.yuv[1] = frame_copy + width * height;         yuv.stride[1] = width / 2;
        yuv.yuv[2] = frame_copy + width * height * 5 / 4; yuv.stride[2] = width / 2;
        rp.frame_type = 0;
        rp.encode_speed = speed;
        if (kbps)
        {
            rp.desired_frame_bytes = kbps * 1000 / 8 / 30;
            rp.qp_min = 10;
            rp.qp_max = 50;
        } else
        {
            rp.qp_min = rp.qp_max = qp;
        }
        clock_gettime(CLOCK_MONOTONIC, &t0);
        err = H264E_encode((H264E_persist_t *)enc, (H264E_scratch_t *)scratch, &rp, &yuv, &coded, &ncoded);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        acc += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        if (err) { pos = -err; break; }
        if (pos + ncoded > out_cap) { pos = -100; break; }
        memcpy(out + pos, coded, ncoded);
        pos += ncoded;
        if (out_sizes) out_sizes[i] = ncoded;
        if (recon)
        {
            h264e_enc_t *e = (h264e_enc_t *)enc;
            size_t rs = (size_t)e->frame.w * e->frame.h;
            unsigned char *r = recon + (size_t)i * rs * 3 / 2;
            ref_get_recon(enc, r, r + rs, r + rs + rs / 4);
        }
    }
    if (seconds) *seconds = acc;
    free(frame_copy);
    free(enc);
    free(scratch);
    return pos;
}

/* ------------------------------------------------------------------ */
/* (3) L1 functions for known-answer tests               */
/* ------------------------------------------------------------------ */
EXPORT int ref_sad_block(const unsigned char *a, int as, const unsigned char *b, int bs, int w, int h)
{ return sad_block(a, as, b, bs, w, h); }                                   /* a1  H:2162 */
EXPORT int ref_sad_mb_8x8(const unsigned char *a, int as, const unsigned char *b, int *sad4)
{ return h264e_sad_mb_unlaign_8x8(a, as, b, sad4); }                        /* a1  H:2178 */
EXPORT void ref_qpel_luma(const unsigned char *src, int stride, unsigned char *dst, int w, int h, int dx, int dy)
{ h264e_qpel_interpolate_luma(src, stride, dst, point(w, h), point(dx, dy)); } /* a2 H:2079 */
EXPORT void ref_qpel_chroma(const unsigned char *src, int stride, unsigned char *dst, int w, int h, int dx, int dy)
{ h264e_qpel_interpolate_chroma(src, stride, dst, point(w, h), point(dx, dy)); } /* a3 H:2133 */
EXPORT void ref_qpel_average(const unsigned char *s0, const unsigned char *s1, unsigned char *dst, int w, int h)
{ h264e_qpel_average_wh_align(s0, s1, dst, point(w, h)); }                   /* a2 H:2065 */

/* left/top may be NULL (unavailable) exactly like mb_encode passes them (H:5742-5743) */
EXPORT void ref_intra16(unsigned char *pred, const unsigned char *left, const unsigned char *top, int mode)
{ h264e_intra_predict_16x16(pred, left, top, mode); }                        /* a7 H:1677 */
/* left32/top32 use the top_line layout (16 Y, 8 U, 8 V); the reference passes left+16 / top+16
 * even when the pointer is NULL and relies on IS_NULL(p) = p < 32 (H:1623, H:5784). */
EXPORT void ref_intra_chroma(unsigned char *pred, const unsigned char *left32, const unsigned char *top32, int mode)
{
    const pix_t *l = (const pix_t *)((uintptr_t)left32 + 16);
    const pix_t *t = (const pix_t *)((uintptr_t)top32 + 16);
    h264e_intra_predict_chroma(pred, l, t, mode);                            /* a7 H:1716 */
}
EXPORT int ref_intra16_estimate(unsigned char *p, int s, int avail, int qp)
{ return intra_estimate_16x16(p, s, avail, qp); }                            /* a7 H:4838 */
/* edge: 32-byte buffer, edge pointer = buf + 16 (needs [-5..7]) */
EXPORT int ref_intra4_choose(const unsigned char *blockin, unsigned char *blockpred, int avail,
                             unsigned char *edge, int mpred, int penalty)
{ return h264e_intra_choose_4x4(blockin, blockpred, avail, edge, mpred, penalty); } /* a8 H:1810 */

/* qdat for a given qp / slice type, as built by rc_set_qp (H:5839). out: 2*42 u16 */
EXPORT void ref_make_qdat(int qp, int is_p_slice, unsigned short *out)
{
    static h264e_enc_t e;
    memset(&e, 0, sizeof(e));
    e.run_param.qp_min = 10;
    e.run_param.qp_max = 51;
    e.slice.type = is_p_slice ? SLICE_TYPE_P : SLICE_TYPE_I;
    e.rc.qp = 0;
    rc_set_qp(&e, qp);
    memcpy(out, e.rc.qdat, sizeof(e.rc.qdat));
}

/* a9/a10/a11: transform+quant (+DC) of one component exactly as mb_write drives it.
 * mode: 2 = INTRA_4 (1 block), 8 = INTER, 9 = INTRA_16, 5 = CHROMA.
 * q_out: n*n quant_t (qv[16], dq[16]); dc_out: 16 quantised DC levels (I16 / chroma only).
 * returns nz mask. The DC step runs when mode is odd. */
EXPORT int ref_transform_quant(const unsigned char *inp, const unsigned char *pred, int inp_stride, int mode,
                               short *q_out, short *dc_out, const unsigned short *qdat)
{
    struct { int16_t dc[16]; quant_t q[16]; } s;
    int n = mode >> 1, nz;
    memset(&s, 0, sizeof(s));
    nz = h264e_transform_sub_quant_dequant(inp, pred, inp_stride, mode, s.q, qdat);
    if (mode == QDQ_MODE_INTRA_16) h264e_quant_luma_dc(s.q, dc_out, qdat);
    if (mode == QDQ_MODE_CHROMA) nz |= h264e_quant_chroma_dc(s.q, dc_out, qdat) << 8;
    memcpy(q_out, s.q, n * n * sizeof(quant_t));
    return nz;
}
EXPORT void ref_transform_add(unsigned char *out, int out_stride, const unsigned char *pred, short *q, int side, int mask)
{ h264e_transform_add(out, out_stride, pred, (quant_t *)q, side, mask); }     /* a11 H:2638 */

/* a13: CAVLC residual block. Returns number of bits; bytes written big-endian to out. */
EXPORT int ref_vlc_encode(const short *qv16, int maxNumCoeff, int nA, int nB, unsigned char *out, int *nnz_out)
{
    static uint32_t buf[64];
    int16_t q[16];
    uint8_t ctx[3];
    bs_t bs;
    int bits;
    memset(buf, 0, sizeof(buf));
    memcpy(q, qv16, sizeof(q));
    ctx[0] = (uint8_t)nA; ctx[1] = 0; ctx[2] = (uint8_t)nB;
    h264e_bs_init_bits(&bs, buf);
    h264e_vlc_encode(&bs, q, maxNumCoeff, ctx + 1);
    bits = h264e_bs_get_pos_bits(&bs);
    h264e_bs_flush(&bs);
    memcpy(out, buf, (bits + 7) / 8 + 4);
    if (nnz_out) *nnz_out = ctx[1];
    return bits;
}

/* a15: one MB of deblocking with explicit parameters (H:1505, H:1469) */
EXPORT void ref_deblock_luma(unsigned char *pix, int stride, const unsigned char *strength32,
                             const unsigned char *tc0, const unsigned char *alpha, const unsigned char *beta)
{
    deblock_params_t par;
    memcpy(par.strength32, strength32, 32);
    memcpy(par.tc0, tc0, 32);
    memcpy(par.alpha, alpha, 4);
    memcpy(par.beta, beta, 4);
    h264e_deblock_luma(pix, stride, &par);
}
EXPORT void ref_deblock_chroma(unsigned char *pix, int stride, const unsigned char *strength32,
                               const unsigned char *tc0, const unsigned char *alpha, const unsigned char *beta)
{
    deblock_params_t par;
    memcpy(par.strength32, strength32, 32);
    memcpy(par.tc0, tc0, 32);
    memcpy(par.alpha, alpha, 4);
    memcpy(par.beta, beta, 4);
    h264e_deblock_chroma(pix, stride, &par);
}
EXPORT void ref_copy_borders(unsigned char *pic, int w, int h, int guard) { h264e_copy_borders(pic, w, h, guard); } /* a16 */

/* host-side helpers, for testing the host mirror */
EXPORT int ref_nal_escape(unsigned char *d, const unsigned char *s, int n) { return nal_put_esc(d, s, n); }
EXPORT unsigned ref_div_q16(unsigned a, unsigned b) { return div_q16(a, b); }
EXPORT unsigned ref_mul32x32shr16(unsigned a, unsigned b) { return mul32x32shr16(a, b); }

#ifdef REF_MB_HOOK_ENABLED
/* ------------------------------------------------------------------ */
/* developer-only per-macroblock decision dump                         */
/* ------------------------------------------------------------------ */
typedef struct
{
    int type, i16_mode, cost, nz_mask, bitpos, skip_run, cl0, cl1;
    int mv[16];
    int mvd[16];
    signed char i4[16];
} ref_mb_dump_t;
static ref_mb_dump_t *g_dump;
static int g_dump_cap, g_dump_n;
EXPORT void ref_set_mb_dump(void *buf, int cap) { g_dump = (ref_mb_dump_t *)buf; g_dump_cap = cap; g_dump_n = 0; }
EXPORT int ref_get_mb_dump_count(void) { return g_dump_n; }
static void ref_mb_hook(struct H264E_persist_tag *enc)
{
    int i;
    ref_mb_dump_t *d;
    if (!g_dump || g_dump_n >= g_dump_cap) return;
    d = g_dump + g_dump_n++;
    d->type = enc->mb.type;
    d->i16_mode = enc->mb.i16.pred_mode_luma;
    d->cost = enc->mb.cost;
    d->nz_mask = enc->scratch->nz_mask;
    d->bitpos = h264e_bs_get_pos_bits(enc->bs);
    d->skip_run = enc->mb.skip_run;
    d->cl0 = enc->mv_clusters[0].u32;
    d->cl1 = enc->mv_clusters[1].u32;
    for (i = 0; i < 16; i++)
    {
        d->mv[i] = enc->mb.mv[i].u32;
        d->mvd[i] = enc->mb.mvd[i].u32;
        d->i4[i] = enc->mb.i4x4_mode[i];
    }
}
#endif
