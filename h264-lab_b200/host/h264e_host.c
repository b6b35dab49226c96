/*
 * h264e_host.c -- host side of the B200 encoder: the reference's public API
 * (include/h264-lab.h) with frame-type / GOP logic, rate control, parameter sets,
 * slice headers and NAL assembly in plain C.  Every macroblock-level operation is
 * submitted to the device through the C-ABI shim (include/h264b200_shim.h); there is
 * no CPU path for macroblock work.
 *
 * Behaviour follows /root/reference/src/h264-lab.h (cited as H:nnn) so that the byte
 * stream is identical; the code is organised around an explicit per-frame "plan"
 * (frame type -> slice parameters -> device job -> NAL assembly) instead of the
 * reference's in-place encoder object.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "h264-lab.h"
#include "h264b200_shim.h"

/* ------------------------------------------------------------------------------ */
/* tables                                                                          */
/* ------------------------------------------------------------------------------ */
typedef struct
{
    uint16_t rnd_inter, deadzone_i, thr_inter, thr_inter2;
    uint16_t skip_thr_inter, skip_thr_i4x4, lambda_q4, lambda_mv_q4, lambda_i4_q4, lambda_i16_q4;
} h264e_qp_tunables_t;
#include "h264e_tables_gen.h"

/* quantiser multiplier / dequantiser scale per QP%6 for the three coefficient position
 * classes (ITU-T H.264 8.5.9 / the usual MF table): {MF, V} x {(0,0), mixed, (1,1)} */
static const int16_t k_quant_mf_v[6][6] = {
    {13107, 10, 8066, 13, 5243, 16},
    {11916, 11, 7490, 14, 4660, 18},
    {10082, 13, 6554, 16, 4194, 20},
    { 9362, 14, 5825, 18, 3647, 23},
    { 8192, 16, 5243, 20, 3355, 25},
    { 7282, 18, 4559, 23, 2893, 29},
};

/* level limits (ITU-T H.264 Table A-1): level_idc, MaxFS, MaxCPB/5 (kbit), MaxDPB (MBs) */
typedef struct { uint8_t level; uint16_t max_fs; uint16_t max_cpb_div5; uint32_t max_dpb; } level_limit_t;
static const level_limit_t k_levels[] = {
    {10, 99, 175 / 5, 396},     {10, 99, 350 / 5, 396},       {11, 396, 500 / 5, 900},
    {12, 396, 1000 / 5, 2376},  {13, 396, 2000 / 5, 2376},    {20, 396, 2000 / 5, 2376},
    {21, 792, 4000 / 5, 4752},  {22, 1620, 4000 / 5, 8100},   {30, 1620, 10000 / 5, 8100},
    {31, 3600, 14000 / 5, 18000}, {32, 5120, 20000 / 5, 20480}, {40, 8192, 25000 / 5, 32768},
    {41, 8192, 62500 / 5, 32768}, {42, 8704, 62500 / 5, 34816}, {50, 22080, 135000 / 5, 110400},
    {51, 36864, 240000 / 5, 184320},
};

#define MIN_QP 10
#define SLICE_P 0
#define SLICE_I 2
#define REF_SIZEOF_ENC 1184      /* sizeof(h264e_enc_t) of the reference's default build (probe) */
#define REF_SIZEOF_SCRATCH_T 2962

#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#define IMAX(a, b) ((a) > (b) ? (a) : (b))

/* ------------------------------------------------------------------------------ */
/* encoder object stored at the start of the caller's persist blob                 */
/* ------------------------------------------------------------------------------ */
#define H264E_MAGIC 0x42323030u

typedef struct h264e_host_tag
{
    uint32_t magic;
    struct h264e_host_tag *self;
    H264E_create_param_t param;
    H264E_run_param_t run;
    h264b200_ctx *ctx;
    int nmbx, nmby, nmb, w16, h16, cropping;
    int frame_num;
    int pic_init_qp;
    int next_idr_pic_id;
    int most_recent_ref_frame_idx;
    int disable_deblock;
    struct
    {
        int qp, vbv_bits, qp_smooth, dqp_smooth, max_dqp, bit_budget, prev_qp, vbv_target_level;
    } rc;
    /* per-call state */
    uint8_t *out;
    unsigned out_pos;
    int out_cap;
    struct h264e_host_tag *next_live;
} h264e_host_t;

static pthread_mutex_t g_live_lock = PTHREAD_MUTEX_INITIALIZER;
#include <time.h>
static double g_host_ms[3], g_host_t_finish0;       /* developer statistic, see H264E_b200_host_timing (not thread-exact) */
static double host_now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static h264e_host_t *g_live = NULL;
static int g_atexit_registered = 0;

static void live_remove_locked(h264e_host_t *e)
{
    h264e_host_t **pp = &g_live;
    while (*pp)
    {
        if (*pp == e) { *pp = e->next_live; return; }
        pp = &(*pp)->next_live;
    }
}
static int live_contains_locked(h264e_host_t *e)
{
    h264e_host_t *p;
    for (p = g_live; p; p = p->next_live) if (p == e) return 1;
    return 0;
}
static void live_cleanup_at_exit(void)
{
    /* the driver reclaims device memory at process exit; only unlink here */
    pthread_mutex_lock(&g_live_lock);
    g_live = NULL;
    pthread_mutex_unlock(&g_live_lock);
}

/* ------------------------------------------------------------------------------ */
/* host bit writer (parameter sets and slice headers only)                          */
/* ------------------------------------------------------------------------------ */
typedef struct { uint8_t buf[256]; uint64_t acc; int nacc; int nbytes; } hbits_t;
static void hb_init(hbits_t *b) { memset(b, 0, sizeof(*b)); }
static void hb_put(hbits_t *b, int n, uint32_t v)
{
    b->acc = (b->acc << n) | v;
    b->nacc += n;
    while (b->nacc >= 8) { b->nacc -= 8; b->buf[b->nbytes++] = (uint8_t)(b->acc >> b->nacc); }
}
static void hb_ue(hbits_t *b, uint32_t v)
{
    uint32_t t = v + 1, u;
    int size = 0;
    for (u = t; u; u >>= 1) size++;
    hb_put(b, 2 * size - 1, t);
}
static void hb_se(hbits_t *b, int v) { v = 2 * v - 1; v ^= v >> 31; hb_ue(b, (uint32_t)v); }
static int hb_bits(const hbits_t *b) { return b->nbytes * 8 + b->nacc; }
static void hb_trailing(hbits_t *b) { hb_put(b, 1, 1); if (b->nacc) hb_put(b, 8 - b->nacc, 0); }

/* Append one NAL unit to the access unit: 4-byte start code, then the payload with
 * emulation-prevention bytes (nal_start/nal_end/nal_put_esc H:3952-4022). */
static void emit_nal(h264e_host_t *e, const uint8_t *rbsp, int n)
{
    uint8_t *d = e->out + e->out_pos, *nal;
    int i, zeros = 0, j = 0;
    d[0] = d[1] = d[2] = 0; d[3] = 1;
    nal = d + 4;
    for (i = 0; i < n; i++)
    {
        uint8_t byte = rbsp[i];
        if (zeros == 2 && byte <= 3) { nal[j++] = 3; zeros = 0; }
        zeros = byte ? 0 : zeros + 1;
        nal[j++] = byte;
    }
    if (e->run.nalu_callback) e->run.nalu_callback(nal, j, e->run.nalu_callback_token);
    e->out_pos += 4 + j;
}

/* ------------------------------------------------------------------------------ */
/* parameter sets (encode_sps H:4040, encode_pps H:4147; AVC single layer only)     */
/* ------------------------------------------------------------------------------ */
static void write_sps(h264e_host_t *e)
{
    hbits_t b;
    const level_limit_t *lim = k_levels;
    hb_init(&b);
    while (lim->level < 51 && (e->nmb > lim->max_fs ||
                               e->param.vbv_size_bytes > lim->max_cpb_div5 * (5 * 1000 / 8) ||
                               (unsigned)(e->nmb * (e->param.max_long_term_reference_frames + 1)) > lim->max_dpb))
        lim++;
    hb_put(&b, 8, 0x67);
    hb_put(&b, 8, 66);                  /* Baseline */
    hb_put(&b, 8, 0);                   /* constraint flags */
    hb_put(&b, 8, lim->level);
    hb_ue(&b, (uint32_t)e->param.sps_id);
    hb_ue(&b, 1);                       /* log2_max_frame_num_minus4 */
    hb_ue(&b, 2);                       /* pic_order_cnt_type */
    hb_ue(&b, (uint32_t)(1 + e->param.max_long_term_reference_frames));
    hb_put(&b, 1, 0);
    hb_ue(&b, (uint32_t)(e->nmbx - 1));
    hb_ue(&b, (uint32_t)(e->nmby - 1));
    hb_put(&b, 3, (uint32_t)(6 + e->cropping));   /* frame_mbs_only, direct_8x8_inference, cropping */
    if (e->cropping)
    {
        hb_ue(&b, 0);
        hb_ue(&b, (uint32_t)((e->w16 - e->param.width) >> 1));
        hb_ue(&b, 0);
        hb_ue(&b, (uint32_t)((e->h16 - e->param.height) >> 1));
    }
    hb_put(&b, 1, 0);                   /* vui_parameters_present_flag */
    hb_trailing(&b);
    emit_nal(e, b.buf, b.nbytes);
}

static void write_pps(h264e_host_t *e)
{
    hbits_t b;
    hb_init(&b);
    hb_put(&b, 8, 0x68);
    hb_ue(&b, (uint32_t)(e->param.sps_id * 4));
    hb_ue(&b, (uint32_t)e->param.sps_id);
    hb_put(&b, 1, 0);                   /* entropy_coding_mode_flag: CAVLC */
    hb_put(&b, 1, 0);
    hb_ue(&b, 0);
    hb_ue(&b, 0);
    hb_ue(&b, 0);
    hb_put(&b, 1, 0);
    hb_put(&b, 2, 0);
    hb_se(&b, e->pic_init_qp - 26);
    hb_put(&b, 5, 0x1C);                /* pic_init_qs=0, chroma_qp_offset=0, deblocking_filter_control=1, 0, 0 */
    hb_trailing(&b);
    emit_nal(e, b.buf, b.nbytes);
}

/* NAL header byte + slice_header() (encode_slice_header H:4182-4333). No trailing bits. */
static void write_slice_header(h264e_host_t *e, hbits_t *b, int slice_type, int is_key, int long_term_idx_update)
{
    hb_init(b);
    hb_put(b, 8, (uint32_t)((is_key ? 5 : 1) | (long_term_idx_update >= 0 ? 0x60 : 0)));
    hb_ue(b, 0);                                        /* first_mb_in_slice */
    hb_ue(b, (uint32_t)slice_type);
    hb_ue(b, (uint32_t)(e->param.sps_id * 4));          /* pic_parameter_set_id */
    hb_put(b, 5, (uint32_t)(e->frame_num & 31));        /* frame_num, log2_max_frame_num = 5 */
    if (is_key) hb_ue(b, (uint32_t)e->next_idr_pic_id);
    if (slice_type == SLICE_P) hb_put(b, 2, 0);         /* no override, no list modification */
    if (long_term_idx_update >= 0)
    {
        if (is_key) hb_put(b, 2, e->param.max_long_term_reference_frames > 0);
        else hb_put(b, 1, 0);                           /* adaptive_ref_pic_marking_mode_flag */
    }
    hb_se(b, e->rc.prev_qp - e->pic_init_qp);           /* slice_qp_delta */
    hb_ue(b, (uint32_t)e->disable_deblock);
    if (e->disable_deblock != 1) hb_put(b, 2, 3);       /* alpha / beta offsets = 0 */
}

/* ------------------------------------------------------------------------------ */
/* rate control (H:5815-6141)                                                       */
/*                                                                                  */
/* This section (mul_q16 / div_q16, rc_frame_start, rc_frame_end) is a RESTATEMENT of */
/* the reference's frame-level rate control, H:3420-3440 and H:5924-6141, statement  */
/* by statement: the Q16 fixed-point arithmetic with its truncations, the budget /   */
/* VBV bookkeeping and the QP update ARE the contract (the QP trajectory must match  */
/* the reference frame for frame), so there is no room to */
/* write it differently.  Only the golden / long-term-frame branches are absent      */
/* (unsupported features).  It runs on the host, outside the device hot path.        */
/* ------------------------------------------------------------------------------ */
static uint32_t mul_q16(uint32_t x, uint32_t y)        /* H:3420 */
{
    return (x >> 16) * (y & 0xFFFFu) + x * (y >> 16) + ((y & 0xFFFFu) * (x & 0xFFFFu) >> 16);
}
static uint32_t div_q16(uint32_t numer, uint32_t denom)      /* H:3430 */
{
    uint32_t f = 1u << __builtin_clz(denom);
    do
    {
        denom = denom * f >> 16;
        numer = mul_q16(numer, f);
        f = ((1u << 17) - denom);
    } while (denom != 0xffff);
    return numer;
}

/* largest t with t*q <= 0x10000 - round (rc_rnd2thr H:5822) */
static uint16_t zero_threshold(int round, int q)
{
    int b, thr = 0;
    for (b = 0x8000; b; b >>= 1)
    {
        int t = (thr | b) * q;
        if (t <= 0x10000 - round) thr |= b;
    }
    return (uint16_t)thr;
}

/* Quantiser tables for luma and chroma at `qp` (rc_set_qp H:5839-5912).  Layout of each
 * 42-entry table: [0..5] {MF', V'} x 3 classes, [6] rounding used, [7] intra dead zone,
 * [8..9] raw thresholds, [10..17] thr1 per position, [18..25] thr2, [26..33] MF per
 * position, [34..41] V per position. */
static void build_qdat(uint16_t qdat[2][42], int qp, int slice_type)
{
    int c;
    for (c = 0; c < 2; c++)
    {
        uint16_t *q = qdat[c];
        int div6 = qp * 86 >> 9, mod6 = qp - div6 * 6, i;
        const h264e_qp_tunables_t *t = &h264e_qp_tunables[qp];
        static const uint8_t cls[8] = {0, 1, 0, 1, 1, 2, 1, 2};   /* position class of coefficient i&7 */
        uint16_t thr1[3], thr2[3];
        for (i = 0; i < 3; i++)
        {
            q[2 * i] = (uint16_t)(k_quant_mf_v[mod6][2 * i] << 1 >> div6);
            q[2 * i + 1] = (uint16_t)(k_quant_mf_v[mod6][2 * i + 1] << div6);
        }
        q[6] = slice_type == SLICE_P ? t->rnd_inter : t->deadzone_i;
        q[7] = t->deadzone_i;
        q[8] = (uint16_t)(t->thr_inter - 0x7fff);
        q[9] = (uint16_t)(t->thr_inter2 - 0x7fff);
        for (i = 0; i < 3; i++)
        {
            thr1[i] = zero_threshold(t->thr_inter - 0x7fff, q[2 * i]);
            thr2[i] = zero_threshold(t->thr_inter2 - 0x7fff, q[2 * i]);
        }
        for (i = 0; i < 8; i++)
        {
            q[10 + i] = thr1[cls[i]];
            q[18 + i] = thr2[cls[i]];
            q[26 + i] = q[2 * cls[i]];
            q[34 + i] = q[2 * cls[i] + 1];
        }
        qp = h264e_qp_chroma[qp];
    }
}

static int clamp_qp(const h264e_host_t *e, int qp)
{
    qp = IMIN(qp, e->run.qp_max);
    qp = IMAX(qp, e->run.qp_min);
    return IMIN(qp, 51);
}

/* frame bit budget and QP (rc_frame_start H:5924-6070, no long-term branch) */
static void rc_frame_start(h264e_host_t *e, int is_intra)
{
    unsigned np = IMIN((unsigned)e->param.gop - 1u, 63u);
    int nmb = e->nmb;
    int qp = -1, add_bits, bit_budget = e->run.desired_frame_bytes * 8;
    int nominal_p, gop_bits, stationary;
    uint32_t peak_q16;

    do
    {
        qp++;
        gop_bits = h264e_bits_per_mb[0][qp] * np + h264e_bits_per_mb[1][qp];
    } while (gop_bits * nmb > (int)(np + 1) * e->run.desired_frame_bytes * 8 && qp < 40);

    peak_q16 = div_q16((uint32_t)h264e_bits_per_mb[1][qp] << 16, (uint32_t)h264e_bits_per_mb[0][qp] << 16);
    if (np)
    {
        uint32_t ratio = div_q16((np + 1) << 16, (np << 16) + peak_q16);
        nominal_p = (int)mul_q16((uint32_t)(e->run.desired_frame_bytes * 8), ratio);
    } else nominal_p = 0;

    stationary = IMIN(e->param.vbv_size_bytes * 8 >> 4, e->run.desired_frame_bytes * 8);

    if (is_intra) add_bits = (int)mul_q16((uint32_t)nominal_p, peak_q16) - bit_budget;
    else
    {
        add_bits = nominal_p - bit_budget;
        if (e->param.vbv_size_bytes) add_bits += (e->rc.vbv_target_level - e->rc.vbv_bits) >> 4;
    }
    if (e->param.vbv_size_bytes) add_bits = IMIN(add_bits, (e->param.vbv_size_bytes * 8 * 7 >> 3) - e->rc.vbv_bits);

    bit_budget += add_bits;
    bit_budget = IMIN(bit_budget, e->run.desired_frame_bytes * 8 * 16);
    bit_budget = IMAX(bit_budget, e->run.desired_frame_bytes * 8 >> 2);

    if (is_intra) e->rc.vbv_target_level = e->rc.vbv_bits + bit_budget - e->run.desired_frame_bytes * 8;
    e->rc.vbv_target_level -= e->run.desired_frame_bytes * 8 - nominal_p;
    e->rc.vbv_target_level = IMAX(e->rc.vbv_target_level, stationary);
    e->rc.bit_budget = bit_budget;

    {
        const uint16_t *bits = h264e_bits_per_mb[!!is_intra];
        for (qp = 0; qp < 42 - 1; qp++)
            if (bits[qp] * nmb < bit_budget) break;
        qp += MIN_QP;
        qp += e->rc.dqp_smooth;
        if (e->rc.prev_qp > qp + 1) qp = (e->rc.prev_qp + qp + 1) / 2;
    }
    qp = clamp_qp(e, qp);
    e->rc.qp = qp;
    e->rc.qp_smooth = qp << 8;
    e->rc.prev_qp = qp;
}

/* rate-control state update after the frame (rc_frame_end H:6075-6141).  Returns the
 * number of filler bytes the caller must emit (vbv_underflow_stuffing_flag). */
static int rc_frame_end(h264e_host_t *e, int intra_flag, int skip_flag)
{
    int filler = 0;
    if (!skip_flag)
    {
        int qp, nmb = e->nmb;
        for (qp = 0; qp != 41 && h264e_bits_per_mb[intra_flag][qp] * nmb > (int)e->out_pos * 8 - 32; qp++) {}
        qp += MIN_QP;
        if ((e->rc.qp_smooth >> 8) - e->rc.dqp_smooth < qp - 1) e->rc.dqp_smooth--;
        else if ((e->rc.qp_smooth >> 8) - e->rc.dqp_smooth > qp + 1) e->rc.dqp_smooth++;
        if (intra_flag) e->rc.max_dqp = e->rc.dqp_smooth;
        else e->rc.max_dqp = IMAX(e->rc.max_dqp, (e->rc.qp_smooth >> 8) - qp);
    }
    e->rc.vbv_bits += e->out_pos * 8 - e->run.desired_frame_bytes * 8;
    if (e->param.vbv_size_bytes)
    {
        if (e->rc.vbv_bits < 0)
        {
            if (e->param.vbv_underflow_stuffing_flag)
            {
                do { filler++; e->rc.vbv_bits += 8; } while (e->rc.vbv_bits < 0);
            } else e->rc.vbv_bits = 0;
        }
        if (e->rc.vbv_bits > e->param.vbv_size_bytes * 8)
        {
            if (!e->param.vbv_overflow_empty_frame_flag) e->rc.vbv_bits = e->param.vbv_size_bytes * 8;
        }
    } else e->rc.vbv_bits = 0;
    return filler;
}

/* ------------------------------------------------------------------------------ */
/* sizes / parameter checks                                                         */
/* ------------------------------------------------------------------------------ */
static int check_create_params(const H264E_create_param_t *par)     /* H:6252 */
{
    if (!par) return H264E_STATUS_BAD_ARGUMENT;
    if ((int)(par->vbv_size_bytes | par->gop) < 0) return H264E_STATUS_BAD_PARAMETER;
    if (par->width <= 0 || par->height <= 0) return H264E_STATUS_BAD_PARAMETER;
    if ((unsigned)(par->const_input_flag | par->fine_rate_control_flag |
                   par->vbv_overflow_empty_frame_flag | par->vbv_underflow_stuffing_flag) > 1)
        return H264E_STATUS_BAD_PARAMETER;
    if ((unsigned)par->max_long_term_reference_frames > 8) return H264E_STATUS_BAD_PARAMETER;
    if ((par->width | par->height) & 1) return H264E_STATUS_SIZE_NOT_MULTIPLE_2;
    if (((par->width | par->height) & 15) && !par->const_input_flag) return H264E_STATUS_SIZE_NOT_MULTIPLE_16;
    return H264E_STATUS_SUCCESS;
}

/* The reference carves its blobs with 16-byte aligned bump allocation starting from
 * address 1 (H:6185-6230, H:6300-6304); reproduce the resulting sizes. */
static size_t bump(size_t p, size_t size) { return ((p + 15) & ~(size_t)15) + size; }

static void ref_sizes(const H264E_create_param_t *par, int *persist, int *scratch)
{
    int nmbx = (par->width + 15) >> 4, nmby = (par->height + 15) >> 4;
    int nref = 1 + par->max_long_term_reference_frames + par->const_input_flag + !!par->temporal_denoise_flag;
    size_t p = bump(1, (size_t)((nmbx + 2) * (nmby + 2) * 384) * nref);
    *persist = (int)(((p - 1) + 15) & ~(size_t)15) + REF_SIZEOF_ENC;
    p = bump(1, REF_SIZEOF_SCRATCH_T);
    p = bump(p, (size_t)(nmbx * nmby * (384 + 2 + 10) * 3 / 2));
    p = bump(p, (size_t)(nmbx * 8 + 8));
    p = bump(p, (size_t)(nmbx * 4 + 8) * 4);
    p = bump(p, (size_t)(nmbx * 4 + 4));
    p = bump(p, (size_t)nmbx);
    p = bump(p, (size_t)nmbx);
    p = bump(p, (size_t)nmbx);
    p = bump(p, (size_t)(nmbx * 32 + 32 + 16));
    *scratch = (int)(p - 1);
}

int H264E_sizeof(const H264E_create_param_t *par, int *sizeof_persist, int *sizeof_scratch)
{
    int error = check_create_params(par);
    if (!sizeof_persist || !sizeof_scratch) error = H264E_STATUS_BAD_ARGUMENT;
    if (error) return error;
    ref_sizes(par, sizeof_persist, sizeof_scratch);
    if (par->num_layers > 1)
    {
        /* the reference adds a second layer's blobs (H:6874-6889); SVC itself is unsupported */
        int p2, s2;
        ref_sizes(par, &p2, &s2);
        *sizeof_persist += p2 - REF_SIZEOF_ENC + REF_SIZEOF_ENC;
        *sizeof_scratch += s2;
    }
    return H264E_STATUS_SUCCESS;
}

static int unsupported_create(const H264E_create_param_t *p)
{
    return p->fine_rate_control_flag || p->max_long_term_reference_frames || p->num_layers > 1;
}

int H264E_init(H264E_persist_t *penc, const H264E_create_param_t *opt)
{
    h264e_host_t *e = (h264e_host_t *)penc;
    h264b200_ctx *ctx = NULL;
    int err;
    if (!e || !opt) return H264E_STATUS_BAD_ARGUMENT;
    err = check_create_params(opt);
    if (err) return err;
    if (unsupported_create(opt)) return H264E_STATUS_UNSUPPORTED;

    pthread_mutex_lock(&g_live_lock);
    if (live_contains_locked(e))
    {   /* re-initialisation of a live session: release its device state first */
        live_remove_locked(e);
        if (e->magic == H264E_MAGIC && e->self == e && e->ctx) h264b200_ctx_destroy(e->ctx);
    }
    pthread_mutex_unlock(&g_live_lock);

    if (h264b200_ctx_create(&ctx, opt->width, opt->height, -1) != 0 || !ctx) return H264E_STATUS_NO_DEVICE;

    memset(e, 0, sizeof(*e));
    e->magic = H264E_MAGIC;
    e->self = e;
    e->param = *opt;
    e->ctx = ctx;
    e->nmbx = (opt->width + 15) >> 4;
    e->nmby = (opt->height + 15) >> 4;
    e->nmb = e->nmbx * e->nmby;
    e->w16 = e->nmbx * 16;
    e->h16 = e->nmby * 16;
    e->cropping = !!((opt->width | opt->height) & 15);

    pthread_mutex_lock(&g_live_lock);
    e->next_live = g_live;
    g_live = e;
    if (!g_atexit_registered) { atexit(live_cleanup_at_exit); g_atexit_registered = 1; }
    pthread_mutex_unlock(&g_live_lock);
    return H264E_STATUS_SUCCESS;
}

void H264E_close(H264E_persist_t *penc)
{
    h264e_host_t *e = (h264e_host_t *)penc;
    int live;
    if (!e) return;
    pthread_mutex_lock(&g_live_lock);
    live = live_contains_locked(e);
    if (live) live_remove_locked(e);
    pthread_mutex_unlock(&g_live_lock);
    if (live && e->magic == H264E_MAGIC && e->ctx)
    {
        h264b200_ctx_destroy(e->ctx);
        e->ctx = NULL;
        e->magic = 0;
    }
}

void H264E_set_vbv_state(H264E_persist_t *penc, int vbv_size_bytes, int vbv_fullness_bytes)    /* H:6898 */
{
    h264e_host_t *e = (h264e_host_t *)penc;
    if (!e) return;
    e->param.vbv_size_bytes = vbv_size_bytes;
    if (vbv_fullness_bytes >= 0)
    {
        e->rc.vbv_bits = vbv_fullness_bytes * 8;
        e->rc.vbv_target_level = e->rc.vbv_bits;
    }
}

/* ------------------------------------------------------------------------------ */
/* one frame = plan (host) -> device job -> assemble (host)                          */
/* ------------------------------------------------------------------------------ */
typedef struct
{
    h264e_host_t *e;
    int frame_type, long_term_idx_use, long_term_idx_update;
    int slice_type, is_key;
    hbits_t hdr;
    int status;
    int transparent;     /* VBV overflow: the frame is coded as one run of skipped macroblocks, no macroblock work (H:6497-6508) */
    int dn_only;         /* ... but the temporal noise suppressor still sees the picture (H:6686 runs before the decision):
                            the device job of such a frame only advances the filter state */
    int job;             /* index of the frame's device job in the submission, -1: none */
    H264E_io_yuv_t inplace;   /* caller planes that receive the reconstruction when const_input_flag == 0 */
} frame_plan_t;

static void fill_frame_params(h264e_host_t *e, h264b200_frame_params *p, int slice_type, int hdr_bits)
{
    const h264e_qp_tunables_t *t = &h264e_qp_tunables[e->rc.qp];
    int c, qp = e->rc.qp;
    memset(p, 0, sizeof(*p));
    p->slice_type = slice_type;
    p->qp = qp;
    p->speed = e->run.encode_speed;
    p->disable_deblock = e->disable_deblock;
    p->lambda_q4 = t->lambda_q4; p->lambda_mv_q4 = t->lambda_mv_q4;
    p->lambda_i4_q4 = t->lambda_i4_q4; p->lambda_i16_q4 = t->lambda_i16_q4;
    p->skip_thr_inter = t->skip_thr_inter; p->skip_thr_i4x4 = t->skip_thr_i4x4;
    for (c = 0; c < 2; c++)
    {   /* every MB of the frame has the same QP, so (qp_p + qp_q + 1) >> 1 == qp (H:5673-5696) */
        const uint8_t *lut = h264e_deblock_tab[qp - 10];
        p->df_alpha[c] = lut[0];
        p->df_beta[c] = lut[4];
        p->df_tc0[c][0] = 0; p->df_tc0[c][1] = lut[1]; p->df_tc0[c][2] = lut[2]; p->df_tc0[c][3] = lut[3];
        qp = h264e_qp_chroma[qp];
    }
    build_qdat(p->qdat, e->rc.qp, slice_type);
    p->hdr_bits = hdr_bits;
    p->denoise = e->param.temporal_denoise_flag && e->run.encode_speed < 2;      /* H:6686 */
}

/* everything H264E_encode does before the macroblock loop (H:6654-6811, H:6477-6486) */
static int plan_frame(h264e_host_t *e, H264E_scratch_t *scratch, const H264E_run_param_t *opt,
                      H264E_io_yuv_t *in, frame_plan_t *pl, h264b200_job *job)
{
    int sp, ss, i;
    memset(pl, 0, sizeof(*pl));
    pl->e = e;
    if (!e || e->magic != H264E_MAGIC || e->self != e || !e->ctx) return H264E_STATUS_BAD_ARGUMENT;
    if (!scratch || !in) return H264E_STATUS_BAD_ARGUMENT;
    ref_sizes(&e->param, &sp, &ss);
    e->out = (uint8_t *)scratch;
    e->out_cap = ss;
    e->out_pos = 0;
    if (opt) e->run = *opt;
    if (e->run.desired_nalu_bytes) return H264E_STATUS_UNSUPPORTED;
    if (!e->run.qp_max || e->run.qp_max > 51) e->run.qp_max = 51;
    if (!e->run.qp_min || e->run.qp_min < MIN_QP) e->run.qp_min = MIN_QP;
    e->disable_deblock = (e->run.encode_speed == 8 || e->run.encode_speed == 10);

    pl->frame_type = e->run.frame_type;
    if (pl->frame_type == H264E_FRAME_TYPE_DEFAULT) pl->frame_type = e->frame_num ? H264E_FRAME_TYPE_P : H264E_FRAME_TYPE_KEY;
    switch (pl->frame_type)
    {
    default:
    case H264E_FRAME_TYPE_I:         pl->long_term_idx_use = -1; pl->long_term_idx_update = 0; break;
    case H264E_FRAME_TYPE_KEY:       pl->long_term_idx_use = -1; pl->long_term_idx_update = 0; break;
    case H264E_FRAME_TYPE_GOLDEN:    pl->long_term_idx_use = 1; pl->long_term_idx_update = 1; break;
    case H264E_FRAME_TYPE_RECOVERY:  pl->long_term_idx_use = 1; pl->long_term_idx_update = 0; break;
    case H264E_FRAME_TYPE_P:         pl->long_term_idx_use = e->most_recent_ref_frame_idx; pl->long_term_idx_update = 0; break;
    case H264E_FRAME_TYPE_DROPPABLE: pl->long_term_idx_use = e->most_recent_ref_frame_idx; pl->long_term_idx_update = -1; break;
    case H264E_FRAME_TYPE_CUSTOM:
        pl->long_term_idx_use = e->run.long_term_idx_use;
        pl->long_term_idx_update = e->run.long_term_idx_update;
        if (!pl->long_term_idx_use) pl->long_term_idx_use = e->most_recent_ref_frame_idx;
        if (pl->long_term_idx_use < 0) pl->frame_type = H264E_FRAME_TYPE_KEY;
        break;
    }
    if (pl->long_term_idx_update >= 0) e->most_recent_ref_frame_idx = pl->long_term_idx_update;
    pl->is_key = pl->frame_type == H264E_FRAME_TYPE_KEY;
    if (pl->is_key)
    {
        int q = 30;
        q = IMIN(q, e->run.qp_max);
        q = IMAX(q, e->run.qp_min);
        e->pic_init_qp = q;
        e->next_idr_pic_id ^= 1;
        e->frame_num = 0;
        write_sps(e);
        write_pps(e);
    } else
    {
        if (!e->pic_init_qp) return H264E_STATUS_BAD_FRAME_TYPE;
        /* no long-term buffers exist (max_long_term_reference_frames == 0), H:6805-6810 */
        if (pl->long_term_idx_use > 0 || pl->long_term_idx_update > 0) return H264E_STATUS_BAD_FRAME_TYPE;
    }
    pl->slice_type = pl->long_term_idx_use < 0 ? SLICE_I : SLICE_P;
    rc_frame_start(e, pl->long_term_idx_use < 0);
    write_slice_header(e, &pl->hdr, pl->slice_type, pl->is_key, pl->long_term_idx_update);

    /* VBV overflow: a "transparent" frame -- slice header, one mb_skip_run covering the picture, and the
     * reconstruction is the reference picture as it is (H:6497-6508).  Never an I / KEY frame
     * (long_term_idx_use is -1 for those), never a frame that updates a long-term buffer. */
    if (e->param.vbv_size_bytes && !pl->long_term_idx_use && pl->long_term_idx_update <= 0 &&
        e->rc.vbv_bits - e->run.desired_frame_bytes * 8 > e->param.vbv_size_bytes * 8)
    {
        pl->transparent = 1;
        memset(&pl->inplace, 0, sizeof(pl->inplace));
        if (!e->param.const_input_flag && in->yuv[0]) pl->inplace = *in;
        if (e->param.temporal_denoise_flag && e->run.encode_speed < 2)
        {   /* H:6686-6696: the filter has already consumed this picture when the reference gets here */
            memset(job, 0, sizeof(*job));
            job->ctx = e->ctx;
            job->p.denoise = 2;
            for (i = 0; i < 3; i++) { job->yuv[i] = in->yuv[i]; job->stride[i] = in->stride[i]; }
            job->preloaded_index = in->yuv[0] ? -1 : in->stride[0];
            pl->dn_only = 1;
        }
        return H264E_STATUS_SUCCESS;
    }
    memset(job, 0, sizeof(*job));
    job->ctx = e->ctx;
    fill_frame_params(e, &job->p, pl->slice_type, hb_bits(&pl->hdr));
    for (i = 0; i < 3; i++) { job->yuv[i] = in->yuv[i]; job->stride[i] = in->stride[i]; }
    job->preloaded_index = in->yuv[0] ? -1 : in->stride[0];
    job->update_ref = pl->long_term_idx_update != -1;
    if (!e->param.const_input_flag && in->yuv[0])
    {   /* the reference reconstructs in place over the caller's frame (H:6719-6723) */
        for (i = 0; i < 3; i++) { job->recon[i] = in->yuv[i]; job->recon_stride[i] = in->stride[i]; }
    }
    return H264E_STATUS_SUCCESS;
}

/* everything after the macroblock loop: slice NAL, RC update, GOP counter
 * (H:6451-6456, H:6596-6614) */
static int finish_frame(frame_plan_t *pl, h264b200_job *job)
{
    h264e_host_t *e = pl->e;
    int nbits, nbytes, i, filler;
    uint8_t *d, *nal;
    int zeros = 0, j = 0;
    const uint32_t *words;
    if (pl->transparent)
    {
        if (pl->dn_only && (!job || job->status)) return H264E_STATUS_DEVICE_ERROR;
        h264b200_note_transparent(e->ctx);
        hb_ue(&pl->hdr, (uint32_t)e->nmb);
        hb_trailing(&pl->hdr);
        if ((int)e->out_pos + 4 + 2 * pl->hdr.nbytes + 64 > e->out_cap) return H264E_STATUS_OUTPUT_OVERFLOW;
        emit_nal(e, pl->hdr.buf, pl->hdr.nbytes);
        if (pl->inplace.yuv[0])
        {   /* in-place mode: the caller's frame receives the reconstruction = the unchanged reference picture */
            unsigned char *planes[3];
            int strides[3], w16 = e->w16, h16 = e->h16, c, r;
            unsigned char *tmp = (unsigned char *)malloc((size_t)w16 * h16 * 3 / 2);
            if (!tmp) return H264E_STATUS_DEVICE_ERROR;
            planes[0] = tmp; planes[1] = tmp + (size_t)w16 * h16; planes[2] = planes[1] + (size_t)w16 * h16 / 4;
            strides[0] = w16; strides[1] = strides[2] = w16 / 2;
            if (h264b200_get_recon(e->ctx, planes, strides)) { free(tmp); return H264E_STATUS_DEVICE_ERROR; }
            for (c = 0; c < 3; c++)
            {
                int cw = c ? e->param.width / 2 : e->param.width, ch = c ? e->param.height / 2 : e->param.height;
                for (r = 0; r < ch; r++) memcpy(pl->inplace.yuv[c] + (size_t)r * pl->inplace.stride[c], planes[c] + (size_t)r * strides[c], (size_t)cw);
            }
            free(tmp);
        }
        filler = rc_frame_end(e, 0, 1);
        goto after_rc;
    }
    if (!job) return H264E_STATUS_DEVICE_ERROR;
    if (job->status) return job->status == -1 ? H264E_STATUS_NO_DEVICE : H264E_STATUS_DEVICE_ERROR;
    words = job->out_words;
    if (!words) return H264E_STATUS_DEVICE_ERROR;      /* a failed submission never hands out a payload */

    /* the device left hdr_bits of room at the front of the payload: merge the header,
     * add the RBSP stop bit and convert MSB-first words to escaped bytes */
    nbits = job->out_bits + 1;
    nbytes = (nbits + 7) >> 3;
    if ((int)e->out_pos + 4 + nbytes + 64 > e->out_cap) return H264E_STATUS_OUTPUT_OVERFLOW;
    d = e->out + e->out_pos;
    d[0] = d[1] = d[2] = 0; d[3] = 1;
    nal = d + 4;
    {
        int hdr_full = pl->hdr.nbytes, hdr_rem = pl->hdr.nacc;
        uint32_t hdr_tail = hdr_rem ? (uint32_t)((pl->hdr.acc & ((1u << hdr_rem) - 1)) << (8 - hdr_rem)) : 0;
        int stop_byte = job->out_bits >> 3;
        uint8_t stop_mask = (uint8_t)(0x80 >> (job->out_bits & 7));
        /* room for the worst case (an emulation-prevention byte after every second byte) is checked once; only a
         * nearly full buffer pays for the per-byte checks */
        const int roomy = (long)e->out_pos + 4 + nbytes + nbytes / 2 + 16 <= (long)e->out_cap;
        for (i = 0; i < nbytes; i++)
        {
            /* fast path: a whole word past the header, before the stop bit, without any zero byte and with
             * fewer than two pending zeros needs no emulation prevention: one byte-swapped 32-bit store */
            if (!(i & 3) && i > hdr_full && i + 4 <= stop_byte && zeros < 2)
            {
                const uint32_t wv = words[i >> 2];
                if (!((wv - 0x01010101u) & ~wv & 0x80808080u))
                {
                    const uint32_t be = __builtin_bswap32(wv);
                    memcpy(nal + j, &be, 4);
                    j += 4; i += 3; zeros = 0;
                    if (!roomy && (int)e->out_pos + 4 + j + 8 > e->out_cap) return H264E_STATUS_OUTPUT_OVERFLOW;
                    continue;
                }
            }
            uint8_t byte = (uint8_t)(words[i >> 2] >> (24 - 8 * (i & 3)));
            if (i < hdr_full) byte |= pl->hdr.buf[i];
            else if (i == hdr_full) byte |= (uint8_t)hdr_tail;
            if (i == stop_byte) byte |= stop_mask;
            if (zeros == 2 && byte <= 3) { nal[j++] = 3; zeros = 0; }
            zeros = byte ? 0 : zeros + 1;
            nal[j++] = byte;
            if (!roomy && (int)e->out_pos + 4 + j + 8 > e->out_cap) return H264E_STATUS_OUTPUT_OVERFLOW;
        }
    }
    if (e->run.nalu_callback) e->run.nalu_callback(nal, j, e->run.nalu_callback_token);
    e->out_pos += 4 + j;

    filler = rc_frame_end(e, pl->long_term_idx_use == -1, job->trailing_skip_run == e->nmb);
after_rc:
    if (filler)
    {   /* filler_data NAL (H:6113-6122).  0xFF bytes need no escapes: 4 + filler + 2 bytes.  The reference does not look at
         * the room left in the scratch buffer here (tiny pictures at high bit rates overrun it); this layer reports it. */
        if ((long)e->out_pos + 4 + (long)filler + 2 > (long)e->out_cap) return H264E_STATUS_OUTPUT_OVERFLOW;
        uint8_t *f = (uint8_t *)malloc((size_t)filler + 2);
        if (f)
        {
            f[0] = 12;
            memset(f + 1, 0xFF, (size_t)filler);
            f[filler + 1] = 0x80;
            emit_nal(e, f, filler + 2);
            free(f);
        }
    }
    if (pl->long_term_idx_update != -1)
    {
        if (++e->frame_num >= e->param.gop && e->param.gop && e->run.frame_type == H264E_FRAME_TYPE_DEFAULT)
            e->frame_num = 0;
    }
    return H264E_STATUS_SUCCESS;
}

int H264E_encode_batch(int n, H264E_persist_t *const *enc, H264E_scratch_t *const *scratch,
                       const H264E_run_param_t *const *run_param, H264E_io_yuv_t *const *frame,
                       unsigned char **coded_data, int *sizeof_coded_data)
{
    frame_plan_t *plans;
    h264b200_job *jobs;
    int i, err = 0, njobs = 0;
    if (n <= 0 || !enc || !scratch || !frame || !coded_data || !sizeof_coded_data) return H264E_STATUS_BAD_ARGUMENT;
    plans = (frame_plan_t *)calloc((size_t)n, sizeof(*plans));
    jobs = (h264b200_job *)calloc((size_t)n, sizeof(*jobs));
    if (!plans || !jobs) { free(plans); free(jobs); return H264E_STATUS_BAD_ARGUMENT; }
    {
    double t0 = host_now(), t1, t2;
    for (i = 0; i < n; i++)
    {
        plans[i].status = plan_frame((h264e_host_t *)enc[i], scratch[i], run_param ? run_param[i] : NULL, frame[i],
                                     &plans[i], &jobs[njobs]);
        plans[i].job = -1;
        if (plans[i].status) { if (!err) err = plans[i].status; }
        else if (!plans[i].transparent || plans[i].dn_only) plans[i].job = njobs++;
    }
    t1 = host_now();
    /* a failed submission leaves status != 0 / out_words == NULL in every job it could not finish: finish_frame turns
     * that into H264E_STATUS_DEVICE_ERROR for the sessions concerned */
    if (njobs) (void)h264b200_encode_frames(njobs, jobs);
    t2 = host_now();
    g_host_ms[0] += (t1 - t0) * 1e3; g_host_ms[1] += (t2 - t1) * 1e3; g_host_t_finish0 = t2;
    }
    for (i = 0; i < n; i++)
    {
        if (plans[i].status) { coded_data[i] = NULL; sizeof_coded_data[i] = 0; continue; }
        plans[i].status = finish_frame(&plans[i], plans[i].job >= 0 ? &jobs[plans[i].job] : NULL);
        if (plans[i].status) { if (!err) err = plans[i].status; coded_data[i] = NULL; sizeof_coded_data[i] = 0; continue; }
        coded_data[i] = plans[i].e->out;
        sizeof_coded_data[i] = (int)plans[i].e->out_pos;
    }
    g_host_ms[2] += (host_now() - g_host_t_finish0) * 1e3;
    free(plans);
    free(jobs);
    return err;
}

/* developer statistic: accumulated host milliseconds of H264E_encode_batch: [0] planning (RC, headers), [1] the device
 * submission (h264b200_encode_frames, blocking), [2] NAL assembly + RC update */
void H264E_b200_host_timing(double out[3]) { int i; for (i = 0; i < 3; i++) out[i] = g_host_ms[i]; }

int H264E_encode(H264E_persist_t *enc, H264E_scratch_t *scratch, const H264E_run_param_t *opt,
                 H264E_io_yuv_t *in, unsigned char **coded_data, int *sizeof_coded_data)
{
    frame_plan_t plan;
    h264b200_job job;
    int err;
    if (!coded_data || !sizeof_coded_data) return H264E_STATUS_BAD_ARGUMENT;
    err = plan_frame((h264e_host_t *)enc, scratch, opt, in, &plan, &job);
    if (err) return err;
    if (!plan.transparent || plan.dn_only)
    {
        const int rc = h264b200_encode_frames(1, &job);
        if (rc && !job.status) job.status = rc;       /* never assemble a NAL from a failed submission */
    }
    err = finish_frame(&plan, (!plan.transparent || plan.dn_only) ? &job : NULL);
    if (err) return err;
    *sizeof_coded_data = (int)plan.e->out_pos;
    *coded_data = plan.e->out;
    return H264E_STATUS_SUCCESS;
}

int H264E_get_recon(H264E_persist_t *penc, unsigned char *y, unsigned char *u, unsigned char *v)
{
    h264e_host_t *e = (h264e_host_t *)penc;
    unsigned char *planes[3];
    int strides[3];
    if (!e || e->magic != H264E_MAGIC || !e->ctx) return H264E_STATUS_BAD_ARGUMENT;
    planes[0] = y; planes[1] = u; planes[2] = v;
    strides[0] = e->w16; strides[1] = strides[2] = e->w16 / 2;
    return h264b200_get_recon(e->ctx, planes, strides) ? H264E_STATUS_DEVICE_ERROR : H264E_STATUS_SUCCESS;
}

int H264E_prefetch(H264E_persist_t *penc, const H264E_io_yuv_t *next)
{
    h264e_host_t *e = (h264e_host_t *)penc;
    const unsigned char *yuv[3];
    int i;
    if (!e || e->magic != H264E_MAGIC || !e->ctx || !next || !next->yuv[0]) return H264E_STATUS_BAD_ARGUMENT;
    for (i = 0; i < 3; i++) yuv[i] = next->yuv[i];
    return h264b200_prefetch_input(e->ctx, yuv, next->stride) ? H264E_STATUS_DEVICE_ERROR : H264E_STATUS_SUCCESS;
}

int H264E_preload(H264E_persist_t *penc, int nframes, const unsigned char *frames)
{
    h264e_host_t *e = (h264e_host_t *)penc;
    if (!e || e->magic != H264E_MAGIC || !e->ctx || nframes <= 0 || !frames) return H264E_STATUS_BAD_ARGUMENT;
    return h264b200_preload(e->ctx, nframes, frames) ? H264E_STATUS_DEVICE_ERROR : H264E_STATUS_SUCCESS;
}

/* device context of a session, for the bench / tests (kernel-only timing) */
h264b200_ctx *H264E_b200_ctx(H264E_persist_t *penc)
{
    h264e_host_t *e = (h264e_host_t *)penc;
    return (e && e->magic == H264E_MAGIC) ? e->ctx : NULL;
}
