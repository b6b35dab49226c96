/*
 * h264b200_shim.h -- the thin C-ABI between the host C encoder (rate control,
 * parameter sets, slice headers, NAL assembly; h264-lab_b200/host/) and the sm_100a
 * CUDA implementation of the macroblock hot path (h264-lab_b200/csrc/).
 *
 * Plain C types only (pointers and sizes, no CUDA or torch types).  The boundary
 * sits between the reference's H264E_encode_one() (H:6477: RC + headers, host) and
 * the body of encode_slice()'s macroblock loop (H:6414-6449, device): everything the
 * loop does -- mb_encode (H:5724), mb_write (H:4378), mb_deblock (H:5642) and the
 * reference-frame border extension (H:3580-3596) -- runs on the GPU.
 *
 * H:nnn = /root/reference/src/h264-lab.h line nnn.
 */
#ifndef H264B200_SHIM_H
#define H264B200_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct h264b200_ctx h264b200_ctx;   /* device-side state of one encoder instance */

/* Per-frame parameters computed by the host (replaces the fields of h264e_enc_t that
 * the macroblock loop reads: rc.qp / rc.qdat H:686-729, slice.type, speed, run_param). */
typedef struct
{
    int slice_type;          /* 0 = P, 2 = I                       (H:3203-3204)        */
    int qp;                  /* frame QP                            (H:688)              */
    int speed;               /* run_param.encode_speed              (H:181)              */
    int disable_deblock;     /* speed == 8 || speed == 10           (H:6717)             */
    /* per-QP tunables (H:1032-1120) */
    int lambda_q4, lambda_mv_q4, lambda_i4_q4, lambda_i16_q4, skip_thr_inter, skip_thr_i4x4;
    /* deblocking constants, [0] luma QP, [1] chroma QP (H:944-987, H:5663-5696) */
    int df_alpha[2], df_beta[2], df_tc0[2][4];
    unsigned short qdat[2][42];   /* quantiser tables of rc_set_qp  (H:5839-5912)        */
    int hdr_bits;            /* bit position inside the NAL at which slice_data() starts */
    int denoise;             /* 1: run the temporal noise suppressor on the input first and encode its output
                                (temporal_denoise_flag && encode_speed < 2, H:6686; h264e_denoise_run H:1547);
                                2: ONLY run the suppressor (the frame itself is coded by the host as a "transparent"
                                all-skip frame, H:6497-6508, yet the reference has filtered it by then): every other
                                field of p is ignored, the job produces no payload */
} h264b200_frame_params;

/* One frame of one encoder instance. */
typedef struct
{
    h264b200_ctx *ctx;
    h264b200_frame_params p;
    const unsigned char *yuv[3];     /* HOST input planes                                  */
    int stride[3];
    int preloaded_index;             /* >= 0: take the input from frame #index of the clip uploaded with
                                        h264b200_preload() instead of the host planes (yuv may be NULL) */
    int update_ref;                  /* 0: droppable frame, the reference picture is kept  */
    unsigned char *recon[3];         /* optional HOST buffers that receive the              */
    int recon_stride[3];             /*   reconstruction (NULL: it stays on the device)     */
    /* results */
    const unsigned int *out_words;   /* [out] HOST (pinned, owned by ctx) slice payload as 32-bit words,
                                        MSB first; slice_data() starts at bit p.hdr_bits, the bits
                                        before it are zero.  Valid until the next call on this ctx. */
    int out_bits;                    /* [out] end of the payload incl. the trailing mb_skip_run */
    int trailing_skip_run;           /* [out] skipped macroblocks at the end of the slice   */
    int status;                      /* [out] 0 = ok, -1 no device, -2 payload overflow (per-MB string or slice buffer),
                                        -3 CUDA error or bad argument (then EVERY job of the submission reports -3 and
                                        out_words is NULL), -4 the exact-wavefront verification did not converge */
} h264b200_job;

/* Create / destroy the device state for a width x height encoder.  device = CUDA
 * ordinal, or -1 for the current device.  Returns 0 or a negative error. */
int h264b200_ctx_create(h264b200_ctx **out, int width, int height, int device);
void h264b200_ctx_destroy(h264b200_ctx *ctx);
/* Reset cross-frame state (mv_clusters, H:766) -- what H264E_init's memset does. */
void h264b200_ctx_reset(h264b200_ctx *ctx);

/* Encode one frame for each of n independent encoder instances concurrently
 * (host->device copy of the inputs, macroblock pass, deblocking, CAVLC + pack,
 * device->host copy of the payload).  Blocks until every job has finished.
 * All contexts of one call must live on the same device (else -3); the call binds the calling thread to that
 * device.  Any n is accepted: batches larger than the device can hold as one wavefront submission are run as
 * consecutive submissions.  Returns 0, or the first failing job's negative status. */
int h264b200_encode_frames(int n, h264b200_job *jobs);

/* Timing of the kernels of the last h264b200_encode_frames call, in milliseconds,
 * measured with CUDA events on the launch stream: [0] whole call on device,
 * [1] macroblock pass, [2] deblock + border, [3] CAVLC + pack.  */
void h264b200_last_timing(float out_ms[4]);
/* Same, n slots (returns how many are defined): [0] whole submission on the device, [1] macroblock sweep incl. the
 * verification / repair passes, [2] in-loop filter + guard bands + half-sample planes, [3] entropy coding (CAVLC, prefix
 * sum, pack; measured on the second stream it runs on, beside [2]), [4] SAD-map pre-pass, [5] speculative motion-estimation pre-pass, [6] intra verification after sweep 0 (part of [1]), [7] reserved (0). */
int h264b200_last_timing_ex(float *out_ms, int n);

/* Upload a clip of nframes tightly packed I420 frames (stride == width) to device memory
 * owned by ctx; jobs with preloaded_index >= 0 then read their input from HBM (no
 * host->device copy inside the call).  Replaces any previously preloaded clip. */
int h264b200_preload(h264b200_ctx *ctx, int nframes, const unsigned char *frames);

/* Name the NEXT frame of ctx (same plane / stride meaning as h264b200_job.yuv) before submitting the current one: the
 * submission of the current frame then also starts the host->device copy of that next frame, on a copy stream, behind
 * its own small uploads, so that the copy runs under the current frame's kernels.  The next job of ctx whose yuv /
 * stride equal these pointers uses the staged copy instead of copying inside h264b200_encode_frames.  The caller must
 * leave the planes untouched until that job has been submitted.  Purely an optimisation: results are identical with
 * or without it. */
int h264b200_prefetch_input(h264b200_ctx *ctx, const unsigned char *const yuv[3], const int stride[3]);

/* Copy the reconstruction of the last frame of ctx (W16 x H16 luma, W16/2 x H16/2 chroma) to host: the picture the
 * last h264b200_encode_frames job wrote (also when that frame was droppable and did not become the reference), or the
 * reference picture after h264b200_note_transparent(). */
int h264b200_get_recon(h264b200_ctx *ctx, unsigned char *const planes[3], const int strides[3]);
/* The host coded a frame of ctx without a device job (transparent frame, H:6497-6508): its reconstruction is the
 * current reference picture, unchanged. */
void h264b200_note_transparent(h264b200_ctx *ctx);

/* Statistics of the exact-wavefront scheme (csrc/h264_wave.h), accumulated since the ctx was
 * created: [0] sweeps, [1] macroblocks re-encoded, [2] candidate-stage re-checks, [3] frames. */
void h264b200_ctx_stats(h264b200_ctx *ctx, int out[4]);
/* same, n slots (returns how many are defined): [4] macroblocks of P-frame sweeps taken by the decide / work fast path
 * (csrc/h264_fast.h), [5] macroblocks of those sweeps encoded by the complete path, [6], [7] reserved */
int h264b200_ctx_stats_ex(h264b200_ctx *ctx, int *out, int n);

/* Test hook: copies the SAD-map records of the last P frame of ctx (csrc/h264_sadmap.h: per macroblock the quadrant SADs
 * at 15 x 15 full-sample offsets and 13 x 13 quarter-sample positions + the motion-estimation record) to out[]; returns
 * the number of 32-bit words per macroblock, or a negative error. */
int h264b200_debug_get_sadmap(h264b200_ctx *ctx, unsigned int *out, int max_words);

/* Number of frames that found their input staged by h264b200_prefetch_input. */
long h264b200_prefetch_hits(void);
/* Number of kernel launches issued since the library was loaded. */
long h264b200_launch_count(void);
const char *h264b200_backend_name(void);

#ifdef __cplusplus
}
#endif
#endif
