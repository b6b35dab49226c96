/*
 * h264-lab.h -- public C API of the B200-native encoder.
 *
 * Drop-in for the API part of the reference's single header
 * (/root/reference/src/h264-lab.h lines 1-318): same function names, argument
 * meaning, status codes and struct layouts (default build of the reference:
 * H264E_SVC_API = 1, H264E_MAX_THREADS = 0, so sizeof(H264E_create_param_t) = 56,
 * sizeof(H264E_run_param_t) = 48, sizeof(H264E_io_yuv_t) = 40).  Declarations only:
 * unlike the reference header this file carries no implementation; link against
 * libh264lab_b200.so.
 *
 * What differs behind the API: the macroblock loop runs on an NVIDIA B200 (sm_100a).
 * There is no CPU fallback -- H264E_init() fails with H264E_STATUS_NO_DEVICE when no
 * CUDA device / library is usable.
 */
#ifndef H264_LAB_B200_API_H
#define H264_LAB_B200_API_H

/* the reference header pulls these in for its users (H:325-329); its CLI (minih264e_test.c) relies on uint8_t,
 * assert and memset through it, so a drop-in header has to provide them as well */
#include <assert.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (reference H:25-34) ---- */
#define H264E_STATUS_SUCCESS                0
#define H264E_STATUS_BAD_ARGUMENT           1
#define H264E_STATUS_BAD_PARAMETER          2
#define H264E_STATUS_BAD_FRAME_TYPE         3
#define H264E_STATUS_SIZE_NOT_MULTIPLE_16   4
#define H264E_STATUS_SIZE_NOT_MULTIPLE_2    5
#define H264E_STATUS_BAD_LUMA_ALIGN         6
#define H264E_STATUS_BAD_LUMA_STRIDE        7
#define H264E_STATUS_BAD_CHROMA_ALIGN       8
#define H264E_STATUS_BAD_CHROMA_STRIDE      9
/* extensions of this implementation */
#define H264E_STATUS_NO_DEVICE              100  /* CUDA device or kernels unavailable     */
#define H264E_STATUS_UNSUPPORTED            101  /* feature outside the B200 hot path */
#define H264E_STATUS_DEVICE_ERROR           102
#define H264E_STATUS_OUTPUT_OVERFLOW        103  /* the coded frame (with its filler-data NAL) does not fit the caller's scratch buffer;
                                                    the reference writes past the buffer in that case */

/* ---- frame types (reference H:63-70) ---- */
#define H264E_FRAME_TYPE_DEFAULT    0    /* by GOP position: KEY when frame.num == 0, else P */
#define H264E_FRAME_TYPE_KEY        6    /* IDR: SPS + PPS + intra slice                      */
#define H264E_FRAME_TYPE_I          5
#define H264E_FRAME_TYPE_GOLDEN     4    /* long-term reference types: not supported          */
#define H264E_FRAME_TYPE_RECOVERY   3
#define H264E_FRAME_TYPE_P          2
#define H264E_FRAME_TYPE_DROPPABLE  1
#define H264E_FRAME_TYPE_CUSTOM     99

/* ---- speed presets (reference H:76-78) ---- */
#define H264E_SPEED_SLOWEST         0
#define H264E_SPEED_BALANCED        5
#define H264E_SPEED_FASTEST         10

/* creation parameters (reference H:83-172); 14 ints */
typedef struct H264E_create_param_tag
{
    int width;                            /* visible width, even                                   */
    int height;                           /* visible height, even                                  */
    int gop;                              /* key-frame period; 0 = only the first frame is a key  */
    int vbv_size_bytes;                   /* VBV size; also selects the level                      */
    int vbv_overflow_empty_frame_flag;
    int vbv_underflow_stuffing_flag;
    int fine_rate_control_flag;           /* per-macroblock QP: not supported (must be 0)          */
    int const_input_flag;                 /* 0: input planes are overwritten with the reconstruction */
    int max_long_term_reference_frames;   /* must be 0                                             */
    int enableNEON;                       /* ignored                                               */
    int temporal_denoise_flag;            /* 1: temporal noise suppression of the input (H:122-125); applied when encode_speed < 2 */
    int sps_id;
    int num_layers;                       /* 1 (0 is accepted as 1); SVC is not supported          */
    int inter_layer_pred_flag;
} H264E_create_param_t;

/* run-time parameters (reference H:177-226) */
typedef struct H264E_run_param_tag
{
    int encode_speed;
    int frame_type;
    int long_term_idx_use;
    int long_term_idx_update;
    int desired_frame_bytes;
    int qp_min;
    int qp_max;
    int desired_nalu_bytes;               /* must be 0 (single slice per frame)                    */
    void (*nalu_callback)(const unsigned char *nalu_data, int sizeof_nalu_data, void *token);
    void *nalu_callback_token;
} H264E_run_param_t;

/* planar YUV 4:2:0 picture (reference H:231-237) */
typedef struct H264E_io_yuv_tag
{
    unsigned char *yuv[3];
    int stride[3];
} H264E_io_yuv_t;

typedef struct H264E_persist_tag H264E_persist_t;
typedef struct H264E_scratch_tag H264E_scratch_t;

/* Memory requirements; same values as the reference reports for the same parameters
 * (reference H:264, H:6868). */
int H264E_sizeof(const H264E_create_param_t *param, int *sizeof_persist, int *sizeof_scratch);

/* Start an encoding session in caller-owned memory (reference H:277, H:6375).  Creates
 * the device-side context; calling it again on the same memory re-initialises it. */
int H264E_init(H264E_persist_t *enc, const H264E_create_param_t *param);

/* Encode one frame; *coded_data points into `scratch` (reference H:291, H:6654). */
int H264E_encode(H264E_persist_t *enc, H264E_scratch_t *scratch, const H264E_run_param_t *run_param,
                 H264E_io_yuv_t *frame, unsigned char **coded_data, int *sizeof_coded_data);

void H264E_set_vbv_state(H264E_persist_t *enc, int vbv_size_bytes, int vbv_fullness_bytes);   /* reference H:308 */

/* ---- extensions (not in the reference) ---------------------------------------- */

/* Release the device resources of a session.  The reference has no destructor (all of
 * its state lives in the caller's blobs); sessions that are never closed are released
 * at process exit. */
void H264E_close(H264E_persist_t *enc);

/* Encode one frame for each of n independent sessions in ONE device submission, so
 * that closed-GOP segments / independent streams fill the GPU together.  Per-session
 * results are exactly those of n separate H264E_encode calls.  Returns the first
 * non-zero status. */
int H264E_encode_batch(int n, H264E_persist_t *const *enc, H264E_scratch_t *const *scratch,
                       const H264E_run_param_t *const *run_param, H264E_io_yuv_t *const *frame,
                       unsigned char **coded_data, int *sizeof_coded_data);

/* Double buffering of the input: name the frame that will be passed to the NEXT H264E_encode / H264E_encode_batch call
 * of this session; the call that encodes the CURRENT frame then also starts the host->device copy of that next frame,
 * which overlaps the current frame's kernels.  The next call recognises the frame by its plane pointers and strides and
 * skips its own copy.  The planes must stay untouched until then.  Optional: output is identical with or without it. */
int H264E_prefetch(H264E_persist_t *enc, const H264E_io_yuv_t *next_frame);

/* Device-resident input: upload nframes tightly packed I420 frames (stride == width) to the
 * session's GPU; a frame whose io_yuv has yuv[0] == NULL is then taken from clip frame
 * number stride[0].  Used to measure the hot path with inputs already in HBM. */
int H264E_preload(H264E_persist_t *enc, int nframes, const unsigned char *frames);

/* Developer statistic: accumulated host milliseconds spent by H264E_encode_batch in [0] planning (rate control, headers),
 * [1] the blocking device submission, [2] NAL assembly and rate-control update. */
void H264E_b200_host_timing(double out_ms[3]);

/* Copy the reconstruction of the last encoded frame (W16 x H16, planes tightly packed
 * with strides W16, W16/2, W16/2) to host memory. */
int H264E_get_recon(H264E_persist_t *enc, unsigned char *y, unsigned char *u, unsigned char *v);

#ifdef __cplusplus
}
#endif
#endif
