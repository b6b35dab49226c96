/* oracle/h264_oracle.h -- TEST INFRASTRUCTURE ONLY: plain-C restatement of the leaf
 * algorithms of the reference's macroblock path; see h264_oracle.c. */
#ifndef H264_ORACLE_H
#define H264_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int orc_sad(const uint8_t *a, int a_stride, const uint8_t *b, int b_stride, int w, int h);
int orc_sad_mb_quadrants(const uint8_t *a, int a_stride, const uint8_t *b16, int sad4[4]);
void orc_qpel_luma(const uint8_t *src, int stride, uint8_t *dst, int w, int h, int dx, int dy);
void orc_qpel_chroma(const uint8_t *src, int stride, uint8_t *dst, int w, int h, int dx, int dy);
void orc_intra16(uint8_t *pred, const uint8_t *left, const uint8_t *top, int mode);
void orc_intra_chroma(uint8_t *pred, const uint8_t *left, const uint8_t *top, int mode);
int orc_intra16_estimate(const uint8_t *p, int avail, int qp);
int orc_intra4_choose(const uint8_t *blockin, uint8_t *blockpred, int avail, const uint8_t *edge, int mpred, int penalty);
void orc_fwd4x4(const uint8_t *inp, int inp_stride, const uint8_t *pred, int16_t *out);
int orc_quant4x4(int16_t *dq, int16_t *qv, int i0, const uint16_t *qdat);
void orc_inv4x4_add(const int16_t *dq, const uint8_t *pred, uint8_t *out, int out_stride);
int orc_transform_quant(const uint8_t *inp, const uint8_t *pred, int inp_stride, int mode,
                        int16_t *q_out, int16_t *dc_out, const uint16_t *qdat);
int orc_cavlc_block(const int16_t *c, int n, int nA, int nB, uint8_t *out, int *total_coeff);
void orc_deblock_luma(uint8_t *pix, int stride, const uint8_t *strength, const uint8_t *tc0,
                      const uint8_t *alpha, const uint8_t *beta);
void orc_deblock_chroma(uint8_t *pix, int stride, const uint8_t *strength, const uint8_t *tc0,
                        const uint8_t *alpha, const uint8_t *beta);
void orc_extend_borders(uint8_t *pic, int w, int h, int guard);
void orc_denoise_run(const uint8_t *cur, uint8_t *prev, int w, int h, int cur_stride, int prev_stride);
#ifdef __cplusplus
}
#endif
#endif
