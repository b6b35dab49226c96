# oracle/opcount.sed -- TEST / MEASUREMENT INFRASTRUCTURE.  Applied by oracle/Makefile to a TEMPORARY copy of the
# reference header (mktemp, deleted after the compile; nothing of the reference enters the repo): adds one counter
# statement at the top of the leaf functions whose work SURVEY.md 8(d) defines as "algorithmic integer work":
#   [0] SAD 1 op / sample            [1] six-tap filter 6 MAC / sample (centre position: 12)
#   [2] quarter-sample average 1 / sample   [3] chroma bilinear 4 MAC / sample
#   [4] forward 4x4 transform 64 add/shift per block   [5] inverse 4x4 transform 64 per block
#   [6] quantiser + dequantiser 2 multiplies / coefficient
#   [7] intra prediction + Intra4x4 mode SADs (2 ops / sample / mode tried)   [8] deblocking line filters (per sample of an edge)
/^static int sad_block(/,/^{/ s/^{/{ REF_CNT(0, w*h);/
/^static void hpel_lpf_diag(/,/^{/ s/^{/{ REF_CNT(1, 12*w*h);/
/^static void hpel_lpf_hor(/,/^{/ s/^{/{ REF_CNT(1, 6*w*h);/
/^static void hpel_lpf_ver(/,/^{/ s/^{/{ REF_CNT(1, 6*w*h);/
/^static void average_16x16_unalign(/,/^{/ s/^{/{ REF_CNT(2, 256);/
/^static void h264e_qpel_average_wh_align(/,/^{/ s/^{/{ REF_CNT(2, wh.s.x*wh.s.y);/
/^static void h264e_qpel_interpolate_chroma(/,/^{/ s/^{/{ REF_CNT(3, dxdy.u32 ? 4*wh.s.x*wh.s.y : 0);/
/^static void FwdTransformResidual4x42(/,/^{/ s/^{/{ REF_CNT(4, 64);/
/^static void TransformResidual4x4(/,/^{/ s/^{/{ REF_CNT(5, 64);/
/^static int quantize(/,/^{/ s/^{/{ REF_CNT(6, 32*(mode>>1)*(mode>>1));/
/^static int h264e_intra_choose_4x4(/,/^{/ s/^{/{ REF_CNT(7, 32*(1 + ((avail\&AVAIL_T)?3:0) + ((avail\&AVAIL_L)?2:0) + (((avail\&7)==7)?3:0)));/
/^static void h264e_intra_predict_16x16(/,/^{/ s/^{/{ REF_CNT(7, 256);/
/^static void h264e_intra_predict_chroma(/,/^{/ s/^{/{ REF_CNT(7, 128);/
/^static void deblock_chroma(/,/^{/ s/^{/{ REF_CNT(8, 6);/
/^static void deblock_luma_v(/,/^{/ s/^{/{ REF_CNT(8, 16*10);/
/^static void deblock_luma_h(/,/^{/ s/^{/{ REF_CNT(8, 16*10);/
/^static void deblock_luma_h_s4(/,/^{/ s/^{/{ REF_CNT(8, 16*14);/
/^static void deblock_luma_v_s4(/,/^{/ s/^{/{ REF_CNT(8, 16*14);/
