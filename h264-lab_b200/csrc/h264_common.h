/*
 * h264_common.h -- shared definitions of the B200 macroblock-encode path.
 *
 * The per-macroblock code in this directory is written once, in a "warp-phase"
 * style, and compiled two ways:
 *   - by nvcc for sm_100a: one warp encodes one macroblock, FOR_LANES() strides
 *     the 32 lanes over a phase's work items, WSYNC() separates phases and
 *     wsum()/wmin() are shuffle reductions (this is the product);
 *   - by g++ with H264_EMU for the developer-only host emulation used to debug
 *     bit-exactness without a GPU (tests/_emu): FOR_LANES() is a plain loop and
 *     the reductions are identities.  The emulation is never linked into the
 *     product library.
 *
 * Rules that keep both builds equivalent: shared ("MBWork") memory is only
 * written inside FOR_LANES()/IF_LANE0 blocks; a phase never reads what the same
 * phase writes; everything outside those blocks is warp-uniform control code.
 *
 * Reference citations (H:nnn) are to /root/reference/src/h264-lab.h.
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#  define H264_DEVICE 1
#  define HD __device__ __forceinline__
#  if defined(H264_HDN_AUTO)       /* developer A/B: leave the inlining of the big leaves to the compiler */
#    define HDN static __device__
#  else
#    define HDN static __device__ __noinline__
#  endif
/* Which of the big leaves are inlined into their (single-copy) callers: bit k of H264_INL = function k in
 * the order of the HDF_ macros below.  A call costs register moves, stack traffic and scheduling freedom;
 * a copy costs instruction-cache sharing between the warps of a CTA (a function kept out of line is ONE
 * copy that every search warp runs).  The default is the best set found by A/B runs on B200. */
#  ifndef H264_INL
#    define H264_INL 0x7EFFF
#  endif
#  if ((H264_INL) >> 0) & 1
#    define HDF_mb_load static __device__ __forceinline__
#  else
#    define HDF_mb_load HDN
#  endif
#  if ((H264_INL) >> 1) & 1
#    define HDF_win_load static __device__ __forceinline__
#  else
#    define HDF_win_load HDN
#  endif
#  if ((H264_INL) >> 2) & 1
#    define HDF_inter_stage_a static __device__ __forceinline__
#  else
#    define HDF_inter_stage_a HDN
#  endif
#  if ((H264_INL) >> 3) & 1
#    define HDF_luma_tq_fast static __device__ __forceinline__
#  else
#    define HDF_luma_tq_fast HDN
#  endif
#  if ((H264_INL) >> 4) & 1
#    define HDF_me_search static __device__ __forceinline__
#  else
#    define HDF_me_search HDN
#  endif
#  if ((H264_INL) >> 5) & 1
#    define HDF_sad_qpel7 static __device__ __forceinline__
#  else
#    define HDF_sad_qpel7 HDN
#  endif
#  if ((H264_INL) >> 6) & 1
#    define HDF_sad_frame_wh static __device__ __forceinline__
#  else
#    define HDF_sad_frame_wh HDN
#  endif
#  if ((H264_INL) >> 7) & 1
#    define HDF_sad_sm_wh static __device__ __forceinline__
#  else
#    define HDF_sad_sm_wh HDN
#  endif
#  if ((H264_INL) >> 8) & 1
#    define HDF_interp_luma_planes static __device__ __forceinline__
#  else
#    define HDF_interp_luma_planes HDN
#  endif
#  if ((H264_INL) >> 9) & 1
#    define HDF_interp_chroma_block static __device__ __forceinline__
#  else
#    define HDF_interp_chroma_block HDN
#  endif
#  if ((H264_INL) >> 10) & 1
#    define HDF_intra16_pred static __device__ __forceinline__
#  else
#    define HDF_intra16_pred HDN
#  endif
#  if ((H264_INL) >> 11) & 1
#    define HDF_intra4_choose static __device__ __forceinline__
#  else
#    define HDF_intra4_choose HDN
#  endif
#  if ((H264_INL) >> 12) & 1
#    define HDF_wave_mb_first static __device__ __forceinline__
#  else
#    define HDF_wave_mb_first HDN
#  endif
#  if ((H264_INL) >> 13) & 1
#    define HDF_sad_nb8 static __device__ __forceinline__
#  else
#    define HDF_sad_nb8 HDN
#  endif
#  if ((H264_INL) >> 14) & 1
#    define HDF_copy_block static __device__ __forceinline__
#  else
#    define HDF_copy_block HDN
#  endif
#  if ((H264_INL) >> 15) & 1
#    define HDF_average_block static __device__ __forceinline__
#  else
#    define HDF_average_block HDN
#  endif
#  if ((H264_INL) >> 16) & 1
#    define HDF_sad_mb_quad static __device__ __forceinline__
#  else
#    define HDF_sad_mb_quad HDN
#  endif
#  if ((H264_INL) >> 17) & 1
#    define HDF_mvp_get static __device__ __forceinline__
#  else
#    define HDF_mvp_get HDN
#  endif
#  if ((H264_INL) >> 18) & 1
#    define HDF_partition_tasks static __device__ __forceinline__
#  else
#    define HDF_partition_tasks HDN
#  endif
#  if ((H264_INL) >> 19) & 1
#    define HDF_mc_chroma_plane static __device__ __forceinline__
#  else
#    define HDF_mc_chroma_plane HDN
#  endif
#  if ((H264_INL) >> 20) & 1
#    define HDF_chroma_tq_fast static __device__ __forceinline__
#  else
#    define HDF_chroma_tq_fast HDN
#  endif
#  if ((H264_INL) >> 21) & 1
#    define HDF_intra_chroma_plane static __device__ __forceinline__
#  else
#    define HDF_intra_chroma_plane HDN
#  endif
#  if ((H264_INL) >> 22) & 1
#    define HDF_mb_store_coefs static __device__ __forceinline__
#  else
#    define HDF_mb_store_coefs HDN
#  endif
#  if ((H264_INL) >> 23) & 1
#    define HDF_inter_mode_search static __device__ __forceinline__
#  else
#    define HDF_inter_mode_search HDN
#  endif
#  if ((H264_INL) >> 24) & 1
#    define HDF_encode_mb static __device__ __forceinline__
#  else
#    define HDF_encode_mb HDN
#  endif
#  if ((H264_INL) >> 25) & 1
#    define HDF_quant4x4 static __device__ __forceinline__
#  else
#    define HDF_quant4x4 HDN
#  endif
#  if ((H264_INL) >> 26) & 1
#    define HDF_fwd4x4 static __device__ __forceinline__
#  else
#    define HDF_fwd4x4 HDN
#  endif
#  if ((H264_INL) >> 27) & 1
#    define HDF_inv4x4_add static __device__ __forceinline__
#  else
#    define HDF_inv4x4_add HDN
#  endif
#  if ((H264_INL) >> 28) & 1
#    define HDF_coefs_small static __device__ __forceinline__
#  else
#    define HDF_coefs_small HDN
#  endif
#  if ((H264_INL) >> 29) & 1
#    define HDF_wave_mb_reencode static __device__ __forceinline__
#  else
#    define HDF_wave_mb_reencode HDN
#  endif
#  define H264_TAB static __device__ const
#else
#  define H264_DEVICE 0
#  define HD static inline
#  define HDN static
#  define HDF_mb_load static
#  define HDF_win_load static
#  define HDF_inter_stage_a static
#  define HDF_luma_tq_fast static
#  define HDF_me_search static
#  define HDF_sad_qpel7 static
#  define HDF_sad_frame_wh static
#  define HDF_sad_sm_wh static
#  define HDF_interp_luma_planes static
#  define HDF_interp_chroma_block static
#  define HDF_intra16_pred static
#  define HDF_intra4_choose static
#  define HDF_wave_mb_first static
#  define HDF_sad_nb8 static
#  define HDF_copy_block static
#  define HDF_average_block static
#  define HDF_sad_mb_quad static
#  define HDF_mvp_get static
#  define HDF_partition_tasks static
#  define HDF_mc_chroma_plane static
#  define HDF_chroma_tq_fast static
#  define HDF_intra_chroma_plane static
#  define HDF_mb_store_coefs static
#  define HDF_inter_mode_search static
#  define HDF_encode_mb static
#  define HDF_quant4x4 static
#  define HDF_fwd4x4 static
#  define HDF_inv4x4_add static
#  define HDF_coefs_small static
#  define HDF_wave_mb_reencode static
#  define H264_TAB static const
#endif

#if H264_DEVICE
#  define LANE_ID ((int)(threadIdx.x & 31))
#  define FOR_LANES(i, n) for (int i = LANE_ID; i < (n); i += 32)
#  define WSYNC() __syncwarp()
#  define IF_LANE0 if (LANE_ID == 0)
HD int wsum(int v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
HD int wmax(int v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}
HD int wor(int v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#else
#  define LANE_ID 0
#  define FOR_LANES(i, n) for (int i = 0; i < (n); i++)
#  define WSYNC() ((void)0)
#  define IF_LANE0
HD int wsum(int v) { return v; }
HD int wmax(int v) { return v; }
HD int wor(int v) { return v; }
#endif

/* One macroblock is encoded by a CTA of MB_WARPS warps: the warps split it by TASK
 * (16x16 search | 16x8 + 8x16 searches | 8x8 search | intra decision, then luma halves |
 * chroma planes).  ON_WARP(k) guards a task; the host emulation runs the tasks one after
 * the other in program order, which satisfies every producer -> consumer dependency. */
#ifndef MB_WARPS
#define MB_WARPS 4           /* 1: single-warp build of the same code (the candidate re-check kernel): tasks run one after the other */
#endif
#if H264_DEVICE && MB_WARPS == 1
#  define WARP_ID 0
#  define ON_WARP(k)
#  define CTA_SYNC() __syncwarp()
#  define FOR_THREADS(i, n) for (int i = LANE_ID; i < (n); i += 32)
#  define FOR_SEARCH_THREADS(i, n) for (int i = LANE_ID; i < (n); i += 32)
#  define IF_THREAD0 if (threadIdx.x == 0)
HD void bar_sync(int, int) {}
#elif H264_DEVICE
/* WARP_ID is the warp's ROLE.  A CTA's warp k always runs on SM sub-partition k, so with the
 * plain numbering every CTA on an SM would put its heaviest task on the same sub-partition
 * (one issue port, one L0 instruction cache); co-resident CTAs therefore rotate the roles. */
#  if defined(H264_PROFILE) || defined(H264_NO_ROLE_SPREAD)
#    define WARP_ID ((int)(threadIdx.x >> 5))
#  else
#    define WARP_ID ((int)(((threadIdx.x >> 5) + blockIdx.x / 148u + blockIdx.y) & 3))
#  endif
#  define ON_WARP(k) if (WARP_ID == (k))
#  define CTA_SYNC() __syncthreads()
#  define FOR_THREADS(i, n) for (int i = (int)threadIdx.x; i < (n); i += MB_WARPS * 32)
/* the 96 threads of the three motion-search warps (roles 0-2) */
#  define FOR_SEARCH_THREADS(i, n) for (int i = WARP_ID * 32 + LANE_ID; i < (n); i += 96)
#  define IF_THREAD0 if (threadIdx.x == 0)
HD void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
#else
#  define WARP_ID 0
#  define ON_WARP(k)
#  define CTA_SYNC() ((void)0)
#  define FOR_THREADS(i, n) for (int i = 0; i < (n); i++)
#  define FOR_SEARCH_THREADS(i, n) for (int i = 0; i < (n); i++)
#  define IF_THREAD0
HD void bar_sync(int, int) {}
#endif

typedef uint8_t pix_t;

HD void atomic_add_stat(int *p)
{
#if H264_DEVICE
    atomicAdd(p, 1);
#else
    (*p)++;
#endif
}

/* ---- constants with the reference's meaning --------------------------------- */
#define MV_NA 0x8000          /* H:3200: "no motion vector" marker (intra / unavailable) */
#define AVAIL_T 1             /* H:511-514 */
#define AVAIL_L 2
#define AVAIL_TL 4
#define AVAIL_TR 8
#define SLICE_P 0             /* H:3203-3204 */
#define SLICE_I 2
#define NNZ_NA 64             /* H:3206 */
#define MV_RANGE_PX 32        /* H:3221 */
#define QDQ_INTRA4 2          /* H:505-508 */
#define QDQ_INTER 8
#define QDQ_INTRA16 9
#define QDQ_CHROMA 5

/* MB types, H:659: -1 skip, 0 P16x16, 1 P16x8, 2 P8x16, 3 P8x8, 5 I4x4, 6 I16x16 */
#define MBT_SKIP (-1)
#define MBT_I4 5
#define MBT_I16 6

/* ---- motion vectors: two int16 packed in an int32, x in the low half (H:533-541) */
HD int mv_pack(int x, int y) { return (int)(((uint32_t)y << 16) | ((uint32_t)x & 0xFFFFu)); }
HD int mv_x(int v) { return (int)(int16_t)(v & 0xFFFF); }
HD int mv_y(int v) { return (int)(int16_t)((uint32_t)v >> 16); }
HD int mv_add2(int a, int b) { return mv_pack(mv_x(a) + mv_x(b), mv_y(a) + mv_y(b)); }
HD int mv_sub2(int a, int b) { return mv_pack(mv_x(a) - mv_x(b), mv_y(a) - mv_y(b)); }
HD int mv_round_fullpel(int v) { return mv_pack((mv_x(v) + 1) & ~3, (mv_y(v) + 1) & ~3); } /* H:3498 */
HD int iabs(int x) { return x < 0 ? -x : x; }
HD int imin(int a, int b) { return a < b ? a : b; }
HD int imax(int a, int b) { return a > b ? a : b; }
HD int clip_u8(int x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }

/* ---- per-macroblock record kept for the whole frame (HBM) ---------------------
 * Replaces the reference's one-row rolling contexts (mv_pred, nnz, i4x4mode,
 * df.*; H:742-745, H:599-605) with frame-wide arrays so that macroblocks can be
 * processed in wavefront order and CAVLC / deblocking can run as separate passes. */
struct MBInfo
{
    int32_t mv[16];      /* final MV of each 4x4 block, raster; MV_NA for intra MBs        */
    int32_t mvd[4];      /* MV difference per partition, partition order (H:4574)         */
    int8_t  type;        /* MBT_*                                                          */
    int8_t  i16_mode;    /* luma 16x16 mode 0=V 1=H 2=DC; also selects the chroma mode     */
    uint8_t cbp;         /* luma 8x8 bits | chroma (0..2) << 4                              */
    uint8_t flags;       /* bit0: row/frame bookkeeping (unused)                           */
    uint16_t nz_mask;    /* luma 4x4 blocks with coefficients, bit 15 = block 0 (H:2589)   */
    uint16_t pad0;
    int8_t  i4_mode[16]; /* actual I4x4 modes (2 = DC for every other MB type, H:4391)     */
    int8_t  i4_code[16]; /* coded value: -1 = predicted, else rem_intra4x4_pred_mode       */
    uint8_t nnz[24];     /* total_coeff of each block as seen by neighbours' CAVLC context:
                            16 luma (raster), 4 U, 4 V                                     */
};

/* Speculation record of one macroblock (P frames).  The reference feeds two running MV
 * averages ("mv_clusters", H:766, H:5263-5278, H:5382-5383) from macroblock to macroblock in
 * RASTER order, which no wavefront can honour directly.  Macroblocks are therefore decided
 * with speculated cluster candidates; this record keeps what a sequential replay of the
 * cluster trajectory needs to verify the speculation and to repair it (h264_wave.h). */
struct MBSpec
{
    int32_t mv0;         /* mb.mv[0] fed to mv_clusters_update                              */
    int32_t flags;       /* SPEC_UPDATES: MB updates the clusters (final type < 5);
                            SPEC_USED_CL: the decision consumed the cluster candidates       */
    int32_t cl_used[2];  /* rounded cluster candidates the decision was made with           */
    int32_t cand_sig[4]; /* candidate-stage outcome: mv_best, sad_best, cost_best, partition hints */
    int32_t mode_cost[4];/* cost of every partition mode that was searched                   */
    int32_t inter_best;  /* partition mode that won the inter decision                      */
    int32_t pad[3];      /* [0] cost of the inter decision                                   */
};
#define SPEC_UPDATES 1
#define SPEC_USED_CL 2
#define SPEC_NO_INTRA 4       /* decided without looking at the intra modes (sweep 0 of a P frame): wave_mb_intra_check has to confirm */

/* per-frame synchronisation words of the wavefront / verification passes */
#define FS_ARRIVE 0       /* rows that finished the current pass (monotonic)                 */
#define FS_STATE 1        /* 0..n: pass to run, FS_DONE when the frame is exact              */
#define FS_TRAJ_CHANGED 2 /* MBs whose cluster-relevant result changed in the current pass    */
#define FS_NDIRTY 3       /* MBs whose speculated candidates differ from the replayed ones    */
#define FS_PASSES 4       /* statistics: passes run                                           */
#define FS_REENC 5        /* statistics: full re-encodes                                      */
#define FS_CHECKS 6       /* statistics: candidate-stage re-checks                            */
#define FS_CL_END 8       /* [8],[9]: cluster state after the last macroblock (raw)           */
#define FS_LIVE 10        /* [10],[11]: cluster state after the raster-contiguous prefix of finished
                             macroblocks of sweep 0, published while the sweep runs; [12] prefix length */
#define FS_NFAIL 13       /* dirty macroblocks whose candidate-stage re-check failed (need a re-encode)  */
#define FS_TRAJ_FIRST 14  /* 0x3fffffff - (first macroblock whose cluster-relevant result changed in the current pass), 0: none */
#define FS_REPLAYED 7      /* GPU: repair pass whose trajectory replay the follower of the repair wave has already done */
#define FS_REPLAY_TF 15    /* GPU: FS_TRAJ_FIRST that replay started from                                   */
#define FS_WAVE_TAGS 16    /* macroblocks tagged for the repair WAVE of the current pass (successors of changes of the last parallel round) */
#define FS_FAST 17         /* statistics: macroblocks of sweep 0 taken by the decide / work fast path (h264_fast.h) */
#define FS_SLOW 18         /* statistics: macroblocks of sweep 0 encoded by the complete path */
#define FS_WAVE_REENC 19   /* statistics: macroblocks the repair WAVES looked at one after the other (h264b200_ctx_stats_ex [8]) */
#define FS_WORDS 24
#define FS_DONE 0x40000000

/* Quantised levels of one macroblock (int16), written by the encode pass and read
 * by the CAVLC pass. Layout in units of int16. */
#define COEF_Y     0      /* 16 blocks x 16 levels, block raster, level index v+4u (H:2391) */
#define COEF_YDC   256    /* 16 luma DC levels (I16x16 only)                                */
#define COEF_C     272    /* 8 blocks x 16: U0..U3, V0..V3                                  */
#define COEF_CDC   400    /* 4 U DC + 4 V DC                                                */
#define COEF_PER_MB 416

/* ---- per-frame parameters (uploaded by the host for every frame) --------------- */
struct FrameParams
{
    int width, height;          /* visible size                                             */
    int nmbx, nmby;
    int slice_type;             /* SLICE_P / SLICE_I                                        */
    int qp;                     /* luma QP of the frame (no per-MB QP: fine RC unsupported) */
    int speed;                  /* run_param.encode_speed                                   */
    int disable_deblock;
    /* per-QP tunables looked up on the host (H:1032-1120) */
    int lambda_q4, lambda_mv_q4, lambda_i4_q4, lambda_i16_q4, skip_thr_inter, skip_thr_i4x4;
    /* MV limits, absolute quarter-pel (H:6322-6325) */
    int mvlim_x0, mvlim_y0, mvlim_x1, mvlim_y1;
    /* deblocking constants for this QP: [0] luma, [1] chroma (H:944-987, H:5673-5696) */
    int df_alpha[2], df_beta[2], df_tc0[2][4];
    uint16_t qdat[2][42];       /* quantiser tables built by the host RC (H:5839-5912)      */
    /* planes */
    const pix_t *inp[3]; int inp_stride[3];
    pix_t *dec[3];              /* reconstruction being built (padded planes, pixel (0,0))  */
    const pix_t *ref[3];        /* previous reconstruction (deblocked, borders extended)    */
    const pix_t *hp[3];         /* half-sample planes b, h, j of ref[0] (same geometry, pixel (0,0)), see hpel_word() */
    pix_t *hp_out;              /* 3 planes of luma_bytes each, filled from dec[0] once the frame is finished */
    pix_t *dec_base;            /* first byte of the padded luma plane of dec                */
    int luma_bytes;             /* size of one padded luma plane                              */
    int update_ref;             /* the frame becomes the next reference picture              */
    int stride[2];              /* luma / chroma stride of dec and ref                      */
    MBInfo *mbi;
    int16_t *coef;
    int32_t *clusters;          /* persistent mv_clusters[2] of this encoder (H:766)        */
    MBSpec *spec;               /* [nmb] speculation records                                */
    int32_t *cl_true;           /* [nmb][2] rounded cluster candidates from the last replay */
    int32_t *cl_ckpt;           /* [(nmb+31)/32][2] raw cluster state before every 32nd macroblock (last replay): a later replay
                                   resumes at the block of the first macroblock that changed                                    */
    int *changed_pass;          /* [nmb] last pass in which the MB's result changed         */
    int *need_reenc;            /* [nmb] pass for which the parallel re-check asked for a re-encode */
    int *fsync;                 /* [FS_WORDS] frame synchronisation words                   */
    int *row_progress;          /* [nmby] macroblocks finished per row (encode pass)        */
    int *row_progress_mv;       /* [nmby] sweep 0: macroblocks DECIDED per row (vectors, types, speculation records written; the
                                   reconstruction may still be in the works), >= row_progress                     */
    int *row_progress_df;       /* [nmby] same for the deblock pass                         */
    int *row_progress_dfc;      /* [nmby] deblock pass, chroma wavefront                     */
    int *row_clean;             /* [nmby] last repair sweep the row went through without anything to do */
    uint32_t *mb_bits;          /* per-MB bit strings, MB_BITS_WORDS words each             */
    int *mb_nbits;              /* [nmb + 1]                                                */
    int *mb_bitoff;             /* [nmb + 1] exclusive prefix sum of mb_nbits + hdr_bits    */
    uint32_t *out_words;        /* packed slice payload                                     */
    int out_cap_words;          /* its capacity (this session's: the jobs of a submission may differ in picture size) */
    int *out_info;              /* [0] total bits, [1] error flags, [2] trailing skip run   */
    int hdr_bits;               /* bit offset at which the slice data starts                */
    int *prof;                  /* developer builds: per-MB phase cycle counts [nmb][10]    */
    int max_passes;             /* safety bound on verification sweeps                      */
    int spec_from_prev;         /* 1: speculate with the previous P frame's replayed trajectory */
    /* SAD maps of the frame's macroblocks (h264_sadmap.h), [nmb][SM_WORDS]; use_sadmap: the pre-pass ran for this frame */
    uint32_t *sadmap;
    int use_sadmap;
    /* speculative motion estimation ahead of the wavefront (h264_wave.h): use_me = the pre-pass ran for this frame;
     * me_field = [nmb][16] motion field the pre-pass predicts for THIS frame (input of its refinement rounds) */
    int use_me;
    int32_t *me_field;
    int *me_list, *me_count;    /* refinement rounds: macroblocks whose record does not belong to the predicted context any more
                                   (k_me_scan), [nmb] indices and the count per round [8] */
    /* 1: sweep 0 of this P frame decides every macroblock among the inter modes only; the intra costs of all macroblocks are
     * verified afterwards, in parallel, against the finished sweep (h264_wave.h, wave_mb_intra_check) */
    int spec_no_intra;
    /* [0]: inter cost from which a macroblock of this frame evaluates its intra modes inside sweep 0 after all (13/8 of the
     * mean inter cost of the previous P frame, written by wave_replay(predict)); [1]: inter cost from which the
     * motion-estimation pre-pass predicts an intra outcome (15/8 of that mean); [2 + y]: the same threshold as [0] for the
     * macroblocks of row y, but never below 13/8 of THAT ROW's mean (a row that is expensive as a whole -- the cropped
     * bottom row of 1080p -- is not sent through the complete path macroblock after macroblock: its intra winners are
     * found by the parallel verification and repaired in parallel rounds); only read when have_cost_stat
     * (a P frame of this session has been finished before -- the statistic survives IDR frames) */
    int *cost_stat;
    int have_cost_stat;
    int thr_eighths;            /* the "13" of the 13/8 above (developer knob H264B200_THR, A/B runs) */
    /* temporal noise suppressor (h264_denoise.h); dn_out[0] == NULL: not used for this frame */
    const pix_t *dn_src[3];     /* picture as submitted                                       */
    const pix_t *dn_prev[3];    /* previous output of the filter                              */
    pix_t *dn_out[3];           /* new output = inp[] of the macroblock path                  */
    int dn_src_stride[3], dn_stride[3];
};

/* shared-memory search window of one macroblock (luma): WIN_W x WIN_H samples */
#define WIN_W 64
#define WIN_H 48

#define MB_BITS_WORDS 512       /* 2048 bytes per macroblock */

/* SAD-map record of one macroblock (h264_sadmap.h) */
#define SM_R 7
#define SM_N (2 * SM_R + 1)
#define SM_QR 2
#define SM_QN (2 * SM_QR + 1)
#define SM_INT_ENTRIES (SM_N * SM_N)
#define SM_Q_ENTRIES (SM_QN * SM_QN)
#define SM_INT_OFF 4
#define SM_Q_OFF (SM_INT_OFF + 2 * SM_INT_ENTRIES)
#define SM_ME_OFF (SM_Q_OFF + 2 * SM_Q_ENTRIES)
/* ... followed by the macroblock's speculative motion-estimation record (h264_wave.h, me_prepass_mb): the inputs the
 * estimation was run with (the key) and everything it produced */
#define ME_KEY 0          /* [0..3] MV context left, [4..7] top-left, [8..12] top (+ top-right), [13], [14] cluster candidates, [15] 1 = valid */
#define ME_IC 16          /* MBWork::ic[16] as the candidate stage / 16x16 search left it */
#define ME_COST 32        /* mode_cost[4] */
#define ME_MV 36          /* part_mv[4][4] */
#define ME_MVD 52         /* part_mvd[4][4] */
#define ME_WORDS 80
#define SM_WORDS (SM_ME_OFF + ME_WORDS)
#define SM_INVALID 0xFFFFFFFFu

/* ---- per-macroblock working set (shared memory on the GPU) --------------------- */
/* private scratch of one motion-search warp */
/* scratch slot of search task k: the single-warp build runs the tasks one after the other in one slot */
#define SS_SLOT(k) (MB_WARPS == 1 ? 0 : (k))
struct SearchScratch
{
    pix_t store[4][256];         /* prediction variants, stride 16 (mb_pix_store, H:567)    */
    pix_t tmpblk[256];           /* second operand of quarter-sample averages               */
    pix_t probe_pad[560];        /* store + tmpblk + this: the probe window of lut_sad_lanes (PROBE_* below) */
    int32_t mvp_left[4], mvp_tl[4], mvp_top[5];   /* rolling MV predictor context (H:742)   */
};
/* probe window: the samples of G, b, h, j that a batch of neighbouring untabulated search positions needs -- 4 planes x
 * up to 19 rows x 24 bytes (+ one word that unaligned reads may touch) */
#define PROBE_ROWS 19
#define PROBE_PITCH 24

/* What prediction + transform / quantisation / reconstruction of one macroblock work on.  MBWork starts with one (its
 * members are used as w->inp_y ... everywhere); the fast path of P frames (h264_fast.h) gives every warp a private one,
 * and the functions that only touch these members (luma_tq_*, chroma_tq_*, mc_chroma_plane) are handed such a buffer
 * through MBState::w. */
#if H264_DEVICE
struct TQBuf
#else
struct alignas(16) TQBuf         /* host emulation: keeps the (MBWork *) view of a private buffer (h264_fast.h) aligned for the sanitizers */
#endif
{
    pix_t inp_y[256];            /* input MB, stride 16 (mb_pix_inp, H:566)                 */
    pix_t inp_c[128];            /* U at +0, V at +8, stride 16                             */
    pix_t predc[128];            /* chroma prediction: U at +0, V at +8, stride 16         */
    int16_t dq_y[16][16];        /* transform coefficients / dequantised (quant_t.dq)       */
    int16_t qv_y[16][16];        /* quantised levels (quant_t.qv)                           */
    int16_t dq_c[8][16];
    int16_t qv_c[8][16];
    int16_t dc_y[16], qdc_y[16]; /* luma DC: transform values / quantised levels            */
    int16_t dc_c[8], qdc_c[8];
    int8_t  zflag1[16], zflag2[16], zflagc[8];
    int32_t tq_res[8];           /* luma nz bits of the two halves, chroma nz bits / dc flags */
};

/* decision of one macroblock of a fast-path batch (h264_fast.h) */
struct FastDec
{
    int32_t type, mv_skip;
    int32_t pmv[4];
    int32_t x, pad;
};
#define FAST_BATCH 8          /* motion-estimation records staged per bulk-copy group */
#define FAST_RING 16          /* decided macroblocks the deciding warp may be ahead of the working warps */

struct MBWork : TQBuf
{
    /* inputs, read-only after mb_load (inp_y, inp_c: TQBuf) */
    pix_t top_y[24];             /* unfiltered row above: 16 + 4 of the top-right MB        */
    pix_t left_y[16];
    pix_t top_c[16];             /* U 0..7, V 8..15                                         */
    pix_t left_c[16];
    pix_t tl[4];                 /* top-left Y, U, V                                        */
    int32_t mvp0_left[4], mvp0_tl[4], mvp0_top[5];   /* MV predictor context at MB start    */
    int32_t nb_i4mode[8];        /* I4x4 modes of the left MB's right column / top MB's bottom row */
#if MB_WARPS == 1 && !defined(CHECK_WITH_WINDOW)
    uint32_t win[4];             /* single-warp build (pre-passes, re-checks): no search window, blocks are read where they lie */
#else
    uint32_t win[(WIN_W * WIN_H + 32) / 4];   /* search window: copy of the reference picture around the MV predictor */
#endif
    /* GPU: asynchronous prefetch for the NEXT macroblock of the row (cp.async): its input samples,
     * and -- once this macroblock's searches are over -- its search window, placed where this
     * macroblock's motion vector points.  Tags = 1 + macroblock index the data belongs to. */
    uint32_t pf_inp[96];         /* 64 luma words (stride 16) + 32 chroma words (U | V, stride 16) */
    int32_t pf_inp_tag, pf_win_tag, pf_win_x0, pf_win_y0;
    /* motion search */
    SearchScratch ss[MB_WARPS == 1 ? 1 : 4];   /* one per warp that may run a partition-mode search (SS_SLOT) */
    int32_t task_next;           /* partition-mode search tasks handed out so far (encode_mb) */
    pix_t mode_store[3][256];    /* assembled prediction of partition modes 1..3 (whichever warp searched them) */
    int32_t ic[16];              /* result of the candidate stage, see IC_* in h264_mbenc.h  */
    int32_t mode_cost[4], mode_pred[4];      /* per partition mode: cost, byte offset of its prediction in MBWork */
    int32_t part_mv[4][4], part_mvd[4][4];   /* per mode, per partition                     */
    /* intra */
    pix_t i16pred[256];
    pix_t i4rec[256];            /* I4x4 reconstruction                                     */
    pix_t i4r[17 * 24];          /* GPU fast path: padded I4x4 reconstruction (row above, left column) */
    pix_t i4z[16];               /* the 13 neighbours of the current 4x4 block              */
    pix_t i4s[64];               /* the 32 source values of its nine predictions (GPU: one table per half-warp) */
    int16_t i4t[16], i4u[16];    /* residual / butterfly exchange of the current 4x4 block  */
    int8_t  i4_mode[16], i4_code[16];
    int32_t intra_res[8];        /* cost16, i16 mode, cost4, nz mask of I4x4                */
    /* chroma prediction, transform / quantisation (predc, dq_*, qv_*, ...: TQBuf) */
    int32_t predc_tag, predc_mv; /* GPU: predc already holds the P16x16 chroma prediction for vector predc_mv (tag = 1) */
    pix_t skip_pred[256];        /* luma prediction at the skip vector                      */
    int32_t scal[16];            /* scalars produced by one lane for everybody              */
#if MB_WARPS == 1
    uint32_t old_mbi[4], old_rec[4];      /* single-warp build: no re-encodes, no staged replay */
    int32_t rp_mv0[1], rp_flags[1], rp_used0[1], rp_used1[1], rp_true0[1], rp_true1[1];
#else
    uint32_t old_mbi[40];        /* previous record / reconstruction of an MB being repaired */
    uint32_t old_rec[96];
    int32_t rp_mv0[32], rp_flags[32], rp_used0[32], rp_used1[32], rp_true0[32], rp_true1[32];   /* replay staging */
#endif
    /* SAD maps (h264_sadmap.h): the record of the current macroblock and, while it is encoded, the one of the next macroblock
     * of the row (bulk copy, double buffered); map_tag = 1 + macroblock index a buffer holds / was asked to hold,
     * map_cnt = copies issued into a buffer so far (phase of its barrier) */
#if MB_WARPS == 1
    uint32_t __attribute__((aligned(16))) maps[2][4];          /* single-warp build: records are read where they lie */
#else
    uint32_t __attribute__((aligned(16))) maps[2][SM_WORDS];
#endif
    unsigned long long map_bar[2];
    int32_t map_tag[2], map_cnt[2];
    int32_t pf_enable;           /* 1: the row loop of sweep 0 runs: stage the next macroblock's record while this one is encoded */
#if MB_WARPS != 1
    /* fast path of P frames (h264_fast.h): private buffers of the warps, the decisions of the current batch, the final
     * vectors of the previous macroblock of the row, the motion-estimation records of the batch (bulk copies) */
    TQBuf wb[MB_WARPS];
    pix_t wpred[MB_WARPS][256];
    FastDec fd[FAST_RING];
    int32_t last_mv[16];
    int32_t batch_n;
    int32_t q_dec;               /* decisions handed to the working warps so far (monotonic within the row) */
    int32_t q_done[MB_WARPS];    /* [k]: decisions worked off by warp k (it takes q = k - 1, k - 1 + 3, ...) */
    int32_t cmd, slow_x;         /* 0: run; 1: all warps meet to encode macroblock slow_x by the complete path; 2: row finished */
    uint32_t __attribute__((aligned(16))) me_stage[FAST_BATCH][ME_WORDS];
    unsigned long long me_bar;
    int32_t me_cnt, me_first, me_num;      /* copies completed on me_bar so far; first macroblock / number of records staged */
#endif
};

HD int mb_avail(int mbx, int mby, int nmbx)   /* single slice per frame: H:3605-3622 */
{
    int f = 0;
    if (mby > 0) f |= AVAIL_T;
    if (mby > 0 && mbx != nmbx - 1) f |= AVAIL_TR;
    if (mbx > 0) f |= AVAIL_L;
    if (mby > 0 && mbx > 0) f |= AVAIL_TL;
    return f;
}
