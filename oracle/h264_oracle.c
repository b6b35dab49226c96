/*
 * oracle/h264_oracle.c -- TEST INFRASTRUCTURE ONLY.  A plain-C restatement of the leaf
 * algorithms of the reference encoder's macroblock path (SURVEY.md section 8(a), rows
 * a1-a3, a7-a11, a13, a15, a16), written from the algorithm descriptions, each function
 * citing the reference lines it follows (H:nnn = /root/reference/src/h264-lab.h).
 *
 * Parity status: PINNED.  tests/test_oracle_kat.py checks every function of this file
 * against the compiled, unmodified reference (oracle/_ref/libh264ref.so, built by
 * oracle/Makefile from /root/reference) on seeded random and adversarial inputs.  The
 * whole-encoder oracle is oracle/_ref itself (the reference compiles from its single
 * header, so it is used directly rather than restated).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 */
#include "h264_oracle.h"
#include <stdlib.h>
#include <string.h>

static int iabs_(int v) { return v < 0 ? -v : v; }
static int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* ---------------- a1: sums of absolute differences (H:2162-2192) ---------------- */
int orc_sad(const uint8_t *a, int a_stride, const uint8_t *b, int b_stride, int w, int h)
{
    int x, y, s = 0;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) s += iabs_((int)a[y * a_stride + x] - (int)b[y * b_stride + x]);
    return s;
}

/* four 8x8 quadrants TL,TR,BL,BR of a 16x16 block against a stride-16 block (H:2178) */
int orc_sad_mb_quadrants(const uint8_t *a, int a_stride, const uint8_t *b16, int sad4[4])
{
    int q;
    for (q = 0; q < 4; q++)
        sad4[q] = orc_sad(a + (q >> 1) * 8 * a_stride + (q & 1) * 8, a_stride, b16 + (q >> 1) * 128 + (q & 1) * 8, 16, 8, 8);
    return sad4[0] + sad4[1] + sad4[2] + sad4[3];
}

/* ---------------- a2: luma quarter-sample interpolation (H:1971-2130) ----------- */
static int six_tap(const uint8_t *p, int step) { return p[0] - 5 * p[step] + 20 * p[2 * step] + 20 * p[3 * step] - 5 * p[4 * step] + p[5 * step]; }

static void half_hor(const uint8_t *src, int stride, uint8_t *dst, int w, int h)      /* H:2029 */
{
    int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) dst[y * 16 + x] = (uint8_t)clip255((six_tap(src + y * stride + x - 2, 1) + 16) >> 5);
}
static void half_ver(const uint8_t *src, int stride, uint8_t *dst, int w, int h)      /* H:2041 */
{
    int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) dst[y * 16 + x] = (uint8_t)clip255((six_tap(src + (y - 2) * stride + x, stride) + 16) >> 5);
}
static void half_diag(const uint8_t *src, int stride, uint8_t *dst, int w, int h)     /* H:1990 */
{
    int16_t mid[21][16];
    int x, y;
    for (y = 0; y < h + 5; y++)
        for (x = 0; x < w; x++) mid[y][x] = (int16_t)six_tap(src + (y - 2) * stride + x - 2, 1);
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++)
        {
            int v = mid[y][x] - 5 * mid[y + 1][x] + 20 * mid[y + 2][x] + 20 * mid[y + 3][x] - 5 * mid[y + 4][x] + mid[y + 5][x];
            dst[y * 16 + x] = (uint8_t)clip255((v + 512) >> 10);
        }
}
static void avg_into(uint8_t *dst, const uint8_t *other, int other_stride, int w, int h)
{
    int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) dst[y * 16 + x] = (uint8_t)((dst[y * 16 + x] + other[y * other_stride + x] + 1) >> 1);
}

/* src = integer sample position, (dx,dy) quarter offsets 0..3, dst stride 16 (H:2079).
 * Sample naming follows ITU-T H.264 figure 8-4: b,h = horizontal/vertical half samples,
 * j = centre; quarter samples are averages of the two nearest integer/half samples. */
void orc_qpel_luma(const uint8_t *src, int stride, uint8_t *dst, int w, int h, int dx, int dy)
{
    uint8_t tmp[256];
    int x, y;
    int need_b = (dy == 0 && dx) || (dy != 2 && dx != 0 && !(dx == 2 && dy == 2)) ;
    (void)need_b;
    if (!dx && !dy)
    {
        for (y = 0; y < h; y++) for (x = 0; x < w; x++) dst[y * 16 + x] = src[y * stride + x];
        return;
    }
    if (dy == 0)
    {   /* a, b, c */
        half_hor(src, stride, dst, w, h);
        if (dx != 2) avg_into(dst, src + (dx >> 1), stride, w, h);
        return;
    }
    if (dx == 0)
    {   /* d, h, n */
        half_ver(src, stride, dst, w, h);
        if (dy != 2) avg_into(dst, src + (dy >> 1) * stride, stride, w, h);
        return;
    }
    if (dx == 2 && dy == 2) { half_diag(src, stride, dst, w, h); return; }        /* j */
    if (dx == 2)
    {   /* f, q: average of j and the nearer horizontal half sample row */
        half_hor(src + (dy >> 1) * stride, stride, dst, w, h);
        half_diag(src, stride, tmp, w, h);
        avg_into(dst, tmp, 16, w, h);
        return;
    }
    if (dy == 2)
    {   /* i, k: average of j and the nearer vertical half sample column */
        half_ver(src + (dx >> 1), stride, dst, w, h);
        half_diag(src, stride, tmp, w, h);
        avg_into(dst, tmp, 16, w, h);
        return;
    }
    /* e, g, p, r: average of the nearest horizontal and vertical half samples */
    half_hor(src + (dy >> 1) * stride, stride, dst, w, h);
    half_ver(src + (dx >> 1), stride, tmp, w, h);
    avg_into(dst, tmp, 16, w, h);
}

/* ---------------- a3: chroma eighth-sample bilinear (H:2133) -------------------- */
void orc_qpel_chroma(const uint8_t *src, int stride, uint8_t *dst, int w, int h, int dx, int dy)
{
    int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++)
        {
            const uint8_t *p = src + y * stride + x;
            if (dx | dy)
                dst[y * 16 + x] = (uint8_t)(((8 - dx) * (8 - dy) * p[0] + dx * (8 - dy) * p[1] +
                                             (8 - dx) * dy * p[stride] + dx * dy * p[stride + 1] + 32) >> 6);
            else dst[y * 16 + x] = p[0];
        }
}

/* ---------------- a7: intra 16x16 / chroma prediction (H:1625-1781) ------------- */
static int dc_of(const uint8_t *left, const uint8_t *top, int n)
{
    int i, s = 0, cnt = 0, sh = n == 16 ? 4 : (n == 8 ? 3 : 2);
    if (left) { for (i = 0; i < n; i++) s += left[i]; cnt++; }
    if (top) { for (i = 0; i < n; i++) s += top[i]; cnt++; }
    if (!cnt) return 128;
    return cnt == 2 ? (s + n) >> (sh + 1) : (s + n / 2) >> sh;
}

void orc_intra16(uint8_t *pred, const uint8_t *left, const uint8_t *top, int mode)   /* H:1677 */
{
    int x, y, dc = mode == 2 ? dc_of(left, top, 16) : 0;
    for (y = 0; y < 16; y++)
        for (x = 0; x < 16; x++) pred[y * 16 + x] = (uint8_t)(mode == 0 ? top[x] : (mode == 1 ? left[y] : dc));
}

/* both planes: pred U at +0, V at +8, stride 16; left/top hold U[8] then V[8] (H:1716).
 * mode in the luma numbering the reference passes: 0 vertical, 1 horizontal, 2 DC. */
void orc_intra_chroma(uint8_t *pred, const uint8_t *left, const uint8_t *top, int mode)
{
    int pl, x, y;
    for (pl = 0; pl < 2; pl++)
    {
        const uint8_t *l = left ? left + 8 * pl : NULL, *t = top ? top + 8 * pl : NULL;
        uint8_t *d = pred + 8 * pl;
        int dc[2][2];
        if (mode == 2)
        {   /* 8.3.4.1-3: each 4x4 quadrant has its own DC rule */
            dc[0][0] = dc_of(l, t, 4);
            dc[1][1] = dc_of(l ? l + 4 : NULL, t ? t + 4 : NULL, 4);
            dc[0][1] = t ? dc_of(NULL, t + 4, 4) : dc_of(l, NULL, 4);
            dc[1][0] = l ? dc_of(l + 4, NULL, 4) : dc_of(NULL, t, 4);
        }
        for (y = 0; y < 8; y++)
            for (x = 0; x < 8; x++)
                d[y * 16 + x] = (uint8_t)(mode == 0 ? t[x] : (mode == 1 ? l[y] : dc[y >> 2][x >> 2]));
    }
}

int orc_intra16_estimate(const uint8_t *p, int avail, int qp)      /* H:4838 */
{
    int gx = iabs_(p[0] - p[15]) + iabs_(p[240] - p[255]) + iabs_(p[128] - p[143]);
    int gy = iabs_(p[0] - p[240]) + iabs_(p[15] - p[255]) + iabs_(p[8] - p[248]);
    if (gx > 30 + 3 * gy && gy < 150 - qp && (avail & 1)) return 0;
    if (gy > 30 + 3 * gx && gx < 150 - qp && (avail & 2)) return 1;
    return 2;
}

/* ---------------- a8: the nine 4x4 predictions + mode choice (H:1810-1962) ------ */
/* edge[-5..-2] = L3..L0, edge[-1] = UL, edge[0..7] = U0..U7 (reference layout H:1163-1175) */
static void pred4(int mode, const uint8_t *edge, int avail, uint8_t o[16])
{
    int x, y;
#define T(i) ((int)edge[(i)])              /* p[i,-1], i = -1..7 */
#define L(i) ((int)edge[-2 - (i)])         /* p[-1,i], i = 0..3; L(-1) = corner */
    for (y = 0; y < 4; y++)
        for (x = 0; x < 4; x++)
        {
            int v = 0, z, k;
            switch (mode)
            {
            case 0: v = T(x); break;
            case 1: v = L(y); break;
            case 2:
                if ((avail & 3) == 3) v = (T(0) + T(1) + T(2) + T(3) + L(0) + L(1) + L(2) + L(3) + 4) >> 3;
                else if (avail & 1) v = (T(0) + T(1) + T(2) + T(3) + 2) >> 2;
                else if (avail & 2) v = (L(0) + L(1) + L(2) + L(3) + 2) >> 2;
                else v = 128;
                break;
            case 3: v = (x + y == 6) ? (T(6) + 3 * T(7) + 2) >> 2 : (T(x + y) + 2 * T(x + y + 1) + T(x + y + 2) + 2) >> 2; break;
            case 4:
                if (x > y) v = (T(x - y - 2) + 2 * T(x - y - 1) + T(x - y) + 2) >> 2;
                else if (x < y) v = (L(y - x - 2) + 2 * L(y - x - 1) + L(y - x) + 2) >> 2;
                else v = (T(0) + 2 * T(-1) + L(0) + 2) >> 2;
                break;
            case 5:
                z = 2 * x - y; k = x - (y >> 1);
                if (z >= 0 && !(z & 1)) v = (T(k - 1) + T(k) + 1) >> 1;
                else if (z >= 0) v = (T(k - 2) + 2 * T(k - 1) + T(k) + 2) >> 2;
                else if (z == -1) v = (L(0) + 2 * T(-1) + T(0) + 2) >> 2;
                else v = (L(y - 1) + 2 * L(y - 2) + L(y - 3) + 2) >> 2;
                break;
            case 6:
                z = 2 * y - x; k = y - (x >> 1);
                if (z >= 0 && !(z & 1)) v = (L(k - 1) + L(k) + 1) >> 1;
                else if (z >= 0) v = (L(k - 2) + 2 * L(k - 1) + L(k) + 2) >> 2;
                else if (z == -1) v = (L(0) + 2 * T(-1) + T(0) + 2) >> 2;
                else v = (T(x - 1) + 2 * T(x - 2) + T(x - 3) + 2) >> 2;
                break;
            case 7:
                k = x + (y >> 1);
                v = (y & 1) ? (T(k) + 2 * T(k + 1) + T(k + 2) + 2) >> 2 : (T(k) + T(k + 1) + 1) >> 1;
                break;
            case 8:
                z = x + 2 * y; k = y + (x >> 1);
                if (z > 5) v = L(3);
                else if (z == 5) v = (L(2) + 3 * L(3) + 2) >> 2;
                else if (z & 1) v = (L(k) + 2 * L(k + 1) + L(k + 2) + 2) >> 2;
                else v = (L(k) + L(k + 1) + 1) >> 1;
                break;
            }
            o[y * 4 + x] = (uint8_t)v;
        }
#undef T
#undef L
}

/* returns mode + (cost << 4); writes the chosen prediction to blockpred (stride 16).
 * Evaluation order and tie-breaking of the reference: DC, then V, DDL, VL (top available),
 * H, HU (left available), DDR, HD, VR (top+left+corner), strict "<" (H:1834-1960). */
int orc_intra4_choose(const uint8_t *blockin, uint8_t *blockpred, int avail, const uint8_t *edge_in, int mpred, int penalty)
{
    static const int order[9] = {2, 0, 3, 7, 1, 8, 4, 6, 5};
    uint8_t buf[16], *edge = buf + 5, p[16];
    int k, i, best = 0x7fffffff, best_mode = 2;
    memcpy(buf, edge_in - 5, 13);
    if ((avail & 1) && !(avail & 8)) edge[4] = edge[5] = edge[6] = edge[7] = edge[3];
    for (k = 0; k < 9; k++)
    {
        int m = order[k], ok, sad = 0;
        if (k == 0) ok = 1;
        else if (k < 4) ok = avail & 1;
        else if (k < 6) ok = avail & 2;
        else ok = (avail & 7) == 7;
        if (!ok) continue;
        pred4(m, edge, avail, p);
        for (i = 0; i < 16; i++) sad += iabs_((int)blockin[(i >> 2) * 16 + (i & 3)] - p[i]);
        if (m != mpred) sad += penalty;
        if (sad < best)
        {
            best = sad; best_mode = m;
            for (i = 0; i < 16; i++) blockpred[(i >> 2) * 16 + (i & 3)] = p[i];
        }
    }
    return best_mode + (best << 4);
}

/* ---------------- a9-a11: transform, quantisation, reconstruction ---------------- */
static void core4(int a, int b, int c, int d, int o[4])     /* forward 4-point kernel (H:2374) */
{
    int s = a + d, t = a - d, u = b + c, w = b - c;
    o[0] = s + u; o[1] = 2 * t + w; o[2] = s - u; o[3] = t - 2 * w;
}

/* out[v + 4u]: v vertical, u horizontal frequency (H:2385, TRANSPOSE_BLOCK) */
void orc_fwd4x4(const uint8_t *inp, int inp_stride, const uint8_t *pred, int16_t *out)
{
    int col[4][4], o[4], x, v;
    for (x = 0; x < 4; x++)
    {
        core4(inp[x] - pred[x], inp[inp_stride + x] - pred[16 + x], inp[2 * inp_stride + x] - pred[32 + x],
              inp[3 * inp_stride + x] - pred[48 + x], o);
        for (v = 0; v < 4; v++) col[x][v] = o[v];
    }
    for (v = 0; v < 4; v++)
    {
        core4(col[0][v], col[1][v], col[2][v], col[3][v], o);
        for (x = 0; x < 4; x++) out[v + 4 * x] = (int16_t)o[x];
    }
}

static int pos_class(int i) { return (i & 1) + ((i >> 2) & 1); }    /* H:2366: 0,1,2 -> qdat pairs 0,2,4 */

/* quantise + dequantise coefficients i0..15 of one block in place (H:2567-2585) */
int orc_quant4x4(int16_t *dq, int16_t *qv, int i0, const uint16_t *qdat)
{
    int i, any = 0;
    for (i = i0; i < 16; i++)
    {
        int c = dq[i], cl = 2 * pos_class(i);
        int rnd = c < 0 ? 0xFFFF - qdat[6] : qdat[6];
        int v = (c * qdat[cl] + rnd) >> 16;
        qv[i] = (int16_t)v;
        dq[i] = (int16_t)(v * qdat[cl + 1]);
        any |= v;
    }
    return any != 0;
}

static int all_small(const int16_t *c, int i0, const uint16_t *thr)         /* H:2491 */
{
    int i;
    for (i = i0; i < 16; i++) if ((unsigned)(c[i] + thr[i & 7]) > 2u * thr[i & 7]) return 0;
    return 1;
}

void orc_inv4x4_add(const int16_t *dq, const uint8_t *pred, uint8_t *out, int out_stride)   /* H:2436, H:2661-2670 */
{
    int t[4][4], x, v;
    for (v = 0; v < 4; v++)
    {
        int a = dq[v], b = dq[v + 4], c = dq[v + 8], d = dq[v + 12];
        int e0 = a + c, e1 = a - c, e2 = (b >> 1) - d, e3 = b + (d >> 1);
        t[v][0] = (int16_t)(e0 + e3); t[v][1] = (int16_t)(e1 + e2); t[v][2] = (int16_t)(e1 - e2); t[v][3] = (int16_t)(e0 - e3);
    }
    for (x = 0; x < 4; x++)
    {
        int a = t[0][x], b = t[1][x], c = t[2][x], d = t[3][x];
        int e0 = a + c, e1 = a - c, e2 = (b >> 1) - d, e3 = b + (d >> 1);
        int r[4], y;
        r[0] = (int16_t)((e0 + e3 + 32) >> 6); r[1] = (int16_t)((e1 + e2 + 32) >> 6);
        r[2] = (int16_t)((e1 - e2 + 32) >> 6); r[3] = (int16_t)((e0 - e3 + 32) >> 6);
        for (y = 0; y < 4; y++) out[y * out_stride + x] = (uint8_t)clip255(r[y] + pred[y * 16 + x]);
    }
}

static void hadamard4(const int in[4], int o[4])
{
    int s = in[0] + in[2], t = in[0] - in[2], u = in[1] + in[3], w = in[1] - in[3];
    o[0] = s + u; o[1] = t + w; o[2] = t - w; o[3] = s - u;
}
static void hadamard4x4(int16_t *x)          /* H:2269: columns into rows, then again */
{
    int16_t tmp[16];
    int i, k, in[4], o[4];
    for (i = 0; i < 4; i++)
    {
        for (k = 0; k < 4; k++) in[k] = x[i + 4 * k];
        hadamard4(in, o);
        for (k = 0; k < 4; k++) tmp[4 * i + k] = (int16_t)o[k];
    }
    for (i = 0; i < 4; i++)
    {
        for (k = 0; k < 4; k++) in[k] = tmp[i + 4 * k];
        hadamard4(in, o);
        for (k = 0; k < 4; k++) x[i + 4 * k] = (int16_t)o[k];
    }
}

/*
 * One component exactly as mb_write drives it (H:4423-4490): mode 2 = one intra 4x4 block,
 * 8 = inter luma (16 blocks, small-coefficient zeroing), 9 = intra 16x16 luma (DC split off),
 * 5 = chroma plane (4 blocks, DC split off).  q_out holds n*n records of qv[16], dq[16];
 * dc_out the quantised DC levels.  Returns the block mask (| dc_flag << 8 for chroma).
 */
int orc_transform_quant(const uint8_t *inp, const uint8_t *pred, int inp_stride, int mode,
                        int16_t *q_out, int16_t *dc_out, const uint16_t *qdat)
{
    int n = mode >> 1, i0 = mode & 1, nb = n * n, b, mask = 0, zmask = 0;
    int16_t dq[16][16], qv[16][16], dc[16];
    memset(qv, 0, sizeof(qv));
    for (b = 0; b < nb; b++)
    {
        int bx = b % n, by = b / n;
        orc_fwd4x4(inp + 4 * bx + 4 * by * inp_stride, inp_stride, pred + 4 * bx + 64 * by, dq[b]);
        if (i0) dc[b] = dq[b][0];
    }
    if (mode == 8 || mode == 5)
    {   /* H:2512: zero blocks whose coefficients are all below thr1; for inter also whole
           8x8 groups below thr2 */
        for (b = 0; b < nb; b++) if (all_small(dq[b], i0, qdat + 10)) zmask |= 1 << b;
        if (mode == 8)
        {
            static const int g0[4] = {0, 2, 8, 10};
            int g;
            for (g = 0; g < 4; g++)
            {
                int m = 0x33 << g0[g], f = g0[g];
                if ((~zmask & m) && all_small(dq[f], i0, qdat + 18) && all_small(dq[f + 1], i0, qdat + 18) &&
                    all_small(dq[f + 4], i0, qdat + 18) && all_small(dq[f + 5], i0, qdat + 18)) zmask |= m;
            }
        }
    }
    for (b = 0; b < nb; b++)
    {
        int nz = 0;
        if (!(zmask & (1 << b))) nz = orc_quant4x4(dq[b], qv[b], i0, qdat);
        else memset(qv[b], 0, sizeof(qv[b]));
        mask = (mask << 1) | nz;
    }
    if (mode == 9)
    {   /* H:2344: 4x4 Hadamard, quantise with half-step rounding, Hadamard, scale */
        int i;
        hadamard4x4(dc);
        for (i = 0; i < 16; i++)
        {
            int v = dc[i];
            v = (v * (int16_t)qdat[0] + (v < 0 ? (1 << 18) - 0x20000 : 0x20000)) >> 18;
            dc_out[i] = dc[i] = (int16_t)v;
        }
        hadamard4x4(dc);
        for (i = 0; i < 16; i++) dq[i][0] = (int16_t)(dc[i] * (int16_t)(qdat[1] >> 2));
    }
    if (mode == 5)
    {   /* H:2355: 2x2 Hadamard, third-step rounding */
        int i, a = dc[0], bb = dc[1], c = dc[2], d = dc[3], t[4], flag;
        t[0] = a + bb + c + d; t[1] = a - bb + c - d; t[2] = a + bb - c - d; t[3] = a - bb - c + d;
        for (i = 0; i < 4; i++)
        {
            int v = (int16_t)t[i];
            v = (v * (int16_t)(qdat[0] << 1) + (v < 0 ? (1 << 18) - 0xAAAA : 0xAAAA)) >> 18;
            dc_out[i] = (int16_t)v; t[i] = (int16_t)v;
        }
        a = t[0]; bb = t[1]; c = t[2]; d = t[3];
        t[0] = (int16_t)(a + bb + c + d); t[1] = (int16_t)(a - bb + c - d); t[2] = (int16_t)(a + bb - c - d); t[3] = (int16_t)(a - bb - c + d);
        flag = (t[0] | t[1] | t[2] | t[3]) != 0;
        for (i = 0; i < 4; i++) dq[i][0] = (int16_t)(t[i] * (int16_t)(qdat[1] >> 1));
        mask |= flag << 8;
    }
    for (b = 0; b < nb; b++)
    {
        memcpy(q_out + b * 32, qv[b], 32);
        memcpy(q_out + b * 32 + 16, dq[b], 32);
    }
    return mask;
}

/* ---------------- a13: CAVLC residual block (H:2775-2949) ------------------------ */
#define H264_TAB static const
#include "../h264-lab_b200/csrc/h264_cavlc_tables.h"     /* ITU-T H.264 tables 9-5, 9-7..9-10 (data only) */

typedef struct { uint8_t *buf; int bits; } obits_t;
static void ob_put(obits_t *b, int n, unsigned v)
{
    int i;
    for (i = n - 1; i >= 0; i--)
    {
        if ((v >> i) & 1) b->buf[b->bits >> 3] |= (uint8_t)(0x80 >> (b->bits & 7));
        b->bits++;
    }
}

/* coefficients coded in plain array order (no zig-zag: the reference's quirk, SURVEY
 * Appendix B.1).  c = first coded coefficient, n = 4, 15 or 16.  Returns bits written. */
int orc_cavlc_block(const int16_t *c, int n, int nA, int nB, uint8_t *out, int *total_coeff)
{
    obits_t b = {out, 0};
    int level[16], pos[16], total = 0, t1 = 0, i, ctx, tab, sl;
    for (i = n - 1; i >= 0; i--) if (c[i]) { level[total] = c[i]; pos[total] = i; total++; }
    while (t1 < total && t1 < 3 && iabs_(level[t1]) == 1) t1++;
    if (total_coeff) *total_coeff = total;
    ctx = nA + nB;
    if (ctx <= 34) ctx = (ctx + 1) >> 1;         /* both neighbours available: average (H:2816-2823) */
    ctx &= 31;                                   /* one neighbour = 64 (not available): the other one */
    tab = ctx < 2 ? 0 : ctx < 4 ? 1 : ctx < 8 ? 2 : ctx < 17 ? 3 : 4;
    ob_put(&b, cavlc_coeff_token_len[tab][total * 4 + t1], cavlc_coeff_token_code[tab][total * 4 + t1]);
    if (!total) return b.bits;
    for (i = 0; i < t1; i++) ob_put(&b, 1, level[i] < 0);
    sl = (total > 10 && t1 < 3) ? 1 : 0;
    for (i = t1; i < total; i++)
    {   /* 9.2.2.1 level_prefix / level_suffix */
        int code = level[i] > 0 ? 2 * level[i] - 2 : -2 * level[i] - 1;
        int prefix, nsuf, suf;
        if (i == t1 && t1 < 3) code -= 2;
        if (sl == 0 && code < 14) { prefix = code; nsuf = 0; suf = 0; }
        else if (sl == 0 && code < 30) { prefix = 14; nsuf = 4; suf = code - 14; }
        else if (sl == 0) { prefix = 15; nsuf = 12; suf = code - 30; }
        else if ((code >> sl) < 15) { prefix = code >> sl; nsuf = sl; suf = code & ((1 << sl) - 1); }
        else { prefix = 15; nsuf = 12; suf = code - (15 << sl); }
        ob_put(&b, prefix + 1, 1);
        if (nsuf) ob_put(&b, nsuf, (unsigned)suf);
        if (sl == 0) sl = 1;
        if (iabs_(level[i]) > (3 << (sl - 1)) && sl < 6) sl++;
    }
    if (total < n)
    {
        int zeros = pos[0] + 1 - total, left = zeros;
        if (n == 4) ob_put(&b, cavlc_total_zeros_dc_len[total - 1][zeros], cavlc_total_zeros_dc_code[total - 1][zeros]);
        else ob_put(&b, cavlc_total_zeros_len[total - 1][zeros], cavlc_total_zeros_code[total - 1][zeros]);
        for (i = 0; i + 1 < total && left > 0; i++)
        {
            int run = pos[i] - pos[i + 1] - 1, t = (left > 7 ? 7 : left) - 1;
            ob_put(&b, cavlc_run_before_len[t][run], cavlc_run_before_code[t][run]);
            left -= run;
        }
    }
    return b.bits;
}

/* ---------------- a15: deblocking of one macroblock (H:1191-1545) ---------------- */
static int clipr(int r, int v) { return v > r ? r : (v < -r ? -r : v); }

static void luma_edge_sample(uint8_t *p, int step, int bs, int alpha, int beta, int tc0)
{
    int p2 = p[-3 * step], p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step], q2 = p[2 * step];
    int ap = iabs_(p2 - p0), aq = iabs_(q2 - q0);
    if (!(iabs_(p0 - q0) < alpha && iabs_(p1 - p0) < beta && iabs_(q1 - q0) < beta)) return;
    if (bs < 4)
    {   /* 8.7.2.3 */
        int tc = tc0 + (ap < beta) + (aq < beta);
        int d = clipr(tc, ((q0 - p0) * 4 + (p1 - q1) + 4) >> 3);
        if (ap < beta) p[-2 * step] = (uint8_t)(p1 + clipr(tc0, ((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1));
        if (aq < beta) p[step] = (uint8_t)(q1 + clipr(tc0, ((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1));
        p[-step] = (uint8_t)clip255(p0 + d);
        p[0] = (uint8_t)clip255(q0 - d);
    } else
    {   /* 8.7.2.4 */
        int strong = iabs_(p0 - q0) < (alpha >> 2) + 2;
        if (strong && ap < beta)
        {
            int p3 = p[-4 * step];
            p[-step] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            p[-2 * step] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
            p[-3 * step] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else p[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        if (strong && aq < beta)
        {
            int q3 = p[3 * step];
            p[0] = (uint8_t)((q2 + 2 * q1 + 2 * q0 + 2 * p0 + p1 + 4) >> 3);
            p[step] = (uint8_t)((q2 + q1 + p0 + q0 + 2) >> 2);
            p[2 * step] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else p[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}

/* strength[32]: [4*e + s] vertical edge e, segment s; [16 + 4*e + s] horizontal (H:613).
 * alpha/beta [0] left edge, [1] inner vertical, [2] top edge, [3] inner horizontal; tc0 per strength slot. */
void orc_deblock_luma(uint8_t *pix, int stride, const uint8_t *strength, const uint8_t *tc0,
                      const uint8_t *alpha, const uint8_t *beta)
{
    int e, i;
    for (e = 0; e < 4; e++)
        for (i = 0; i < 16; i++)
        {
            int k = 4 * e + (i >> 2), a = alpha[e ? 1 : 0], b = beta[e ? 1 : 0];
            int bs = strength[4 * e] == 4 ? 4 : strength[k];
            if (bs && (bs == 4 || a)) luma_edge_sample(pix + i * stride + 4 * e, 1, bs, a, b, tc0[k]);
        }
    for (e = 0; e < 4; e++)
        for (i = 0; i < 16; i++)
        {
            int k = 16 + 4 * e + (i >> 2), a = alpha[e ? 3 : 2], b = beta[e ? 3 : 2];
            int bs = strength[16 + 4 * e] == 4 ? 4 : strength[k];
            if (bs && (bs == 4 || a)) luma_edge_sample(pix + (4 * e) * stride + i, stride, bs, a, b, tc0[k]);
        }
}

static void chroma_edge_sample(uint8_t *p, int step, int bs, int alpha, int beta, int tc0)    /* H:1217 */
{
    int p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step];
    if (!bs || iabs_(p0 - q0) >= alpha || iabs_(p1 - p0) >= beta || iabs_(q1 - q0) >= beta) return;
    if (bs < 4)
    {
        int d = clipr(tc0 + 1, ((q0 - p0) * 4 + (p1 - q1) + 4) >> 3);
        p[-step] = (uint8_t)clip255(p0 + d);
        p[0] = (uint8_t)clip255(q0 - d);
    } else
    {
        p[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        p[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}

void orc_deblock_chroma(uint8_t *pix, int stride, const uint8_t *strength, const uint8_t *tc0,
                        const uint8_t *alpha, const uint8_t *beta)           /* H:1469 */
{
    int e, i;
    for (e = 0; e < 2; e++)
    {
        int a = alpha[e ? 1 : 0], b = beta[e ? 1 : 0], any = 0;
        for (i = 0; i < 4; i++) any |= strength[8 * e + i];
        if (!any || !a) continue;
        for (i = 0; i < 8; i++) chroma_edge_sample(pix + i * stride + 4 * e, 1, strength[8 * e + (i >> 1)], a, b, tc0[8 * e + (i >> 1)]);
    }
    for (e = 0; e < 2; e++)
    {
        int a = alpha[e ? 3 : 2], b = beta[e ? 3 : 2], any = 0;
        for (i = 0; i < 4; i++) any |= strength[16 + 8 * e + i];
        if (!any || !a) continue;
        for (i = 0; i < 8; i++) chroma_edge_sample(pix + (4 * e) * stride + i, stride, strength[16 + 8 * e + (i >> 1)], a, b, tc0[16 + 8 * e + (i >> 1)]);
    }
}

/* ---------------- a16: guard band replication (H:2232) --------------------------- */
void orc_extend_borders(uint8_t *pic, int w, int h, int guard)
{
    int stride = w + 2 * guard, x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < guard; x++)
        {
            pic[y * stride - 1 - x] = pic[y * stride];
            pic[y * stride + w + x] = pic[y * stride + w - 1];
        }
    for (y = 0; y < guard; y++)
    {
        memcpy(pic - guard + (-1 - y) * stride, pic - guard, (size_t)stride);
        memcpy(pic - guard + (h + y) * stride, pic - guard + (h - 1) * stride, (size_t)stride);
    }
}

/* ---------------- temporal noise suppressor (h264e_denoise_run H:1547-1620) -------- */
#include "../h264-lab_b200/csrc/h264_denoise_tab.h"       /* 255 - gain(d), 255 - min(4 gain(d), 255): data only (H:1122) */
/* In place like the reference: `prev` (the previous output, w x h at prev_stride) becomes the new output.
 * The reference parks every output row one row up while it works and shifts the rows back afterwards
 * (H:1561-1619); with a scratch copy of the old picture that is simply: border samples = current picture,
 * inner samples blended by the weight of the sample's own and its 4-neighbourhood's temporal difference. */
void orc_denoise_run(const uint8_t *cur, uint8_t *prev, int w, int h, int cur_stride, int prev_stride)
{
    int x, y;
    uint8_t *old;
    if (w <= 2 || h <= 2) return;                                         /* H:1550 */
    old = (uint8_t *)malloc((size_t)w * h);
    for (y = 0; y < h; y++) memcpy(old + (size_t)y * w, prev + (size_t)y * prev_stride, (size_t)w);
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++)
        {
            const uint8_t *c = cur + (size_t)y * cur_stride + x;
            const uint8_t *p = old + (size_t)y * w + x;
            int v = c[0];
            if (x > 0 && y > 0 && x < w - 1 && y < h - 1)
            {
                int d = iabs_(c[0] - p[0]);                                                         /* H:1572, 1578 */
                int nb = iabs_((c[-1] - p[-1]) + (c[1] - p[1]) + (c[-cur_stride] - p[-w]) + (c[cur_stride] - p[w])) >> 2;   /* H:1573-1582 */
                unsigned g = (unsigned)denoise_weight[d][0] * (unsigned)denoise_weight[nb][1];       /* H:1584-1594 */
                v = (int)(((unsigned)p[0] * g + (0xffffu - g) * (unsigned)c[0] + (1u << 15)) >> 16); /* H:1598 */
            }
            prev[(size_t)y * prev_stride + x] = (uint8_t)v;
        }
    free(old);
}
