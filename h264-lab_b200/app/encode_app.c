/*
 * encode_app -- command-line encoder on top of libh264lab_b200.so, with the command
 * line of the reference's test application (/root/reference/src/minih264e_test.c,
 * cited as T:nnn): --input/-i --output/-o --recon/-r --gen --gop --qp --kbps
 * --maxframes --threads --speed --denoise --stats --psnr.
 *
 * Kept quirks (they change nothing in the bit stream): every --long option consumes the
 * following argument (T:212); -r sets the INPUT name (T:215); --maxframes only
 * distinguishes zero from non-zero (T:576); --recon and --threads are accepted and
 * ignored; --gen encodes 301 frames of 1024x768 rotating chessboard (T:435-452,
 * T:578-582); --denoise sets temporal_denoise_flag (T:525).  Difference: --segments N (extension) encodes N closed-GOP segments of the
 * input concurrently, one fresh encoder session per segment, and concatenates them --
 * the GOP-sharded mode of DESIGN.md; its output equals the reference run once per
 * segment.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "h264-lab.h"

#define DEFAULT_GOP 20       /* T:10 */
#define DEFAULT_QP 33        /* T:11 */

static struct
{
    char input_file[512], output_file[512], recon_file[512];
    int gen, gop, qp, kbps, max_frames, threads, speed, denoise, stats, psnr, segments;
} opt;

typedef struct { const char *name; int w, h; } size_name_t;
static const size_name_t k_sizes[] = {     /* T:256-281 */
    {"sqcif", 128, 96}, {"qvga", 320, 240}, {"svga", 800, 600}, {"4vga", 1280, 960}, {"sxga", 1280, 1024},
    {"xga", 1024, 768}, {"vga", 640, 480}, {"qcif", 176, 144}, {"4cif", 704, 576}, {"4sif", 704, 480},
    {"cif", 352, 288}, {"sif", 352, 240}, {"pal", 720, 576}, {"ntsc", 720, 480}, {"d1", 720, 480},
    {"16cif", 1408, 1152}, {"16sif", 1408, 960}, {"720p", 1280, 720}, {"4SVGA", 1600, 1200},
    {"4XGA", 2048, 1536}, {"16VGA", 2560, 1920}, {NULL, 0, 0}};

/* the last size specification ("352x288", "cif", ...) found in the file name (T:288-329) */
static int guess_size(const char *name, int *w, int *h)
{
    int i = (int)strlen(name), found = 0;
    while (--i >= 0)
    {
        const size_name_t *s;
        const char *p = name + i;
        int prev = found;
        found = 0;
        if (*p >= '0' && *p <= '9')
        {
            char *end;
            int ww = (int)strtoul(p, &end, 10);
            if (ww && (*end == 'x' || *end == 'X') && end[1] >= '1' && end[1] <= '9')
            {
                int hh = (int)strtoul(end + 1, &end, 10);
                if (hh) { *w = ww; *h = hh; found = 1; }
            }
        }
        for (s = k_sizes; s->name; s++)
            if (!strncmp(p, s->name, strlen(s->name))) { *w = s->w; *h = s->h; found = 1; }
        if (!found && prev) return prev;
    }
    return found;
}

static int long_option(const char *p, const char *val)
{
    if (!val) val = "";
    if (!strncmp(p, "gen", 3)) opt.gen = 1;
    else if (!strncmp(p, "gop", 3)) opt.gop = atoi(val);
    else if (!strncmp(p, "qp", 2)) opt.qp = atoi(val);
    else if (!strncmp(p, "kbps", 4)) opt.kbps = atoi(val);
    else if (!strncmp(p, "maxframes", 9)) opt.max_frames = atoi(val);
    else if (!strncmp(p, "threads", 7)) opt.threads = atoi(val);
    else if (!strncmp(p, "speed", 5)) opt.speed = atoi(val);
    else if (!strncmp(p, "segments", 8)) opt.segments = atoi(val);
    else if (!strncmp(p, "denoise", 7)) opt.denoise = 1;
    else if (!strncmp(p, "stats", 5)) opt.stats = 1;
    else if (!strncmp(p, "psnr", 4)) opt.psnr = 1;
    else if (!strncmp(p, "output", 6)) snprintf(opt.output_file, sizeof(opt.output_file), "%s", val);
    else if (!strncmp(p, "input", 5)) snprintf(opt.input_file, sizeof(opt.input_file), "%s", val);
    else if (!strncmp(p, "recon", 5)) snprintf(opt.recon_file, sizeof(opt.recon_file), "%s", val);
    else { printf("ERROR: Unknown option %s\n", p); return 0; }
    return 1;
}

static int parse_args(int argc, char **argv)
{
    int i;
    opt.gop = DEFAULT_GOP; opt.qp = DEFAULT_QP; opt.max_frames = 99999;
    for (i = 1; i < argc; i++)
    {
        const char *p = argv[i];
        if (*p++ != '-') { printf("ERROR: Unknown option %s\n", p); return 0; }
        switch (*p)
        {
        case '-': long_option(p + 1, i + 1 < argc ? argv[i + 1] : NULL); i++; break;
        case 'o': if (++i < argc) snprintf(opt.output_file, sizeof(opt.output_file), "%s", argv[i]); break;
        case 'i':
        case 'r': if (++i < argc) snprintf(opt.input_file, sizeof(opt.input_file), "%s", argv[i]); break;
        default: break;
        }
    }
    return 1;
}

/* the reference's synthetic source (T:407-452) */
static int chess_pixel(double x, double y)
{
    int mid = (fabs(x) < 4 && fabs(y) < 4);
    int i = (int)x, j = (int)y;
    int black = mid ? 128 : i / 16, white = mid ? 128 : 255 - j / 16;
    int c00 = (((i >> 4) + (j >> 4)) & 1) ? white : black;
    int c01 = ((((i + 1) >> 4) + (j >> 4)) & 1) ? white : black;
    int c10 = (((i >> 4) + ((j + 1) >> 4)) & 1) ? white : black;
    int c11 = ((((i + 1) >> 4) + ((j + 1) >> 4)) & 1) ? white : black;
    int s = (int)((c00 * (1 - (x - i)) + c01 * (x - i)) * (1 - (y - j)) + (c10 * (1 - (x - i)) + c11 * (x - i)) * (y - j) + 0.5);
    return s < 0 ? 0 : s > 255 ? 255 : s;
}
static void gen_frame(unsigned char *p, int w, int h, int frm)
{
    double co = cos(.01 * frm), si = sin(.01 * frm);
    int r, c, hw = w >> 1, hh = h >> 1;
    for (r = 0; r < h; r++)
        for (c = 0; c < w; c++)
            p[r * w + c] = (unsigned char)chess_pixel(co * (c - hw) + si * (r - hh), -si * (c - hw) + co * (r - hh));
    memset(p + w * h, 128, (size_t)w * h / 2);
}

static double g_noise[3], g_count[3], g_bytes;
static int g_frames;
static void psnr_add(const unsigned char *a, const unsigned char *b, int w, int h, int bytes)
{
    int k, i;
    for (k = 0; k < 3; k++)
    {
        double s = 0;
        for (i = 0; i < w * h; i++) { int d = *a++ - *b++; s += d * d; }
        g_count[k] += w * h; g_noise[k] += s;
        if (!k) { w >>= 1; h >>= 1; }
    }
    g_frames++; g_bytes += bytes;
}
static void psnr_print(void)
{
    static const char *nm[3] = {"YPSNR", "UPSNR", "VPSNR"};
    double kbps = g_bytes * 8. / ((double)g_frames / 30) / 1000;
    double db = 10 * log10(255. * 255 / (g_noise[0] / g_count[0]));
    int i;
    printf("%5.0f kbps@30fps  ", kbps);
    for (i = 0; i < 3; i++) printf(" %s=%.2f db ", nm[i], 10 * log10(255. * 255 / (g_noise[i] / g_count[i])));
    printf("  %6.2f db/rate ", 10 * log10(g_count[0] * g_count[0] * 3 / 2 * 255 * 255 / (g_noise[0] * g_bytes)));
    printf("  %6.3f db/lgrate   \n", db / log10(kbps));
}

static void fill_run_param(H264E_run_param_t *rp)
{
    memset(rp, 0, sizeof(*rp));
    rp->encode_speed = opt.speed;
    if (opt.kbps) { rp->desired_frame_bytes = opt.kbps * 1000 / 8 / 30; rp->qp_min = 10; rp->qp_max = 50; }
    else rp->qp_min = rp->qp_max = opt.qp;
}

static void set_yuv(H264E_io_yuv_t *y, unsigned char *buf, int w, int h)
{
    y->yuv[0] = buf; y->stride[0] = w;
    y->yuv[1] = buf + w * h; y->stride[1] = w / 2;
    y->yuv[2] = buf + w * h * 5 / 4; y->stride[2] = w / 2;
}

/* extension: closed-GOP segments encoded concurrently, one session per segment */
static int run_segments(FILE *fin, FILE *fout, int w, int h, const H264E_create_param_t *cp)
{
    size_t frame_size = (size_t)w * h * 3 / 2;
    long nframes;
    int nseg = opt.segments, seglen, s, t, sp = 0, ss = 0, err;
    unsigned char *clip, **outbuf;
    size_t *outlen, *outcap;
    H264E_persist_t **enc;
    H264E_scratch_t **scr;
    H264E_run_param_t rp, **rps;
    H264E_io_yuv_t *yuv, **yuvs;
    unsigned char **coded;
    int *ncoded;
    fseek(fin, 0, SEEK_END);
    nframes = ftell(fin) / (long)frame_size;
    fseek(fin, 0, SEEK_SET);
    if (nframes <= 0) return 1;
    if (nseg > nframes) nseg = (int)nframes;
    seglen = (int)((nframes + nseg - 1) / nseg);
    clip = (unsigned char *)malloc(frame_size * (size_t)nframes);
    if (!clip || fread(clip, frame_size, (size_t)nframes, fin) != (size_t)nframes) return 1;
    err = H264E_sizeof(cp, &sp, &ss);
    if (err) { printf("H264E_init error = %d\n", err); return 1; }
    enc = calloc(nseg, sizeof(*enc)); scr = calloc(nseg, sizeof(*scr)); rps = calloc(nseg, sizeof(*rps));
    yuv = calloc(nseg, sizeof(*yuv)); yuvs = calloc(nseg, sizeof(*yuvs)); coded = calloc(nseg, sizeof(*coded));
    ncoded = calloc(nseg, sizeof(*ncoded)); outbuf = calloc(nseg, sizeof(*outbuf));
    outlen = calloc(nseg, sizeof(*outlen)); outcap = calloc(nseg, sizeof(*outcap));
    fill_run_param(&rp);
    for (s = 0; s < nseg; s++)
    {
        enc[s] = (H264E_persist_t *)aligned_alloc(64, ((size_t)sp + 63) & ~(size_t)63);
        scr[s] = (H264E_scratch_t *)aligned_alloc(64, ((size_t)ss + 63) & ~(size_t)63);
        err = H264E_init(enc[s], cp);
        if (err) { printf("H264E_init error = %d\n", err); return 1; }
        rps[s] = &rp;
        yuvs[s] = &yuv[s];
    }
    for (t = 0; t < seglen; t++)
    {
        int n = 0;
        int idx[4096];
        for (s = 0; s < nseg && s < 4096; s++)
        {
            long f = (long)s * seglen + t;
            if (f >= nframes || f >= (long)(s + 1) * seglen) continue;
            set_yuv(&yuv[n], clip + frame_size * (size_t)f, w, h);
            yuvs[n] = &yuv[n];
            idx[n] = s;
            n++;
        }
        if (!n) break;
        {
            H264E_persist_t *e2[4096]; H264E_scratch_t *s2[4096];
            int k;
            for (k = 0; k < n; k++) { e2[k] = enc[idx[k]]; s2[k] = scr[idx[k]]; }
            err = H264E_encode_batch(n, e2, s2, (const H264E_run_param_t *const *)rps, yuvs, coded, ncoded);
            if (err) { printf("H264E_encode error = %d\n", err); return 1; }
            for (k = 0; k < n; k++)
            {
                s = idx[k];
                if (outlen[s] + (size_t)ncoded[k] > outcap[s])
                {
                    outcap[s] = (outlen[s] + (size_t)ncoded[k]) * 2 + 65536;
                    outbuf[s] = (unsigned char *)realloc(outbuf[s], outcap[s]);
                }
                memcpy(outbuf[s] + outlen[s], coded[k], (size_t)ncoded[k]);
                outlen[s] += (size_t)ncoded[k];
                if (opt.stats) printf("segment=%d frame=%d, bytes=%d\n", s, t, ncoded[k]);
            }
        }
    }
    for (s = 0; s < nseg; s++)
    {
        if (fout && outlen[s]) fwrite(outbuf[s], outlen[s], 1, fout);
        H264E_close(enc[s]);
        free(enc[s]); free(scr[s]); free(outbuf[s]);
    }
    free(clip);
    return 0;
}

int main(int argc, char **argv)
{
    H264E_create_param_t cp;
    H264E_run_param_t rp;
    H264E_io_yuv_t yuv;
    FILE *fin = NULL, *fout;
    int w = 352, h = 288, i, frames = 0, sp = 0, ss = 0, err;
    size_t frame_size;
    unsigned char *buf_in, *buf_save, *coded = NULL;
    int ncoded = 0;
    H264E_persist_t *enc;
    H264E_scratch_t *scratch;
    struct timespec t0, t1;

    if (!parse_args(argc, argv)) return 1;
    if (!opt.gen)
    {
        guess_size(opt.input_file, &w, &h);
        fin = fopen(opt.input_file, "rb");
        if (!fin) { printf("ERROR: cant open input file %s\n", opt.input_file); return 1; }
    } else { w = 1024; h = 768; }
    fout = fopen(opt.output_file[0] ? opt.output_file : "out.264", "wb");
    if (!fout) { printf("ERROR: cant open output file %s\n", opt.output_file); return 1; }

    memset(&cp, 0, sizeof(cp));           /* T:507-526 */
    cp.enableNEON = 1;
    cp.num_layers = 1;
    cp.gop = opt.gop;
    cp.width = w;
    cp.height = h;
    cp.const_input_flag = opt.psnr ? 0 : 1;
    cp.temporal_denoise_flag = opt.denoise;       /* T:525 */
    cp.vbv_size_bytes = 100000 / 8;

    if (opt.segments > 1 && fin)
    {
        int rc;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        rc = run_segments(fin, fout, w, h, &cp);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if (opt.stats) printf("elapsed %.3f s\n", (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
        fclose(fin); fclose(fout);
        return rc;
    }

    frame_size = (size_t)w * h * 3 / 2;
    buf_in = (unsigned char *)aligned_alloc(64, (frame_size + 63) & ~(size_t)63);
    buf_save = (unsigned char *)aligned_alloc(64, (frame_size + 63) & ~(size_t)63);
    if (!buf_in || !buf_save) { printf("ERROR: not enough memory\n"); return 1; }

    err = H264E_sizeof(&cp, &sp, &ss);
    if (err) { printf("H264E_init error = %d\n", err); return 0; }
    printf("sizeof_persist = %d sizeof_scratch = %d\n", sp, ss);
    enc = (H264E_persist_t *)aligned_alloc(64, ((size_t)sp + 63) & ~(size_t)63);
    scratch = (H264E_scratch_t *)aligned_alloc(64, ((size_t)ss + 63) & ~(size_t)63);
    err = H264E_init(enc, &cp);
    if (err) { printf("H264E_init error = %d (100: no CUDA device; this build has no CPU path)\n", err); return 1; }

    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (i = 0; opt.max_frames; i++)
    {
        if (!fin) { if (i > 300) break; gen_frame(buf_in, w, h, i); }
        else if (!fread(buf_in, frame_size, 1, fin)) break;
        if (opt.psnr) memcpy(buf_save, buf_in, frame_size);
        set_yuv(&yuv, buf_in, w, h);
        fill_run_param(&rp);
        err = H264E_encode(enc, scratch, &rp, &yuv, &coded, &ncoded);
        if (err) { printf("ERROR: H264E_encode = %d\n", err); return 1; }
        if (opt.stats) printf("frame=%d, bytes=%d\n", frames++, ncoded);
        if (!fwrite(coded, (size_t)ncoded, 1, fout)) { printf("ERROR writing output file\n"); break; }
        if (opt.psnr) psnr_add(buf_save, buf_in, w, h, ncoded);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (opt.stats) printf("elapsed %.3f s\n", (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
    if (opt.psnr) psnr_print();
    H264E_close(enc);
    free(enc); free(scratch); free(buf_in); free(buf_save);
    if (fin) fclose(fin);
    fclose(fout);
    return 0;
}
