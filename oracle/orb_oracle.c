/* TEST INFRASTRUCTURE ONLY -- see orb_oracle.h.  CPU oracle for the ORB front end.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: float ops individually rounded,
 * matching the reference build, which sets no -march/-O flags: CMakeLists.txt:54-55). */
#include "orb_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAXL 16
#define PATCH_SIZE 31
#define HALF_PATCH 15
#define EDGE_TH 19

static const int8_t k_pattern[1024] = {
#include "orb_pattern_data.inc"
};

/* cvRound: round half to even (A.5) */
static inline int cv_round_f(float v) { return (int)nearbyintf(v); }
static inline int cv_round_d(double v) { return (int)nearbyint(v); }

struct orc_extractor {
    int nfeatures, nlevels, ini_th, min_th;
    double scale_factor; /* the member is double: include/orb_extractor.h:115 */
    float scale[MAXL], inv_scale[MAXL], sigma2[MAXL], inv_sigma2[MAXL];
    int per_level[MAXL];
    int umax[HALF_PATCH + 1];
    int lw[MAXL], lh[MAXL];
    uint8_t *level[MAXL], *blur[MAXL];
    float *cands[MAXL], *dist[MAXL];
    int ncands[MAXL], ndist[MAXL], tie_cut[MAXL];
};

/* ---- ctor tables: src/orb_extractor.cpp:410-470 ------------------------------------ */
orc_extractor *orc_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th,
                                    int min_th) {
    if (nlevels < 1 || nlevels > MAXL || nfeatures < 1) return NULL;
    orc_extractor *ex = (orc_extractor *)calloc(1, sizeof(*ex));
    ex->nfeatures = nfeatures;
    ex->nlevels = nlevels;
    ex->ini_th = ini_th;
    ex->min_th = min_th;
    ex->scale_factor = (double)scale_factor;
    ex->scale[0] = 1.0f;
    ex->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        ex->scale[i] = (float)((double)ex->scale[i - 1] * ex->scale_factor);
        ex->sigma2[i] = ex->scale[i] * ex->scale[i];
    }
    for (int i = 0; i < nlevels; i++) {
        ex->inv_scale[i] = 1.0f / ex->scale[i];
        ex->inv_sigma2[i] = 1.0f / ex->sigma2[i];
    }
    float factor = (float)(1.0 / ex->scale_factor);
    float denom = 1 - (float)pow((double)factor, (double)nlevels);
    float want = (float)nfeatures * (1 - factor) / denom;
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        ex->per_level[l] = cv_round_f(want);
        sum += ex->per_level[l];
        want *= factor;
    }
    ex->per_level[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    /* circular patch row ends */
    int vmax = (int)floorf(HALF_PATCH * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceilf(HALF_PATCH * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH * HALF_PATCH;
    for (int v = 0; v <= vmax; ++v) ex->umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (int v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (ex->umax[v0] == ex->umax[v0 + 1]) ++v0;
        ex->umax[v] = v0;
        ++v0;
    }
    return ex;
}

static void free_stages(orc_extractor *ex) {
    for (int l = 0; l < MAXL; l++) {
        free(ex->level[l]); ex->level[l] = NULL;
        free(ex->blur[l]); ex->blur[l] = NULL;
        free(ex->cands[l]); ex->cands[l] = NULL;
        free(ex->dist[l]); ex->dist[l] = NULL;
        ex->ncands[l] = ex->ndist[l] = 0;
    }
}

void orc_extractor_destroy(orc_extractor *ex) {
    if (!ex) return;
    free_stages(ex);
    free(ex);
}

void orc_extractor_tables(const orc_extractor *ex, float *scale, float *inv_scale, float *sigma2,
                          float *inv_sigma2, int *per_level, int *umax16) {
    for (int i = 0; i < ex->nlevels; i++) {
        if (scale) scale[i] = ex->scale[i];
        if (inv_scale) inv_scale[i] = ex->inv_scale[i];
        if (sigma2) sigma2[i] = ex->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = ex->inv_sigma2[i];
        if (per_level) per_level[i] = ex->per_level[i];
    }
    if (umax16)
        for (int i = 0; i <= HALF_PATCH; i++) umax16[i] = ex->umax[i];
}

/* level size: src/orb_extractor.cpp:1111-1112 */
void orc_level_size(const orc_extractor *ex, int w, int h, int level, int *lw, int *lh) {
    float s = ex->inv_scale[level];
    *lw = cv_round_f((float)w * s);
    *lh = cv_round_f((float)h * s);
}

/* ---- A.2: cv::resize(INTER_LINEAR) for CV_8UC1 ------------------------------------- */
void orc_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw,
                          int dh, int dstride) {
    int *xofs = (int *)malloc(sizeof(int) * dw);
    short *ax = (short *)malloc(sizeof(short) * 2 * dw);
    double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ax[2 * dx] = (short)cv_round_f((1.f - fx) * 2048);
        ax[2 * dx + 1] = (short)cv_round_f(fx * 2048);
    }
    int *row0 = (int *)malloc(sizeof(int) * dw), *row1 = (int *)malloc(sizeof(int) * dw);
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        short b0 = (short)cv_round_f((1.f - fy) * 2048), b1 = (short)cv_round_f(fy * 2048);
        int y0 = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
        int y1 = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
        const uint8_t *s0 = src + (size_t)y0 * sstride, *s1 = src + (size_t)y1 * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            row0[dx] = s0[sx] * ax[2 * dx] + s0[sx1] * ax[2 * dx + 1];
            row1[dx] = s1[sx] * ax[2 * dx] + s1[sx1] * ax[2 * dx + 1];
        }
        uint8_t *d = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++) {
            int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
            d[dx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    free(xofs); free(ax); free(row0); free(row1);
}

/* ---- A.3: cv::GaussianBlur(7x7, sigma 2, REFLECT_101) for CV_8UC1 ------------------ */
static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

void orc_gaussian7_s2_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride) {
    static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
    uint16_t *hbuf = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t *s = src + (size_t)y * sstride;
        for (int x = 0; x < w; x++) {
            int acc = 0;
            if (x >= 3 && x < w - 3)
                for (int k = 0; k < 7; k++) acc += K[k] * s[x + k - 3];
            else
                for (int k = 0; k < 7; k++) acc += K[k] * s[reflect101(x + k - 3, w)];
            hbuf[(size_t)y * w + x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; y++) {
        const uint16_t *r[7];
        for (int k = 0; k < 7; k++) r[k] = hbuf + (size_t)reflect101(y + k - 3, h) * w;
        uint8_t *d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; k++) acc += (uint32_t)K[k] * r[k][x];
            d[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
    free(hbuf);
}

/* ---- A.1: cv::FAST TYPE_9_16 -------------------------------------------------------- */
static const int k_circ_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int k_circ_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* best(p) = max over the 16 arcs of 9 consecutive circle pixels of
 * max(min_k d_k, min_k -d_k), d_k = I(p) - I(p + o_k).  Corner at t iff best > t. */
static inline int fast_best(const uint8_t *p, const int *off) {
    int d[25];
    int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[off[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int best = -256;
    for (int s = 0; s < 16; s++) {
        int mn = d[s], mx = d[s];
        for (int k = 1; k < 9; k++) {
            int t = d[s + k];
            mn = t < mn ? t : mn;
            mx = t > mx ? t : mx;
        }
        if (mn > best) best = mn;
        if (-mx > best) best = -mx;
    }
    return best;
}

void orc_fast_score_u8(const uint8_t *img, int w, int h, int stride, uint8_t *score) {
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = k_circ_dy[k] * stride + k_circ_dx[k];
    memset(score, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int b = fast_best(img + (size_t)y * stride + x, off);
            score[(size_t)y * w + x] = (uint8_t)(b < 0 ? 0 : b);
        }
}

/* FAST-9-16 with non-max suppression on a sub-image, raster output order;
 * response = best - 1; neighbours that are not corners (or lie outside the
 * tested region of THIS sub-image) count as 0; strict > against all 8. */
int orc_fast_nms(const uint8_t *img, int w, int h, int stride, int t, float *xyr, int cap) {
    if (w < 7 || h < 7) return 0;
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = k_circ_dy[k] * stride + k_circ_dx[k];
    int *sc = (int *)calloc((size_t)w * h, sizeof(int));
    for (int y = 3; y < h - 3; y++) {
        const uint8_t *row = img + (size_t)y * stride;
        for (int x = 3; x < w - 3; x++) {
            const uint8_t *p = row + x;
            int v = p[0], lo = v - t, hi = v + t;
            /* every 9-arc contains one pixel of each antipodal pair */
            int a = p[off[0]], b = p[off[8]];
            if (a >= lo && a <= hi && b >= lo && b <= hi) continue;
            a = p[off[4]]; b = p[off[12]];
            if (a >= lo && a <= hi && b >= lo && b <= hi) continue;
            int best = fast_best(p, off);
            if (best > t) sc[(size_t)y * w + x] = best - 1;
        }
    }
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = sc[(size_t)y * w + x];
            if (s == 0) continue; /* corners have best-1 >= t >= 1 for t >= 1 */
            const int *c = sc + (size_t)y * w + x;
            if (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] &&
                s > c[w - 1] && s > c[w] && s > c[w + 1]) {
                if (n < cap) {
                    xyr[3 * n] = (float)x;
                    xyr[3 * n + 1] = (float)y;
                    xyr[3 * n + 2] = (float)s;
                }
                n++;
            }
        }
    free(sc);
    return n;
}

/* ---- A.4: cv::fastAtan2 ---------------------------------------------------------- */
float orc_fast_atan2(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s,
                p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ---- DistributeOctTree: src/orb_extractor.cpp:481-763, literal std::list emulation --- */
typedef struct {
    int x0, x1, y0, y1; /* UL.x, UR.x, UL.y, BL.y */
    int *keys;          /* candidate indices, inherited order */
    int cnt;
    int nomore;
    int prev, next; /* list links */
} onode;

typedef struct {
    onode *n;
    int used, cap;
    int head, tail, size;
} olist;

static int node_new(olist *L, int x0, int x1, int y0, int y1, int reserve) {
    if (L->used == L->cap) {
        L->cap = L->cap ? 2 * L->cap : 256;
        L->n = (onode *)realloc(L->n, sizeof(onode) * L->cap);
    }
    onode *nd = &L->n[L->used];
    nd->x0 = x0; nd->x1 = x1; nd->y0 = y0; nd->y1 = y1;
    nd->keys = (int *)malloc(sizeof(int) * (reserve > 0 ? reserve : 1));
    nd->cnt = 0; nd->nomore = 0; nd->prev = nd->next = -1;
    return L->used++;
}
static void list_push_back(olist *L, int i) {
    L->n[i].prev = L->tail; L->n[i].next = -1;
    if (L->tail >= 0) L->n[L->tail].next = i; else L->head = i;
    L->tail = i; L->size++;
}
static void list_push_front(olist *L, int i) {
    L->n[i].next = L->head; L->n[i].prev = -1;
    if (L->head >= 0) L->n[L->head].prev = i; else L->tail = i;
    L->head = i; L->size++;
}
static void list_erase(olist *L, int i) {
    int p = L->n[i].prev, q = L->n[i].next;
    if (p >= 0) L->n[p].next = q; else L->head = q;
    if (q >= 0) L->n[q].prev = p; else L->tail = p;
    L->size--;
}

/* DivideNode: src/orb_extractor.cpp:481-537 */
static void divide(olist *L, int pi, const float *xyr, int c[4]) {
    int x0 = L->n[pi].x0, x1 = L->n[pi].x1, y0 = L->n[pi].y0, y1 = L->n[pi].y1;
    int cnt = L->n[pi].cnt;
    int hx = (int)ceilf((float)(x1 - x0) / 2), hy = (int)ceilf((float)(y1 - y0) / 2);
    int mx = x0 + hx, my = y0 + hy;
    c[0] = node_new(L, x0, mx, y0, my, cnt);
    c[1] = node_new(L, mx, x1, y0, my, cnt);
    c[2] = node_new(L, x0, mx, my, y1, cnt);
    c[3] = node_new(L, mx, x1, my, y1, cnt);
    const int *keys = L->n[pi].keys;
    for (int i = 0; i < cnt; i++) {
        int k = keys[i];
        float x = xyr[3 * k], y = xyr[3 * k + 1];
        int q = (x < (float)mx) ? ((y < (float)my) ? 0 : 2) : ((y < (float)my) ? 1 : 3);
        onode *ch = &L->n[c[q]];
        ch->keys[ch->cnt++] = k;
    }
    for (int q = 0; q < 4; q++)
        if (L->n[c[q]].cnt == 1) L->n[c[q]].nomore = 1;
}

typedef struct { int cnt, seq; } szref;
static int szref_cmp(const void *a, const void *b) {
    const szref *p = (const szref *)a, *q = (const szref *)b;
    if (p->cnt != q->cnt) return p->cnt < q->cnt ? -1 : 1;
    return p->seq < q->seq ? -1 : (p->seq > q->seq ? 1 : 0); /* T1: creation order replaces the address */
}

/* tie_cut (optional): set to 1 when the careful phase stopped (:730-731) inside a group of nodes of EQUAL point count --
 * or would have under another order of that group -- i.e. when the reference's heap-address order (:684) decides
 * which of them are split; 0 = the survivor set does not depend on the tie rule (only its order does). */
int orc_distribute_ex(const float *xyr, int n, int min_x, int max_x, int min_y, int max_y,
                      int n_want, float *out, int cap, int *tie_cut) {
    if (tie_cut) *tie_cut = 0;
    if (n <= 0) return 0;
    const int W = max_x - min_x, H = max_y - min_y;
    const int n_ini = (int)roundf((float)W / H);
    if (n_ini < 1) return -1; /* reference divides by zero here */
    const float hX = (float)W / n_ini;
    olist L = {0};
    L.head = L.tail = -1;
    int *roots = (int *)malloc(sizeof(int) * n_ini);
    for (int i = 0; i < n_ini; i++) {
        roots[i] = node_new(&L, (int)(hX * (float)i), (int)(hX * (float)(i + 1)), 0, H, n);
        list_push_back(&L, roots[i]);
    }
    for (int i = 0; i < n; i++) {
        int r = (int)(xyr[3 * i] / hX);
        if (r >= n_ini) r = n_ini - 1; /* cannot happen for FAST output (x <= W-4) */
        onode *nd = &L.n[roots[r]];
        nd->keys[nd->cnt++] = i;
    }
    for (int it = L.head; it >= 0;) {
        int nx = L.n[it].next;
        if (L.n[it].cnt == 1) L.n[it].nomore = 1;
        else if (L.n[it].cnt == 0) list_erase(&L, it);
        it = nx;
    }
    szref *vsz = (szref *)malloc(sizeof(szref) * (4 * (size_t)n + 16));
    szref *vprev = (szref *)malloc(sizeof(szref) * (4 * (size_t)n + 16));
    int nsz = 0, finish = 0;
    while (!finish) {
        int prev_size = L.size, n_expand = 0;
        nsz = 0;
        for (int it = L.head; it >= 0;) {
            if (L.n[it].nomore) { it = L.n[it].next; continue; }
            int c[4];
            divide(&L, it, xyr, c);
            for (int q = 0; q < 4; q++)
                if (L.n[c[q]].cnt > 0) {
                    list_push_front(&L, c[q]);
                    if (L.n[c[q]].cnt > 1) {
                        n_expand++;
                        vsz[nsz].cnt = L.n[c[q]].cnt; vsz[nsz].seq = c[q]; nsz++;
                    }
                }
            int nx = L.n[it].next;
            list_erase(&L, it);
            it = nx;
        }
        if (L.size >= n_want || L.size == prev_size) {
            finish = 1;
        } else if (L.size + n_expand * 3 > n_want) {
            while (!finish) {
                prev_size = L.size;
                int np = nsz;
                memcpy(vprev, vsz, sizeof(szref) * np);
                nsz = 0;
                qsort(vprev, np, sizeof(szref), szref_cmp);
                for (int j = np - 1; j >= 0; j--) {
                    int pi = vprev[j].seq, c[4];
                    const int size_before = L.size;
                    divide(&L, pi, xyr, c);
                    for (int q = 0; q < 4; q++)
                        if (L.n[c[q]].cnt > 0) {
                            list_push_front(&L, c[q]);
                            if (L.n[c[q]].cnt > 1) {
                                vsz[nsz].cnt = L.n[c[q]].cnt; vsz[nsz].seq = c[q]; nsz++;
                            }
                        }
                    list_erase(&L, pi);
                    vprev[j].seq = L.size - size_before; /* the slot now holds the node's growth (its id is spent) */
                    if (L.size >= n_want) {
                        if (tie_cut) {
                            /* the stop is order-independent iff no other order of the equally-full group containing j
                             * reaches n_want before the group is exhausted, and the group ends here */
                            int e = j, total = 0, min_inc = 4;
                            while (e + 1 < np && vprev[e + 1].cnt == vprev[j].cnt) e++;
                            for (int t = j; t <= e; t++) {
                                total += vprev[t].seq;
                                if (vprev[t].seq < min_inc) min_inc = vprev[t].seq;
                            }
                            if (j > 0 && vprev[j - 1].cnt == vprev[j].cnt) *tie_cut = 1;
                            else if (e > j && (L.size - total) + total - min_inc >= n_want) *tie_cut = 1;
                        }
                        break;
                    }
                }
                if (L.size >= n_want || L.size == prev_size) finish = 1;
            }
        }
    }
    int m = 0;
    for (int it = L.head; it >= 0; it = L.n[it].next) {
        const onode *nd = &L.n[it];
        int bk = nd->keys[0];
        float br = xyr[3 * bk + 2];
        for (int k = 1; k < nd->cnt; k++) {
            int kk = nd->keys[k];
            if (xyr[3 * kk + 2] > br) { bk = kk; br = xyr[3 * kk + 2]; }
        }
        if (m < cap) { out[3 * m] = xyr[3 * bk]; out[3 * m + 1] = xyr[3 * bk + 1]; out[3 * m + 2] = br; }
        m++;
    }
    for (int i = 0; i < L.used; i++) free(L.n[i].keys);
    free(L.n); free(roots); free(vsz); free(vprev);
    return m;
}

int orc_distribute(const float *xyr, int n, int min_x, int max_x, int min_y, int max_y,
                   int n_want, float *out, int cap) {
    return orc_distribute_ex(xyr, n, min_x, max_x, min_y, max_y, n_want, out, cap, NULL);
}

/* ---- per-cell FAST: src/orb_extractor.cpp:765-829 ------------------------------------ */
static int level_candidates(const orc_extractor *ex, const uint8_t *img, int cols, int rows,
                            float **out, int *bx) {
    const int min_bx = EDGE_TH - 3, min_by = min_bx;
    const int max_bx = cols - EDGE_TH + 3, max_by = rows - EDGE_TH + 3;
    bx[0] = min_bx; bx[1] = max_bx; bx[2] = min_by; bx[3] = max_by;
    const float width = (float)(max_bx - min_bx), height = (float)(max_by - min_by);
    const int n_cols = (int)(width / 30.f), n_rows = (int)(height / 30.f);
    *out = NULL;
    if (n_cols < 1 || n_rows < 1) return 0; /* reference would divide by zero */
    const int w_cell = (int)ceilf(width / n_cols), h_cell = (int)ceilf(height / n_rows);
    int cap = 4096, n = 0;
    float *buf = (float *)malloc(sizeof(float) * 3 * cap);
    float cell[3 * 2048]; /* a cell tests at most ~40x40 px */
    for (int i = 0; i < n_rows; i++) {
        const int ini_y = min_by + i * h_cell;
        int max_y = ini_y + h_cell + 6;
        if (ini_y >= max_by - 3) continue;
        if (max_y > max_by) max_y = max_by;
        for (int j = 0; j < n_cols; j++) {
            const int ini_x = min_bx + j * w_cell;
            int max_x = ini_x + w_cell + 6;
            if (ini_x >= max_bx - 6) continue;
            if (max_x > max_bx) max_x = max_bx;
            const uint8_t *sub = img + (size_t)ini_y * cols + ini_x;
            int k = orc_fast_nms(sub, max_x - ini_x, max_y - ini_y, cols, ex->ini_th, cell, 2048);
            if (k == 0) k = orc_fast_nms(sub, max_x - ini_x, max_y - ini_y, cols, ex->min_th, cell, 2048);
            if (k > 2048) k = 2048;
            if (n + k > cap) {
                while (n + k > cap) cap *= 2;
                buf = (float *)realloc(buf, sizeof(float) * 3 * cap);
            }
            for (int q = 0; q < k; q++) {
                buf[3 * (n + q)] = cell[3 * q] + (float)(j * w_cell);
                buf[3 * (n + q) + 1] = cell[3 * q + 1] + (float)(i * h_cell);
                buf[3 * (n + q) + 2] = cell[3 * q + 2];
            }
            n += k;
        }
    }
    *out = buf;
    return n;
}

/* ---- IC_Angle: src/orb_extractor.cpp:77-104 ----------------------------------------- */
static float ic_angle(const uint8_t *img, int step, int x, int y, const int *umax) {
    int m_01 = 0, m_10 = 0;
    const uint8_t *center = img + (size_t)y * step + x;
    for (int u = -HALF_PATCH; u <= HALF_PATCH; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = center[u + v * step], vm = center[u - v * step];
            v_sum += (vp - vm);
            m_10 += u * (vp + vm);
        }
        m_01 += v * v_sum;
    }
    return orc_fast_atan2((float)m_01, (float)m_10);
}

/* ---- computeOrbDescriptor: src/orb_extractor.cpp:107-147 (T2 float model) ------------ */
static void orb_descriptor(const uint8_t *img, int step, int x, int y, float angle_deg, uint8_t *desc) {
    const float factor_pi = (float)(3.1415926535897932384626433832795 / 180.f);
    float angle = angle_deg * factor_pi;
    float a = (float)cos((double)angle), b = (float)sin((double)angle);
    const uint8_t *center = img + (size_t)y * step + x;
    const int8_t *pat = k_pattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            int t[2];
            for (int s = 0; s < 2; s++) {
                float px = (float)pat[4 * k + 2 * s], py = (float)pat[4 * k + 2 * s + 1];
                int ry = cv_round_f(px * b + py * a), rx = cv_round_f(px * a - py * b);
                t[s] = center[ry * step + rx];
            }
            val |= (t[0] < t[1]) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ---- extract: src/orb_extractor.cpp:1043-1105 ---------------------------------------- */
int orc_extract(orc_extractor *ex, const uint8_t *img, int w, int h, int stride, orc_keypoint *kps,
                uint8_t *desc, int cap) {
    if (!ex) return -1;
    free_stages(ex);
    if (!img || w <= 0 || h <= 0) return 0; /* empty image: silent return (:1046-1047) */
    int n = 0;
    for (int l = 0; l < ex->nlevels; l++) {
        orc_level_size(ex, w, h, l, &ex->lw[l], &ex->lh[l]);
        int lw = ex->lw[l], lh = ex->lh[l];
        if (lw < 1 || lh < 1) return -2;
        ex->level[l] = (uint8_t *)malloc((size_t)lw * lh);
        if (l == 0)
            for (int y = 0; y < h; y++) memcpy(ex->level[0] + (size_t)y * w, img + (size_t)y * stride, w);
        else
            orc_resize_linear_u8(ex->level[l - 1], ex->lw[l - 1], ex->lh[l - 1], ex->lw[l - 1],
                                 ex->level[l], lw, lh, lw);
    }
    for (int l = 0; l < ex->nlevels; l++) {
        int lw = ex->lw[l], lh = ex->lh[l], bx[4];
        ex->ncands[l] = level_candidates(ex, ex->level[l], lw, lh, &ex->cands[l], bx);
        int nc = ex->ncands[l];
        ex->dist[l] = (float *)malloc(sizeof(float) * 3 * (nc > 0 ? nc : 1));
        ex->tie_cut[l] = 0;
        int nd = nc > 0 ? orc_distribute_ex(ex->cands[l], nc, bx[0], bx[1], bx[2], bx[3],
                                             ex->per_level[l], ex->dist[l], nc, &ex->tie_cut[l]) : 0;
        if (nd < 0) return -3;
        ex->ndist[l] = nd;
    }
    for (int l = 0; l < ex->nlevels; l++) {
        int lw = ex->lw[l], lh = ex->lh[l], nd = ex->ndist[l];
        if (nd == 0) continue;
        ex->blur[l] = (uint8_t *)malloc((size_t)lw * lh);
        orc_gaussian7_s2_u8(ex->level[l], lw, lh, lw, ex->blur[l], lw);
        const float size = (float)(int)(PATCH_SIZE * ex->scale[l]);
        const float sc = ex->scale[l];
        for (int i = 0; i < nd; i++) {
            if (n >= cap) return -4;
            int x = (int)ex->dist[l][3 * i] + (EDGE_TH - 3), y = (int)ex->dist[l][3 * i + 1] + (EDGE_TH - 3);
            orc_keypoint *kp = &kps[n];
            kp->angle = ic_angle(ex->level[l], lw, x, y, ex->umax);
            orb_descriptor(ex->blur[l], lw, x, y, kp->angle, desc + (size_t)32 * n);
            kp->x = (float)x; kp->y = (float)y;
            if (l != 0) { kp->x *= sc; kp->y *= sc; }
            kp->size = size;
            kp->response = ex->dist[l][3 * i + 2];
            kp->octave = l;
            kp->class_id = -1;
            n++;
        }
    }
    return n;
}

static int copy_plane(const uint8_t *src, int w, int h, uint8_t *out) {
    if (!src) return -1;
    memcpy(out, src, (size_t)w * h);
    return 0;
}
int orc_get_level(const orc_extractor *ex, int l, uint8_t *out) { return copy_plane(ex->level[l], ex->lw[l], ex->lh[l], out); }
int orc_get_blur(const orc_extractor *ex, int l, uint8_t *out) { return copy_plane(ex->blur[l], ex->lw[l], ex->lh[l], out); }
int orc_get_score(const orc_extractor *ex, int l, uint8_t *out) {
    if (!ex->level[l]) return -1;
    orc_fast_score_u8(ex->level[l], ex->lw[l], ex->lh[l], ex->lw[l], out);
    return 0;
}
int orc_get_candidates(const orc_extractor *ex, int l, float *xyr, int cap) {
    int n = ex->ncands[l] < cap ? ex->ncands[l] : cap;
    if (n > 0) memcpy(xyr, ex->cands[l], sizeof(float) * 3 * n);
    return ex->ncands[l];
}
int orc_get_tie_cut(const orc_extractor *ex, int l) { return ex->tie_cut[l]; }
int orc_get_distributed(const orc_extractor *ex, int l, float *xyr, int cap) {
    int n = ex->ndist[l] < cap ? ex->ndist[l] : cap;
    if (n > 0) memcpy(xyr, ex->dist[l], sizeof(float) * 3 * n);
    return ex->ndist[l];
}

/* ---- DescriptorDistance: include/orb_extractor.h:87-103 ------------------------------ */
int orc_hamming256(const void *a, const void *b) {
    uint32_t pa[8], pb[8];
    memcpy(pa, a, 32);
    memcpy(pb, b, 32);
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

/* ---- StereoMatch: src/matcher.cpp:54-132.  The int(y/10) row buckets (:60-66,83-95)
 * only pre-select: |dy|<=3 < 10 keeps every passing candidate inside buckets b-1..b+1,
 * so the candidate set is "all right keypoints passing :103-110", visited in ascending j. */
void orc_stereo_match(const orc_keypoint *kl, const uint8_t *dl, int nl, const orc_keypoint *kr,
                      const uint8_t *dr, int nr, double y_thr, double max_dx, double ratio,
                      int *out_idx, int *out_dist) {
    for (int i = 0; i < nl; i++) {
        double dist0 = 999999999., dist1 = dist0;
        int champ0 = -1;
        for (int j = 0; j < nr; j++) {
            double dx = kl[i].x - kr[j].x; /* float subtraction, widened (as in the reference) */
            double dy = kl[i].y - kr[j].y;
            if (fabs(dy) > y_thr) continue;
            if (dx < 0.) continue;
            if (dx > max_dx) continue;
            double d = orc_hamming256(dl + (size_t)32 * i, dr + (size_t)32 * j);
            if (d < dist0) { dist1 = dist0; dist0 = d; champ0 = j; }
            else if (d < dist1) { dist1 = d; }
        }
        out_idx[i] = -1;
        if (out_dist) out_dist[i] = -1;
        if (champ0 < 0) continue;
        if (dist0 < dist1 * ratio) {
            out_idx[i] = champ0;
            if (out_dist) out_dist[i] = (int)dist0;
        }
    }
}

/* ---- Xc = predicted_Tcw * Xw (src/matcher.cpp:151).  The reference holds the pose as a g2o::SE3Quat, so the product is
 * g2o's `_t + _r * v` with Eigen's quaternion-vector product (Eigen/src/Geometry/Quaternion.h, _transformVector):
 *   uv = q.vec x v;  uv += uv;  r = (v + q.w * uv) + q.vec x uv;   cross(a, b) = (a1 b2 - a2 b1, a2 b0 - a0 b2, a0 b1 - a1 b0)
 * quat != 0: pose = {qx, qy, qz, qw, tx, ty, tz} (the unit quaternion as the SE3Quat holds it).
 * quat == 0: pose = row-major 3x4 [R|t], rows evaluated left to right (for callers that hold a matrix; rule T6).
 * tests/test_ref_pinning.py checks the quaternion form bit-for-bit against oracle/_ref. */
static void apply_pose(const double *p, int quat, double X, double Y, double Z, double *xc, double *yc, double *zc) {
    if (!quat) {
        *xc = ((p[0] * X + p[1] * Y) + p[2] * Z) + p[3];
        *yc = ((p[4] * X + p[5] * Y) + p[6] * Z) + p[7];
        *zc = ((p[8] * X + p[9] * Y) + p[10] * Z) + p[11];
        return;
    }
    const double qx = p[0], qy = p[1], qz = p[2], qw = p[3];
    double ux = qy * Z - qz * Y, uy = qz * X - qx * Z, uz = qx * Y - qy * X;
    ux += ux; uy += uy; uz += uz;
    const double cx = qy * uz - qz * uy, cy = qz * ux - qx * uz, cz = qx * uy - qy * ux;
    *xc = p[4] + ((X + qw * ux) + cx);
    *yc = p[5] + ((Y + qw * uy) + cy);
    *zc = p[6] + ((Z + qw * uz) + cz);
}
void orc_se3_apply(const double qt[7], const double *x, int n, double *out) {
    for (int i = 0; i < n; i++) apply_pose(qt, 1, x[3 * i], x[3 * i + 1], x[3 * i + 2], out + 3 * i, out + 3 * i + 1, out + 3 * i + 2);
}

/* ---- ProjectionMatch: src/matcher.cpp:134-209 + Camera::Project src/camera.cpp:50-79 +
 * IsInImage :26-36 + FLANN radius search (strict d^2 < r^2, T4).  Map-point order = array
 * order (T3).  Tcw given as row-major [R|t] 3x4 (T: reference uses g2o::SE3Quat). */
static void projection_match_impl(const double *xw, const uint8_t *mp_desc, const uint8_t *skip, int n,
                          const double *rt, int quat, const orc_camera *cam, const orc_keypoint *kps,
                          const uint8_t *kp_desc, int m, double radius, double ratio,
                          int *kp_to_query, int *kp_dist) {
    for (int j = 0; j < m; j++) { kp_to_query[j] = -1; kp_dist[j] = -1; }
    const double r2max = radius * radius;
    for (int i = 0; i < n; i++) {
        if (skip && skip[i]) continue;
        const double X = xw[3 * i], Y = xw[3 * i + 1], Z = xw[3 * i + 2];
        double xc, yc, zc;
        apply_pose(rt, quat, X, Y, Z, &xc, &yc, &zc);
        if (zc < 0.) continue;
        double x = xc / zc, y = yc / zc;
        double r2 = x * x + y * y, r4 = r2 * r2;
        double a1 = 2. * x * y, a2 = r2 + 2. * x * x, a3 = r2 + 2. * y * y;
        double cdist = 1. + cam->d[0] * r2 + cam->d[1] * r4;
        double xd = x * cdist + cam->d[2] * a1 + cam->d[3] * a2;
        double yd = y * cdist + cam->d[2] * a3 + cam->d[3] * a1;
        double u = cam->fx * xd + cam->cx, v = cam->fy * yd + cam->cy;
        if (u < 0. || v < 0. || u > cam->width || v > cam->height) continue;
        if (!(u == u) || !(v == v)) continue; /* NaN (zc == 0): FLANN finds nothing */
        double dist0 = 999999999., dist1 = dist0;
        int champ0 = -1;
        for (int j = 0; j < m; j++) {
            double ddx = u - (double)kps[j].x, ddy = v - (double)kps[j].y;
            double d2 = ddx * ddx + ddy * ddy;
            if (!(d2 < r2max)) continue;
            double d = orc_hamming256(mp_desc + (size_t)32 * i, kp_desc + (size_t)32 * j);
            if (d < dist0) { dist1 = dist0; dist0 = d; champ0 = j; }
            else if (d < dist1) { dist1 = d; }
        }
        if (champ0 < 0) continue;
        if (dist0 < dist1 * ratio) {
            if (kp_to_query[champ0] >= 0 && (double)kp_dist[champ0] < dist0) continue;
            kp_to_query[champ0] = i; /* ties: the later query wins (:197-204) */
            kp_dist[champ0] = (int)dist0;
        }
    }
}

/* The same ProjectionMatch with the candidate loop restricted by a 32-px bucket grid over the frame's keypoints: the
 * CPU-baseline stand-in for the reference's per-frame FLANN kd-tree (src/frame.cpp:59-68,170-178), so that the timed
 * CPU path is not charged an O(n m) scan the reference does not do.  Same candidate set (d^2 < r^2), hence the same
 * accepted matches for ratio <= 1 (a tie for the best distance fails the ratio test whatever the visiting order,
 * SURVEY §8a); tests/test_oracle_matchers.py checks it against orc_projection_match. */
void orc_projection_match(const double *xw, const uint8_t *mp_desc, const uint8_t *skip, int n, const double rt[12],
                          const orc_camera *cam, const orc_keypoint *kps, const uint8_t *kp_desc, int m, double radius,
                          double ratio, int *kp_to_query, int *kp_dist) {
    projection_match_impl(xw, mp_desc, skip, n, rt, 0, cam, kps, kp_desc, m, radius, ratio, kp_to_query, kp_dist);
}
void orc_projection_match_se3(const double *xw, const uint8_t *mp_desc, const uint8_t *skip, int n, const double qt[7],
                              const orc_camera *cam, const orc_keypoint *kps, const uint8_t *kp_desc, int m, double radius,
                              double ratio, int *kp_to_query, int *kp_dist) {
    projection_match_impl(xw, mp_desc, skip, n, qt, 1, cam, kps, kp_desc, m, radius, ratio, kp_to_query, kp_dist);
}

static void projection_match_grid_impl(const double *xw, const uint8_t *mp_desc, const uint8_t *skip, int n,
                               const double *rt, int quat, const orc_camera *cam, const orc_keypoint *kps,
                               const uint8_t *kp_desc, int m, double radius, double ratio,
                               int *kp_to_query, int *kp_dist) {
    for (int j = 0; j < m; j++) { kp_to_query[j] = -1; kp_dist[j] = -1; }
    const int gw = ((cam->width > 1 ? cam->width : 1) >> 5) + 1, gh = ((cam->height > 1 ? cam->height : 1) >> 5) + 1;
    int *start = (int *)calloc((size_t)gw * gh + 1, sizeof(int)), *fill = (int *)calloc((size_t)gw * gh, sizeof(int));
    int *order = (int *)malloc(sizeof(int) * (m > 0 ? m : 1)), *cell = (int *)malloc(sizeof(int) * (m > 0 ? m : 1));
    for (int j = 0; j < m; j++) {
        int cx = (int)floorf(kps[j].x) >> 5, cy = (int)floorf(kps[j].y) >> 5;
        cx = cx < 0 ? 0 : cx >= gw ? gw - 1 : cx;
        cy = cy < 0 ? 0 : cy >= gh ? gh - 1 : cy;
        cell[j] = cy * gw + cx;
        start[cell[j] + 1]++;
    }
    for (int c = 0; c < gw * gh; c++) start[c + 1] += start[c];
    for (int j = 0; j < m; j++) order[start[cell[j]] + fill[cell[j]]++] = j;
    const double r2max = radius * radius;
    for (int i = 0; i < n; i++) {
        if (skip && skip[i]) continue;
        const double X = xw[3 * i], Y = xw[3 * i + 1], Z = xw[3 * i + 2];
        double xc, yc, zc;
        apply_pose(rt, quat, X, Y, Z, &xc, &yc, &zc);
        if (zc < 0.) continue;
        double x = xc / zc, y = yc / zc;
        double r2 = x * x + y * y, r4 = r2 * r2;
        double a1 = 2. * x * y, a2 = r2 + 2. * x * x, a3 = r2 + 2. * y * y;
        double cdist = 1. + cam->d[0] * r2 + cam->d[1] * r4;
        double xd = x * cdist + cam->d[2] * a1 + cam->d[3] * a2;
        double yd = y * cdist + cam->d[2] * a3 + cam->d[3] * a1;
        double u = cam->fx * xd + cam->cx, v = cam->fy * yd + cam->cy;
        if (u < 0. || v < 0. || u > cam->width || v > cam->height) continue;
        if (!(u == u) || !(v == v)) continue;
        int cx0 = (int)floor(u - radius) >> 5, cx1 = (int)floor(u + radius) >> 5;
        int cy0 = (int)floor(v - radius) >> 5, cy1 = (int)floor(v + radius) >> 5;
        cx0 = cx0 < 0 ? 0 : cx0 >= gw ? gw - 1 : cx0; cx1 = cx1 < 0 ? 0 : cx1 >= gw ? gw - 1 : cx1;
        cy0 = cy0 < 0 ? 0 : cy0 >= gh ? gh - 1 : cy0; cy1 = cy1 < 0 ? 0 : cy1 >= gh ? gh - 1 : cy1;
        double dist0 = 999999999., dist1 = dist0;
        int champ0 = -1;
        for (int cy = cy0; cy <= cy1; cy++)
            for (int t = start[cy * gw + cx0]; t < start[cy * gw + cx1 + 1]; t++) {
                const int j = order[t];
                double ddx = u - (double)kps[j].x, ddy = v - (double)kps[j].y;
                double d2 = ddx * ddx + ddy * ddy;
                if (!(d2 < r2max)) continue;
                double d = orc_hamming256(mp_desc + (size_t)32 * i, kp_desc + (size_t)32 * j);
                if (d < dist0) { dist1 = dist0; dist0 = d; champ0 = j; }
                else if (d < dist1) { dist1 = d; }
            }
        if (champ0 < 0) continue;
        if (dist0 < dist1 * ratio) {
            if (kp_to_query[champ0] >= 0 && (double)kp_dist[champ0] < dist0) continue;
            kp_to_query[champ0] = i;
            kp_dist[champ0] = (int)dist0;
        }
    }
    free(start); free(fill); free(order); free(cell);
}

void orc_projection_match_grid(const double *xw, const uint8_t *mp_desc, const uint8_t *skip, int n, const double rt[12],
                               const orc_camera *cam, const orc_keypoint *kps, const uint8_t *kp_desc, int m, double radius,
                               double ratio, int *kp_to_query, int *kp_dist) {
    projection_match_grid_impl(xw, mp_desc, skip, n, rt, 0, cam, kps, kp_desc, m, radius, ratio, kp_to_query, kp_dist);
}
void orc_projection_match_grid_se3(const double *xw, const uint8_t *mp_desc, const uint8_t *skip, int n, const double qt[7],
                                   const orc_camera *cam, const orc_keypoint *kps, const uint8_t *kp_desc, int m, double radius,
                                   double ratio, int *kp_to_query, int *kp_dist) {
    projection_match_grid_impl(xw, mp_desc, skip, n, qt, 1, cam, kps, kp_desc, m, radius, ratio, kp_to_query, kp_dist);
}

/* ---- Frame glue -------------------------------------------------------------
 * Camera::NormalizedUndistort src/camera.cpp:95-109 (called per keypoint by Frame::Frame, src/frame.cpp:50-56):
 * 5 fixed-point iterations x += (x_n - Distort(D, x)), Distort as src/camera.cpp:50-68.  Doubles, no contraction. */
static void orc_distort(const double d[4], double x, double y, double *xd, double *yd) {
    double r2 = x * x + y * y, r4 = r2 * r2;
    double a1 = 2. * x * y, a2 = r2 + 2. * x * x, a3 = r2 + 2. * y * y;
    double cdist = 1. + d[0] * r2 + d[1] * r4;
    *xd = x * cdist + d[2] * a1 + d[3] * a2;
    *yd = y * cdist + d[2] * a3 + d[3] * a1;
}

void orc_normalized_undistort(const orc_camera *cam, const orc_keypoint *kps, int n, double *xy) {
    for (int i = 0; i < n; i++) {
        const double nx = ((double)kps[i].x - cam->cx) / cam->fx, ny = ((double)kps[i].y - cam->cy) / cam->fy;
        double x = nx, y = ny;
        for (int it = 0; it < 5; it++) {
            double xd, yd;
            orc_distort(cam->d, x, y, &xd, &yd);
            x += nx - xd;
            y += ny - yd;
        }
        xy[2 * i] = x;
        xy[2 * i + 1] = y;
    }
}

/* StereoFrame::GetDepth src/frame.cpp:391-409 for every left keypoint: valid[i] = 1 and xc = (n_x, n_y, 1) * depth with
 * depth = fx * baseline / dx, dx = (double)(float)(x_l - x_r) (the reference subtracts two floats); 0 = no stereo
 * correspondence; 2 = dx < 0 (the reference throws std::invalid_argument). */
void orc_stereo_depth(const orc_camera *cam, double baseline, const orc_keypoint *kps_l, const double *norm_xy, int n,
                      const orc_keypoint *kps_r, const int *stereo_idx, double *xc, uint8_t *valid) {
    for (int i = 0; i < n; i++) {
        xc[3 * i] = xc[3 * i + 1] = xc[3 * i + 2] = 0.;
        valid[i] = 0;
        const int j = stereo_idx[i];
        if (j < 0) continue;
        const float dxf = kps_l[i].x - kps_r[j].x;
        const double dx = (double)dxf;
        if (dx < 0.) { valid[i] = 2; continue; }
        const double depth = cam->fx * baseline / dx;
        xc[3 * i] = norm_xy[2 * i] * depth;
        xc[3 * i + 1] = norm_xy[2 * i + 1] * depth;
        xc[3 * i + 2] = depth;
        valid[i] = 1;
    }
}

/* ReprojectionFilter::GetOutlier src/posetracker.cpp:106-137, the quantity its threshold test sees: err[i] = the distance
 * between keypoint i and Camera::Project(Tcw Xw_i) (src/camera.cpp:50-79, no IsInImage test here), +inf for z < 0 (:122-125
 * flags those unconditionally), -1 when keypoint i has no map point (:115-116). */
static void reprojection_error_impl(const orc_camera *cam, const double *rt, int quat, const orc_keypoint *kps, int n,
                                    const double *xw, const uint8_t *has_mp, double *err) {
    for (int i = 0; i < n; i++) {
        err[i] = -1.;
        if (!has_mp[i]) continue;
        const double X = xw[3 * i], Y = xw[3 * i + 1], Z = xw[3 * i + 2];
        double xc, yc, zc;
        apply_pose(rt, quat, X, Y, Z, &xc, &yc, &zc);
        if (zc < 0.) { err[i] = INFINITY; continue; }
        double x = xc / zc, y = yc / zc, xd, yd;
        orc_distort(cam->d, x, y, &xd, &yd);
        const double u = cam->fx * xd + cam->cx, v = cam->fy * yd + cam->cy;
        const double dx = u - (double)kps[i].x, dy = v - (double)kps[i].y;
        err[i] = sqrt(dx * dx + dy * dy);
    }
}

void orc_reprojection_error(const orc_camera *cam, const double rt[12], const orc_keypoint *kps, int n, const double *xw,
                            const uint8_t *has_mp, double *err) {
    reprojection_error_impl(cam, rt, 0, kps, n, xw, has_mp, err);
}
void orc_reprojection_error_se3(const orc_camera *cam, const double qt[7], const orc_keypoint *kps, int n, const double *xw,
                                const uint8_t *has_mp, double *err) {
    reprojection_error_impl(cam, qt, 1, kps, n, xw, has_mp, err);
}

/* Frame::SearchRadius src/frame.cpp:157-178 (FLANN radiusSearch with radius^2, L2<double> on the keypoints stored as
 * doubles): all keypoints with d^2 < r^2 (T4), canonical result order = ascending keypoint index.  Returns the count;
 * at most cap indices are written. */
int orc_search_radius(const orc_keypoint *kps, int m, double u, double v, double radius, int *idx, int cap) {
    const double r2max = radius * radius;
    int n = 0;
    for (int j = 0; j < m; j++) {
        double ddx = u - (double)kps[j].x, ddy = v - (double)kps[j].y;
        double d2 = ddx * ddx + ddy * ddy;
        if (!(d2 < r2max)) continue;
        if (n < cap) idx[n] = j;
        n++;
    }
    return n;
}

/* Frame::SearchNeareast src/frame.cpp:180-193 (FLANN knnSearch, k = 1): nearest keypoint and its SQUARED distance
 * (FLANN L2 returns squared distances); canonical tie rule T6 = the smaller index.  -1 when the frame is empty. */
void orc_search_nearest(const orc_keypoint *kps, int m, double u, double v, int *kpt_index, double *dist2) {
    *kpt_index = -1;
    *dist2 = 0.;
    for (int j = 0; j < m; j++) {
        double ddx = u - (double)kps[j].x, ddy = v - (double)kps[j].y;
        double d2 = ddx * ddx + ddy * ddy;
        if (*kpt_index < 0 || d2 < *dist2) { *kpt_index = j; *dist2 = d2; }
    }
}

/* ---- BoW transform: DBoW2 TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup),
 * thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1259, on a vocabulary held as loadFromTextFile builds it (:1338-1421):
 * node 0 = root; node i >= 1 has parent[i], a 32-byte descriptor and a weight; the children of a node are the nodes that
 * name it as parent, in id order; words are numbered in id order over the nodes flagged is_leaf; isLeaf() = no children.
 * F::distance = 256-bit Hamming (FORB.cpp:81-101), strict < keeps the first child on ties.
 * T7 (canonical): the reference leaves *nid unwritten when the descent reaches a leaf above level L - levelsup; here 0. */
void orc_vocab_transform(int n_nodes, const int32_t *parent, const uint8_t *is_leaf, const uint8_t *node_desc,
                         const double *node_weight, int L, const uint8_t *features, int n, int levelsup,
                         int32_t *word_id, double *weight, int32_t *node_id) {
    int *n_child = (int *)calloc((size_t)n_nodes + 1, sizeof(int));
    int *start = (int *)calloc((size_t)n_nodes + 2, sizeof(int));
    int *list = (int *)calloc((size_t)n_nodes + 1, sizeof(int));
    int *wid = (int *)calloc((size_t)n_nodes + 1, sizeof(int));
    for (int i = 1; i < n_nodes; i++) n_child[parent[i]]++;
    for (int i = 0; i < n_nodes; i++) start[i + 1] = start[i] + n_child[i];
    for (int i = 0; i < n_nodes; i++) n_child[i] = 0;
    int words = 0;
    for (int i = 1; i < n_nodes; i++) {
        list[start[parent[i]] + n_child[parent[i]]++] = i;
        if (is_leaf[i]) wid[i] = words++;
    }
    const int nid_level = L - levelsup;
    for (int f = 0; f < n; f++) {
        const uint8_t *feat = features + (size_t)32 * f;
        int nid = 0, final_id = 0, level = 0;
        do {
            ++level;
            const int *ch = list + start[final_id];
            const int nc = start[final_id + 1] - start[final_id];
            final_id = ch[0];
            double best = (double)orc_hamming256(feat, node_desc + (size_t)32 * final_id);
            for (int c = 1; c < nc; c++) {
                const double d = (double)orc_hamming256(feat, node_desc + (size_t)32 * ch[c]);
                if (d < best) { best = d; final_id = ch[c]; }
            }
            if (level == nid_level) nid = final_id;
        } while (start[final_id + 1] > start[final_id]);
        word_id[f] = wid[final_id];
        weight[f] = node_weight[final_id];
        node_id[f] = nid;
    }
    free(n_child); free(start); free(list); free(wid);
}

/* ---- brute-force top-2: the StereoMatch/ProjectionMatch inner loop
 * with the whole database as candidate set; strict < in ascending index = lexicographic
 * (dist, idx). */
/* Brute-force top-2: per query the lexicographically smallest two (distance, row index), i.e. the strict-< ascending-index
 * loop of src/matcher.cpp:114-123 over the whole database.  Queries are split over threads; every thread walks the
 * database in cache-sized tiles of rows (ascending, so ties still resolve towards the smaller index) and scans each tile
 * for all of its queries, which is what makes the 2000 x 10 M case of BASELINE config 4 checkable in seconds.
 * The distance is DescriptorDistance (orc_hamming256's SWAR count) computed with the popcount instruction when the CPU
 * has one -- the same integer. */
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("popcnt"))) static int hamming256_popcnt(const uint64_t *a, const uint64_t *b) {
    return __builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1]) + __builtin_popcountll(a[2] ^ b[2]) +
           __builtin_popcountll(a[3] ^ b[3]);
}
__attribute__((target("popcnt"))) static void knn2_tile_popcnt(const uint64_t *qv, const uint8_t *rows, int64_t j0, int64_t j1,
                                                               int *d0, int *d1, int64_t *i0, int64_t *i1) {
    for (int64_t j = j0; j < j1; j++) {
        uint64_t r[4];
        memcpy(r, rows + (size_t)32 * j, 32);
        const int d = hamming256_popcnt(qv, r);
        if (d < *d0) { *d1 = *d0; *i1 = *i0; *d0 = d; *i0 = j; }
        else if (d < *d1) { *d1 = d; *i1 = j; }
    }
}
#define ORC_HAVE_POPCNT_PATH 1
#endif

typedef struct {
    const uint8_t *queries, *db;
    int q0, q1;
    int64_t m, idx_base;
    int32_t *out;
    int use_popcnt;
} knn_job;

static void *knn_worker(void *arg) {
    knn_job *jb = (knn_job *)arg;
    const int nq = jb->q1 - jb->q0;
    if (nq <= 0) return NULL;
    int *d0 = (int *)malloc(sizeof(int) * nq), *d1 = (int *)malloc(sizeof(int) * nq);
    int64_t *i0 = (int64_t *)malloc(sizeof(int64_t) * nq), *i1 = (int64_t *)malloc(sizeof(int64_t) * nq);
    for (int i = 0; i < nq; i++) { d0[i] = d1[i] = 999999999; i0[i] = i1[i] = -1; }
    const int64_t tile = 1024; /* 32 KB of rows: stays in L1/L2 while every query of this thread scans it */
    for (int64_t t0 = 0; t0 < jb->m; t0 += tile) {
        const int64_t t1 = t0 + tile < jb->m ? t0 + tile : jb->m;
        for (int i = 0; i < nq; i++) {
            const uint8_t *qp = jb->queries + (size_t)32 * (jb->q0 + i);
#ifdef ORC_HAVE_POPCNT_PATH
            if (jb->use_popcnt) {
                uint64_t qv[4];
                memcpy(qv, qp, 32);
                knn2_tile_popcnt(qv, jb->db, t0, t1, &d0[i], &d1[i], &i0[i], &i1[i]);
                continue;
            }
#endif
            for (int64_t j = t0; j < t1; j++) {
                const int d = orc_hamming256(qp, jb->db + (size_t)32 * j);
                if (d < d0[i]) { d1[i] = d0[i]; i1[i] = i0[i]; d0[i] = d; i0[i] = j; }
                else if (d < d1[i]) { d1[i] = d; i1[i] = j; }
            }
        }
    }
    for (int i = 0; i < nq; i++) {
        int32_t *o = jb->out + 4 * (size_t)(jb->q0 + i);
        o[0] = (int32_t)(i0[i] < 0 ? -1 : i0[i] + jb->idx_base);
        o[1] = d0[i];
        o[2] = (int32_t)(i1[i] < 0 ? -1 : i1[i] + jb->idx_base);
        o[3] = d1[i];
    }
    free(d0); free(d1); free(i0); free(i1);
    return NULL;
}

void orc_knn2_mt(const uint8_t *queries, int q, const uint8_t *db, int64_t m, int64_t idx_base, int32_t *out, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads > q) nthreads = q > 0 ? q : 1;
    int use_popcnt = 0;
#ifdef ORC_HAVE_POPCNT_PATH
    use_popcnt = __builtin_cpu_supports("popcnt") ? 1 : 0;
#endif
    knn_job jobs[256];
    pthread_t th[256];
    for (int t = 0; t < nthreads; t++) {
        knn_job jb = {queries, db, (int)((int64_t)q * t / nthreads), (int)((int64_t)q * (t + 1) / nthreads), m, idx_base, out, use_popcnt};
        jobs[t] = jb;
    }
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, knn_worker, &jobs[t]);
    knn_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}

void orc_knn2(const uint8_t *queries, int q, const uint8_t *db, int64_t m, int64_t idx_base, int32_t *out) {
    orc_knn2_mt(queries, q, db, m, idx_base, out, 1);
}

/* ---- Tracking step between consecutive stereo frames: the previous frame's keypoints that have a stereo
 * correspondence become map points through StereoFrame::GetDepth (src/frame.cpp:391-409, normalised keypoints of
 * Camera::NormalizedUndistort, src/camera.cpp:95-109), in keypoint order, and ProjectionMatch (src/matcher.cpp:134-209)
 * with Tcw = rt matches them against the current frame's keypoints.  track_idx[j] = previous-frame keypoint matched to
 * keypoint j, or -1.  Points with dx < 0 (the reference throws) are skipped. */
void orc_track_pair(const orc_camera *cam, double baseline, const double rt[12], double radius, double ratio,
                    const orc_keypoint *kl_prev, const uint8_t *dl_prev, int nl_prev, const orc_keypoint *kr_prev,
                    const int *sidx_prev, const orc_keypoint *kps, const uint8_t *desc, int n, int *track_idx,
                    int *track_dist, int use_grid) {
    const int np = nl_prev > 0 ? nl_prev : 1;
    double *nrm = (double *)malloc(sizeof(double) * 2 * np), *xc = (double *)malloc(sizeof(double) * 3 * np);
    uint8_t *valid = (uint8_t *)malloc(np);
    orc_normalized_undistort(cam, kl_prev, nl_prev, nrm);
    orc_stereo_depth(cam, baseline, kl_prev, nrm, nl_prev, kr_prev, sidx_prev, xc, valid);
    for (int i = 0; i < nl_prev; i++) valid[i] = valid[i] != 1; /* -> skip mask */
    (use_grid ? orc_projection_match_grid : orc_projection_match)(xc, dl_prev, valid, nl_prev, rt, cam, kps, desc, n, radius, ratio,
                                                                  track_idx, track_dist);
    free(nrm); free(xc); free(valid);
}

/* ---- CPU baseline driver ---------------------------------------------------------- */
typedef struct {
    const uint8_t *left, *right;
    int count, w, h, nfeatures, nlevels, ini_th, min_th;
    float scale_factor;
    int next;
    int64_t matches, kps, tracked;
    pthread_mutex_t mu;
    /* sequence mode: per-frame results kept for the tracking phase */
    int cap;
    orc_keypoint *kl_all, *kr_all;
    uint8_t *dl_all;
    int *idx_all, *nl_all;
    const orc_camera *cam;
    double baseline, radius;
} job_t;

static void *worker(void *arg) {
    job_t *jb = (job_t *)arg;
    orc_extractor *ex = orc_extractor_create(jb->nfeatures, jb->scale_factor, jb->nlevels, jb->ini_th, jb->min_th);
    int cap = jb->kl_all ? jb->cap : jb->nfeatures + 4 * jb->nlevels + 64;
    orc_keypoint *kl = (orc_keypoint *)malloc(sizeof(orc_keypoint) * cap), *kr = (orc_keypoint *)malloc(sizeof(orc_keypoint) * cap);
    uint8_t *dl = (uint8_t *)malloc((size_t)32 * cap), *dr = (uint8_t *)malloc((size_t)32 * cap);
    int *idx = (int *)malloc(sizeof(int) * cap);
    int64_t matches = 0, kps = 0;
    for (;;) {
        pthread_mutex_lock(&jb->mu);
        int f = jb->next++;
        pthread_mutex_unlock(&jb->mu);
        if (f >= jb->count) break;
        size_t off = (size_t)f * jb->w * jb->h;
        int nl = orc_extract(ex, jb->left + off, jb->w, jb->h, jb->w, kl, dl, cap);
        int nr = orc_extract(ex, jb->right + off, jb->w, jb->h, jb->w, kr, dr, cap);
        if (nl < 0 || nr < 0) continue;
        orc_stereo_match(kl, dl, nl, kr, dr, nr, 3., 100., 0.5, idx, NULL);
        for (int i = 0; i < nl; i++) matches += idx[i] >= 0;
        kps += nl + nr;
        if (jb->kl_all) {
            const size_t o = (size_t)f * jb->cap;
            memcpy(jb->kl_all + o, kl, sizeof(orc_keypoint) * nl);
            memcpy(jb->kr_all + o, kr, sizeof(orc_keypoint) * nr);
            memcpy(jb->dl_all + o * 32, dl, (size_t)32 * nl);
            memcpy(jb->idx_all + o, idx, sizeof(int) * nl);
            jb->nl_all[f] = nl;
        }
    }
    pthread_mutex_lock(&jb->mu);
    jb->matches += matches;
    jb->kps += kps;
    pthread_mutex_unlock(&jb->mu);
    free(kl); free(kr); free(dl); free(dr); free(idx);
    orc_extractor_destroy(ex);
    return NULL;
}

int64_t orc_stereo_frames(const uint8_t *left, const uint8_t *right, int count, int w, int h,
                          int nthreads, int nfeatures, float scale_factor, int nlevels, int ini_th,
                          int min_th, int64_t *total_kps) {
    job_t jb;
    memset(&jb, 0, sizeof(jb));
    jb.left = left; jb.right = right; jb.count = count; jb.w = w; jb.h = h;
    jb.nfeatures = nfeatures; jb.scale_factor = scale_factor; jb.nlevels = nlevels;
    jb.ini_th = ini_th; jb.min_th = min_th;
    pthread_mutex_init(&jb.mu, NULL);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, worker, &jb);
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    pthread_mutex_destroy(&jb.mu);
    if (total_kps) *total_kps = jb.kps;
    return jb.matches;
}

static void *track_worker(void *arg) {
    job_t *jb = (job_t *)arg;
    static const double ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    int *tidx = (int *)malloc(sizeof(int) * jb->cap), *tdist = (int *)malloc(sizeof(int) * jb->cap);
    int64_t tracked = 0;
    for (;;) {
        pthread_mutex_lock(&jb->mu);
        int f = jb->next++;
        pthread_mutex_unlock(&jb->mu);
        if (f >= jb->count) break;
        if (f == 0) continue;
        const size_t o = (size_t)f * jb->cap, q = (size_t)(f - 1) * jb->cap;
        orc_track_pair(jb->cam, jb->baseline, ident, jb->radius, 0.5, jb->kl_all + q, jb->dl_all + q * 32, jb->nl_all[f - 1],
                       jb->kr_all + q, jb->idx_all + q, jb->kl_all + o, jb->dl_all + o * 32, jb->nl_all[f], tidx, tdist, 1);
        for (int j = 0; j < jb->nl_all[f]; j++) tracked += tidx[j] >= 0;
    }
    pthread_mutex_lock(&jb->mu);
    jb->tracked += tracked;
    pthread_mutex_unlock(&jb->mu);
    free(tidx); free(tdist);
    return NULL;
}

/* CPU baseline helper for the sequence workload: orc_stereo_frames, then the tracking step of every consecutive pair
 * (identity motion prior, bucket-grid candidate search), both phases over `nthreads` threads. */
int64_t orc_stereo_sequence(const uint8_t *left, const uint8_t *right, int count, int w, int h, int nthreads, int nfeatures,
                            float scale_factor, int nlevels, int ini_th, int min_th, const orc_camera *cam, double baseline,
                            double radius, int64_t *total_kps, int64_t *total_tracked) {
    job_t jb;
    memset(&jb, 0, sizeof(jb));
    jb.left = left; jb.right = right; jb.count = count; jb.w = w; jb.h = h;
    jb.nfeatures = nfeatures; jb.scale_factor = scale_factor; jb.nlevels = nlevels;
    jb.ini_th = ini_th; jb.min_th = min_th;
    jb.cap = nfeatures + 4 * nlevels + 64;
    jb.cam = cam; jb.baseline = baseline; jb.radius = radius;
    const size_t tot = (size_t)count * jb.cap;
    jb.kl_all = (orc_keypoint *)malloc(sizeof(orc_keypoint) * tot);
    jb.kr_all = (orc_keypoint *)malloc(sizeof(orc_keypoint) * tot);
    jb.dl_all = (uint8_t *)malloc(32 * tot);
    jb.idx_all = (int *)malloc(sizeof(int) * tot);
    jb.nl_all = (int *)calloc(count, sizeof(int));
    pthread_mutex_init(&jb.mu, NULL);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, worker, &jb);
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    jb.next = 0;
    for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, track_worker, &jb);
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    pthread_mutex_destroy(&jb.mu);
    free(jb.kl_all); free(jb.kr_all); free(jb.dl_all); free(jb.idx_all); free(jb.nl_all);
    if (total_kps) *total_kps = jb.kps;
    if (total_tracked) *total_tracked = jb.tracked;
    return jb.matches;
}
