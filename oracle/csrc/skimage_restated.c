/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * CPU restatement of the two scikit-image routines the `bs segment` ws path
 * calls.  scikit-image is a third-party dependency that is NOT vendored in
 * /root/reference and NOT installed in this image (pyproject.toml:21, unpinned),
 * so these are written from the published algorithm; PARITY UNPINNED.
 *
 *   sk_watershed  <- skimage.segmentation.watershed(image, markers, mask=mask)
 *                    with defaults connectivity=1, compactness=0,
 *                    watershed_line=False.  Call site: post/ws.py:26-28.
 *   sk_label      <- skimage.measure.label(x, return_num=True) (full
 *                    connectivity, background 0).  Call site:
 *                    post/blockwise/watershed_frags.py:222.
 *
 * Build: see oracle/build.py (gcc -O2 -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* binary heap exactly as skimage's heap_general.pxi / heap_watershed  */
/* ------------------------------------------------------------------ */
typedef struct {
    double value;
    int64_t age;
    int64_t index;
} item_t;

typedef struct {
    item_t *d;
    int64_t n, cap;
} heap_t;

static inline int smaller(const item_t *a, const item_t *b) {
    if (a->value != b->value) return a->value < b->value;
    return a->age < b->age;
}

static void heap_push(heap_t *h, item_t e) {
    if (h->n == h->cap) {
        h->cap *= 2;
        h->d = (item_t *)realloc(h->d, sizeof(item_t) * (size_t)h->cap);
    }
    int64_t child = h->n;
    h->d[child] = e;
    h->n++;
    while (child > 0) {
        int64_t parent = (child + 1) / 2 - 1;
        if (smaller(&h->d[child], &h->d[parent])) {
            item_t t = h->d[child];
            h->d[child] = h->d[parent];
            h->d[parent] = t;
            child = parent;
        } else
            break;
    }
}

static item_t heap_pop(heap_t *h) {
    item_t out = h->d[0];
    h->n--;
    if (h->n == 0) return out;
    /* swap(0, last) then sift down */
    item_t t = h->d[0];
    h->d[0] = h->d[h->n];
    h->d[h->n] = t;
    int64_t i = 0, smallest = 0;
    for (;;) {
        int64_t l = 2 * i + 1, r = 2 * i + 2;
        if (l < h->n) {
            if (smaller(&h->d[l], &h->d[i])) smallest = l;
            if (r < h->n && smaller(&h->d[r], &h->d[smallest])) smallest = r;
        } else
            break;
        if (smallest == i) break;
        t = h->d[i];
        h->d[i] = h->d[smallest];
        h->d[smallest] = t;
        i = smallest;
    }
    return out;
}

/*
 * image   : float64, C order, shape[ndim]   (ndim 2 or 3)
 * markers : int64,   same shape  (already offset; multiplied by mask here)
 * mask    : uint8,   same shape
 * out     : int64,   same shape
 * seed_tie: 0 = faithful: all seeds pushed with age 0, ties among equal
 *               (value, age) resolved by the binary heap's layout history,
 *               exactly as skimage does.
 *           1 = "index" rule: seeds get distinct negative ages in ascending
 *               raveled index, which makes (value, age) a strict total order
 *               (= FIFO bucket queue).  This is the order the CUDA path
 *               implements; see DESIGN.md "declared deviation D1".
 * returns the number of pops (for statistics).
 */
int64_t sk_watershed(const double *image, const int64_t *markers,
                     const uint8_t *mask, int ndim, const int64_t *shape,
                     int64_t *out, int seed_tie) {
    int64_t sz = 1, psz = 1;
    int64_t pshape[3], pstride[3];
    for (int d = 0; d < ndim; d++) {
        sz *= shape[d];
        pshape[d] = shape[d] + 2;
        psz *= pshape[d];
    }
    pstride[ndim - 1] = 1;
    for (int d = ndim - 2; d >= 0; d--) pstride[d] = pstride[d + 1] * pshape[d + 1];

    double *pimg = (double *)calloc((size_t)psz, sizeof(double));
    uint8_t *pmask = (uint8_t *)calloc((size_t)psz, 1);
    int64_t *pout = (int64_t *)calloc((size_t)psz, sizeof(int64_t));

    /* pad(…, 1, constant 0) ; markers * mask */
    int64_t c[3] = {0, 0, 0};
    for (int64_t i = 0; i < sz; i++) {
        int64_t p = 0;
        for (int d = 0; d < ndim; d++) p += (c[d] + 1) * pstride[d];
        pimg[p] = image[i];
        pmask[p] = mask[i] ? 1 : 0;
        pout[p] = mask[i] ? markers[i] : 0;
        for (int d = ndim - 1; d >= 0; d--) {
            if (++c[d] < shape[d]) break;
            c[d] = 0;
        }
    }

    /* conn-1 neighbours, stable-sorted by distance, centre dropped:
       2-D (-row,-col,+col,+row); 3-D (-z,-y,-x,+x,+y,+z) */
    int64_t nb[6];
    int nnb = 0;
    for (int d = 0; d < ndim; d++) nb[nnb++] = -pstride[d];
    for (int d = ndim - 1; d >= 0; d--) nb[nnb++] = pstride[d];

    heap_t h;
    h.cap = 1024;
    h.n = 0;
    h.d = (item_t *)malloc(sizeof(item_t) * (size_t)h.cap);

    int64_t nseeds = 0;
    if (seed_tie == 1)
        for (int64_t p = 0; p < psz; p++)
            if (pout[p] != 0) nseeds++;
    int64_t k = 0;
    for (int64_t p = 0; p < psz; p++) {
        if (pout[p] == 0) continue;
        item_t e;
        e.value = pimg[p];
        e.age = (seed_tie == 1) ? (k - nseeds) : 0;
        e.index = p;
        heap_push(&h, e);
        k++;
    }
    int64_t age = 1, pops = 0;
    while (h.n > 0) {
        item_t e = heap_pop(&h);
        pops++;
        for (int i = 0; i < nnb; i++) {
            int64_t q = e.index + nb[i];
            if (!pmask[q]) continue;
            if (pout[q]) continue;
            age++;
            pout[q] = pout[e.index]; /* labelled at push time */
            item_t ne;
            ne.value = pimg[q]; /* not clamped to e.value */
            ne.age = age;
            ne.index = q;
            heap_push(&h, ne);
        }
    }

    /* crop */
    c[0] = c[1] = c[2] = 0;
    for (int64_t i = 0; i < sz; i++) {
        int64_t p = 0;
        for (int d = 0; d < ndim; d++) p += (c[d] + 1) * pstride[d];
        out[i] = pout[p];
        for (int d = ndim - 1; d >= 0; d--) {
            if (++c[d] < shape[d]) break;
            c[d] = 0;
        }
    }
    free(h.d);
    free(pimg);
    free(pmask);
    free(pout);
    return pops;
}

/* ------------------------------------------------------------------ */
/* skimage.measure.label: CC of equal-valued non-zero elements, full   */
/* connectivity (8 / 26), ids 1..n in raster order of first element.   */
/* ------------------------------------------------------------------ */
static int64_t uf_find(int64_t *p, int64_t x) {
    while (p[x] != x) {
        p[x] = p[p[x]];
        x = p[x];
    }
    return x;
}

int64_t sk_label(const int64_t *x, int ndim, const int64_t *shape, int64_t *out) {
    int64_t Z = ndim == 3 ? shape[0] : 1;
    int64_t Y = shape[ndim - 2], X = shape[ndim - 1];
    int64_t n = Z * Y * X;
    int64_t *par = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    for (int64_t i = 0; i < n; i++) par[i] = i;
    for (int64_t z = 0; z < Z; z++)
        for (int64_t y = 0; y < Y; y++)
            for (int64_t xx = 0; xx < X; xx++) {
                int64_t i = (z * Y + y) * X + xx;
                int64_t v = x[i];
                if (v == 0) continue;
                /* all 13 raster-preceding neighbours */
                for (int64_t dz = -1; dz <= 0; dz++)
                    for (int64_t dy = -1; dy <= 1; dy++)
                        for (int64_t dx = -1; dx <= 1; dx++) {
                            if (dz == 0 && (dy > 0 || (dy == 0 && dx >= 0))) continue;
                            int64_t zz = z + dz, yy = y + dy, x2 = xx + dx;
                            if (zz < 0 || yy < 0 || yy >= Y || x2 < 0 || x2 >= X) continue;
                            int64_t j = (zz * Y + yy) * X + x2;
                            if (x[j] != v) continue;
                            int64_t a = uf_find(par, i), b = uf_find(par, j);
                            if (a < b)
                                par[b] = a;
                            else if (b < a)
                                par[a] = b;
                        }
            }
    /* roots are minimal raveled indices -> raster order numbering */
    int64_t next = 0;
    for (int64_t i = 0; i < n; i++) {
        if (x[i] == 0) {
            out[i] = 0;
            continue;
        }
        int64_t r = uf_find(par, i);
        if (r == i)
            out[i] = ++next;
        else
            out[i] = out[r];
    }
    free(par);
    return next;
}
