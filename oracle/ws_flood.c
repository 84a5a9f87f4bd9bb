/*
 * oracle/ws_flood.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement of the marker-controlled watershed that the reference calls at
 * postproc_other.py:47  (`watershed(-dist, marker, mask=pred)`, scikit-image,
 * connectivity 1, compactness 0, no watershed line).  scikit-image is NOT vendored in
 * /root/reference and is absent from this image (unpinned dependency, probably
 * 0.16-0.18), so the published algorithm is restated here:
 *
 *   - every non-zero marker pixel inside the mask is pushed with age 0, in raster order;
 *   - repeatedly pop the smallest (value, age) element e; for each 4-neighbour q of e in
 *     raveled-offset order (-W, -1, +1, +W): skip if q is outside the mask or already
 *     labelled; otherwise age += 1, out[q] = out[e] (labelled when PUSHED) and push
 *     (image[q], age, q).
 *
 * Two orderings are provided:
 *   ws_flood_stable : total order (value, age, raster index).  This is the canonical order
 *                     of this build (SURVEY.md section 7 hard-part 1): it decomposes per foreground
 *                     component and is what the CUDA kernel implements.
 *   ws_flood_heap   : the order produced by a plain array binary heap comparing (value, age)
 *                     only, as recalled from scikit-image's heap_general.pxi (push = append +
 *                     sift up by swapping; pop = move last to root + sift down choosing the
 *                     smaller child).  Ties between age-0 marker pixels of equal value follow
 *                     heap mechanics.  Used only to QUANTIFY how many pixels depend on that
 *                     implementation detail ("parity unpinned").
 *
 * Also here: conv11_fma, the 11x11 two-channel correlation that restates what
 * torch.nn.functional.conv2d (CPU, f32) produces at my_transforms_direction.py:827-830 -- measured
 * to be bit-identical to a sequential f32 FMA chain over the taps in (kh, kw) row-major order
 * starting from +0 (SURVEY.md section 7 hard-part 2; re-asserted by tests/test_conv_fma_property.py).
 *
 * Build: gcc -O2 -shared -fPIC -o _build/liboracle.so ws_flood.c -lm   (see oracle/build.py)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct {
    double  value;
    int64_t age;
    int64_t index;
} item_t;

typedef int (*less_fn)(const item_t *, const item_t *);

static int less_stable(const item_t *a, const item_t *b) {
    if (a->value != b->value) return a->value < b->value;
    if (a->age != b->age) return a->age < b->age;
    return a->index < b->index;
}

static int less_value_age(const item_t *a, const item_t *b) {
    if (a->value != b->value) return a->value < b->value;
    return a->age < b->age;
}

typedef struct {
    item_t *a;
    int64_t n, cap;
    less_fn less;
} heap_t;

static int heap_push(heap_t *h, const item_t *e) {
    if (h->n == h->cap) {
        int64_t nc = h->cap ? h->cap * 2 : 1024;
        item_t *na = (item_t *)realloc(h->a, (size_t)nc * sizeof(item_t));
        if (!na) return -1;
        h->a = na;
        h->cap = nc;
    }
    int64_t child = h->n++;
    h->a[child] = *e;
    while (child > 0) {
        int64_t parent = (child + 1) / 2 - 1;
        if (h->less(&h->a[child], &h->a[parent])) {
            item_t t = h->a[child]; h->a[child] = h->a[parent]; h->a[parent] = t;
            child = parent;
        } else {
            break;
        }
    }
    return 0;
}

static void heap_pop(heap_t *h, item_t *dest) {
    *dest = h->a[0];
    h->n -= 1;
    if (h->n == 0) return;
    h->a[0] = h->a[h->n];
    int64_t i = 0;
    for (;;) {
        int64_t smallest = i, l = 2 * i + 1, r = 2 * i + 2;
        if (l >= h->n) break;
        if (h->less(&h->a[l], &h->a[i])) smallest = l;
        if (r < h->n && h->less(&h->a[r], &h->a[smallest])) smallest = r;
        if (smallest == i) break;
        item_t t = h->a[i]; h->a[i] = h->a[smallest]; h->a[smallest] = t;
        i = smallest;
    }
}

/* image: H*W doubles; markers: H*W int32 (already multiplied by the mask by the caller or not --
 * it is re-masked here as scikit-image does); mask: H*W uint8; out: H*W int32. */
static int flood(const double *image, const int32_t *markers, const uint8_t *mask,
                 int H, int W, int32_t *out, less_fn less) {
    const int64_t n = (int64_t)H * W;
    heap_t hp = {0, 0, 0, less};
    item_t e, ne;
    int64_t age = 0;
    for (int64_t i = 0; i < n; ++i) out[i] = mask[i] ? markers[i] : 0;
    for (int64_t i = 0; i < n; ++i) {
        if (out[i]) {
            e.value = image[i]; e.age = 0; e.index = i;
            if (heap_push(&hp, &e)) { free(hp.a); return -1; }
        }
    }
    while (hp.n > 0) {
        heap_pop(&hp, &e);
        const int64_t p = e.index;
        const int y = (int)(p / W), x = (int)(p % W);
        for (int k = 0; k < 4; ++k) {
            int64_t q;
            if (k == 0) { if (y == 0) continue; q = p - W; }
            else if (k == 1) { if (x == 0) continue; q = p - 1; }
            else if (k == 2) { if (x == W - 1) continue; q = p + 1; }
            else { if (y == H - 1) continue; q = p + W; }
            if (!mask[q] || out[q]) continue;
            age += 1;
            out[q] = out[p];
            ne.value = image[q]; ne.age = age; ne.index = q;
            if (heap_push(&hp, &ne)) { free(hp.a); return -1; }
        }
    }
    free(hp.a);
    return 0;
}

int ws_flood_stable(const double *image, const int32_t *markers, const uint8_t *mask,
                    int H, int W, int32_t *out) {
    return flood(image, markers, mask, H, W, out, less_stable);
}

int ws_flood_heap(const double *image, const int32_t *markers, const uint8_t *mask,
                  int H, int W, int32_t *out) {
    return flood(image, markers, mask, H, W, out, less_value_age);
}

/* img: H*W f32; ker: 2*11*11 f32 (channel 0 = d/dy, 1 = d/dx, SegFix_offset_helper.py:122-132);
 * out: 2*H*W f32; zero padding 5.  If sel != NULL only pixels with sel[p] != 0 are computed
 * (the others are written as +0). */
void conv11_fma(const float *img, int H, int W, const float *ker, float *out, const uint8_t *sel) {
    for (int c = 0; c < 2; ++c) {
        const float *k = ker + c * 121;
        float *o = out + (int64_t)c * H * W;
        for (int y = 0; y < H; ++y) {
            for (int x = 0; x < W; ++x) {
                float acc = 0.0f;
                if (!sel || sel[(int64_t)y * W + x]) {
                    for (int kh = 0; kh < 11; ++kh) {
                        const int yy = y + kh - 5;
                        for (int kw = 0; kw < 11; ++kw) {
                            const int xx = x + kw - 5;
                            const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W)
                                                ? img[(int64_t)yy * W + xx] : 0.0f;
                            acc = fmaf(k[kh * 11 + kw], v, acc);
                        }
                    }
                }
                o[(int64_t)y * W + x] = acc;
            }
        }
    }
}
