/* ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of one image of yolo-re's non_max_suppression
 * (src/yolo/utils/nms.py:46-92) including the greedy step the reference delegates to
 * torchvision.ops.nms (nms.py:99-102; semantics restated in nms.py:107-152).
 * Same results as oracle/nms_ref.py::nms_image_numpy; exists because the stress
 * configuration has 33 600 candidates per image.  Build with -ffp-contract=off so that
 * every fp32 operation is rounded separately, as in the reference's tensor expressions.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float score; int32_t idx; } cand_t;

/* stable descending by score == ascending by (-score, idx) */
static int cmp_cand(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

static inline float fmax2(float a, float b) { return a > b ? a : b; }
static inline float fmin2(float a, float b) { return a < b ? a : b; }

/* pred [A][4+nc] fp32. classes may be NULL (n_classes < 0 => no filter).
 * det [max_det][6], keep_anchor [max_det]. Returns number of detections, <0 on error. */
int yre_oracle_nms_image(const float* pred, int A, int nc, float conf_thres, double iou_thres, int max_det,
                         const int32_t* classes, int n_classes, int agnostic,
                         float* det, int64_t* keep_anchor)
{
    const int ch = 4 + nc;
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)(A > 0 ? A : 1));
    int32_t* cls = (int32_t*)malloc(sizeof(int32_t) * (size_t)(A > 0 ? A : 1));
    float* box = (float*)malloc(sizeof(float) * 4 * (size_t)(A > 0 ? A : 1));   /* un-offset xyxy */
    float* nb = (float*)malloc(sizeof(float) * 4 * (size_t)(A > 0 ? A : 1));    /* offset xyxy   */
    float* area = (float*)malloc(sizeof(float) * (size_t)(A > 0 ? A : 1));
    uint8_t* dead = (uint8_t*)calloc((size_t)(A > 0 ? A : 1), 1);
    if (!c || !cls || !box || !nb || !area || !dead) { free(c); free(cls); free(box); free(nb); free(area); free(dead); return -1; }

    int n = 0;
    float max_coord = 0.f; int have_max = 0;
    for (int a = 0; a < A; ++a) {
        const float* row = pred + (size_t)a * ch;
        int best = 0; float bs = row[4];
        for (int k = 1; k < nc; ++k) if (row[4 + k] > bs) { bs = row[4 + k]; best = k; }  /* first max */
        if (!(bs > conf_thres)) continue;                                                /* strict > */
        if (n_classes >= 0) {
            int ok = 0;
            for (int k = 0; k < n_classes; ++k) ok |= (classes[k] == best);
            if (!ok) continue;
        }
        const float hw = row[2] / 2.f, hh = row[3] / 2.f;
        float* b = box + 4 * (size_t)a;
        b[0] = row[0] - hw; b[1] = row[1] - hh; b[2] = row[0] + hw; b[3] = row[1] + hh;
        for (int k = 0; k < 4; ++k) if (!have_max || b[k] > max_coord) { max_coord = b[k]; have_max = 1; }
        cls[a] = best;
        c[n].score = bs; c[n].idx = a; ++n;
    }
    const float scale = max_coord + 1.f;
    for (int i = 0; i < n; ++i) {
        const int a = c[i].idx;
        const float off = agnostic ? 0.f : (float)cls[a] * scale;
        const float* b = box + 4 * (size_t)a; float* o = nb + 4 * (size_t)a;
        if (agnostic) { o[0] = b[0]; o[1] = b[1]; o[2] = b[2]; o[3] = b[3]; }
        else { o[0] = b[0] + off; o[1] = b[1] + off; o[2] = b[2] + off; o[3] = b[3] + off; }
        area[a] = (o[2] - o[0]) * (o[3] - o[1]);
    }
    qsort(c, (size_t)n, sizeof(cand_t), cmp_cand);

    int kept = 0;
    for (int p = 0; p < n && kept < max_det; ++p) {
        const int i = c[p].idx;
        if (dead[i]) continue;
        const float* bi = box + 4 * (size_t)i;
        det[6 * kept + 0] = bi[0]; det[6 * kept + 1] = bi[1]; det[6 * kept + 2] = bi[2]; det[6 * kept + 3] = bi[3];
        det[6 * kept + 4] = c[p].score; det[6 * kept + 5] = (float)cls[i];
        keep_anchor[kept] = i; ++kept;
        const float* oi = nb + 4 * (size_t)i; const float ai = area[i];
        for (int q = p + 1; q < n; ++q) {
            const int j = c[q].idx;
            if (dead[j]) continue;
            const float* oj = nb + 4 * (size_t)j;
            const float w = fmax2(0.f, fmin2(oi[2], oj[2]) - fmax2(oi[0], oj[0]));
            const float h = fmax2(0.f, fmin2(oi[3], oj[3]) - fmax2(oi[1], oj[1]));
            const float inter = w * h;
            const float iou = inter / (ai + area[j] - inter);
            if ((double)iou > iou_thres) dead[j] = 1;   /* fp32 IoU vs DOUBLE threshold, as torchvision's CPU kernel */
        }
    }
    free(c); free(cls); free(box); free(nb); free(area); free(dead);
    return kept;
}
