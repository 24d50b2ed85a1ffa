// K11: detection -> ground-truth matching for mAP (the evaluator fast path).
//   replaces the per-prediction Python loop of compute_map   reference src/yolo/eval/metrics.py:137-176
//            box_iou                                          reference src/yolo/eval/metrics.py:10-31
// One CTA per image.  Phase 1 (parallel over detections): best-IoU ground truth of the detection's class, first maximum,
// IoU in fp32 with torch's operation order (explicit _rn intrinsics: no FMA contraction).  Phase 2 (one thread per IoU
// threshold): the greedy scan in matching order -- a detection is a true positive iff its best IoU reaches the threshold
// (compared in fp32, as torch does with a Python float) and that ground truth is still unmatched at this threshold.
// Index / flag work: results are bit-exact with the reference.
#include "yre_common.cuh"

namespace {

constexpr int MT_CHUNK = 1024;     // detections per pass
constexpr int MT_MAXGT = 4096;     // ground truths per image (bitmask per threshold in shared memory)

struct MatchParams {
    const float* det; int det_stride; const int* det_off;
    const float* gt; const int* gt_cls; const int* gt_off;
    int n_thr; float thr[16];
    unsigned char* tp;
};

__global__ void __launch_bounds__(256) match_kernel(const MatchParams p) {
    __shared__ float s_iou[MT_CHUNK];
    __shared__ int s_gt[MT_CHUNK];
    __shared__ unsigned s_matched[16][MT_MAXGT / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int d0 = p.det_off[b], n = p.det_off[b + 1] - d0;
    const int g0 = p.gt_off[b], m = p.gt_off[b + 1] - g0;
    for (int i = tid; i < 16 * (MT_MAXGT / 32); i += blockDim.x) (&s_matched[0][0])[i] = 0u;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += MT_CHUNK) {
        const int cnt = min(MT_CHUNK, n - c0);
        for (int i = tid; i < cnt; i += blockDim.x) {
            const float* r = p.det + (long long)(d0 + c0 + i) * p.det_stride;
            const float x1 = r[0], y1 = r[1], x2 = r[2], y2 = r[3];
            const int cls = (int)r[5];
            const float a1 = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
            float best = -1.f; int bj = -1;
            for (int j = 0; j < m; ++j) {
                if (p.gt_cls[g0 + j] != cls) continue;
                const float4 g = *reinterpret_cast<const float4*>(p.gt + 4ll * (g0 + j));
                const float a2 = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
                const float w = fmaxf(__fsub_rn(fminf(x2, g.z), fmaxf(x1, g.x)), 0.f);
                const float h = fmaxf(__fsub_rn(fminf(y2, g.w), fmaxf(y1, g.y)), 0.f);
                const float inter = __fmul_rn(w, h);
                const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a1, a2), inter));
                if (bj < 0 || iou > best) { best = iou; bj = j; }          // first maximum
            }
            s_iou[i] = best; s_gt[i] = bj;
        }
        __syncthreads();
        if (tid < p.n_thr) {
            const float thr = p.thr[tid];
            for (int i = 0; i < cnt; ++i) {
                const int j = s_gt[i];
                unsigned char flag = 0;
                if (j >= 0 && s_iou[i] >= thr && !((s_matched[tid][j >> 5] >> (j & 31)) & 1u)) {
                    flag = 1;
                    s_matched[tid][j >> 5] |= 1u << (j & 31);
                }
                p.tp[(long long)(d0 + c0 + i) * p.n_thr + tid] = flag;
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int yre_match_detections(const yre_match_desc* d, yre_stream_t s) {
    if (!d || d->B < 0 || d->n_thr < 1 || d->n_thr > 16 || d->det_stride < 6) YRE_FAIL(YRE_EINVAL, "match: bad descriptor");
    if (d->B == 0) return YRE_OK;
    if (!d->det_off || !d->gt_off || !d->tp) YRE_FAIL(YRE_EINVAL, "match: null pointer");
    if (d->max_gt_per_image > MT_MAXGT) YRE_FAIL(YRE_EUNSUPPORTED, "match: %d ground truths in one image (max %d)", d->max_gt_per_image, MT_MAXGT);
    MatchParams p;
    p.det = d->det; p.det_stride = d->det_stride; p.det_off = d->det_off;
    p.gt = d->gt_boxes; p.gt_cls = d->gt_cls; p.gt_off = d->gt_off;
    p.n_thr = d->n_thr;
    for (int i = 0; i < 16; ++i) p.thr[i] = i < d->n_thr ? (float)d->thr[i] : 2.f;      // fp32 compare, like torch
    p.tp = d->tp;
    match_kernel<<<d->B, 256, 0, (cudaStream_t)s>>>(p);
    YRE_LAUNCH_CHECK("match_detections");
    return YRE_OK;
}
