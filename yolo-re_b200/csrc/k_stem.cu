// K2: first convolution, straight from the fp32 NCHW image (Cin <= 4, 3x3, pad 1, stride 1|2).
//
//   y = act(conv(x, w) + bias)          -- reference: layers.stem1, src/yolo/blocks/conv.py:88-89
//
// HBM-bound (AI ~ 23 FLOP/B): one thread owns one output pixel, keeps its 3x3xCin patch in
// registers, and walks the output channels 16 at a time with the folded weights broadcast from
// shared memory.  Output is written channels-last (optionally as the 4 parity planes the
// stride-2 tcgen05 conv that follows wants), 32 B per store.
#include "yre_common.cuh"

namespace {

struct StemParams {
    const float* x;
    DView y;
    const float* w;     // [Cout][3][3][Cin]
    const float* bias;
    int B, Cin, H, W, Ho, Wo, Cout, stride, act;
};

template <typename TOut, int CIN>
__global__ void __launch_bounds__(128) stem_kernel(const StemParams p) {
    extern __shared__ __align__(16) float sw[];               // [9*CIN][Cout] + bias[Cout]
    const int K = 9 * CIN;
    for (int i = threadIdx.x; i < K * p.Cout; i += blockDim.x) {
        const int co = i / K, kk = i % K;       // global order [co][tap][ci]
        sw[kk * p.Cout + co] = p.w[i];
    }
    float* sb = sw + K * p.Cout;
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) sb[i] = p.bias ? p.bias[i] : 0.f;
    __syncthreads();

    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)p.B * p.Ho * p.Wo;
    if (pix >= total) return;
    const int ox = (int)(pix % p.Wo);
    const long long t = pix / p.Wo;
    const int oy = (int)(t % p.Ho), b = (int)(t / p.Ho);

    float in[9 * CIN];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int iy = oy * p.stride + dy - 1, ix = ox * p.stride + dx - 1;
            const bool ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci)
                in[(dy * 3 + dx) * CIN + ci] = ok ? __ldg(p.x + (((long long)b * CIN + ci) * p.H + iy) * p.W + ix) : 0.f;
        }

    const long long obase = dview_pix(p.y, b, oy, ox);
    for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = sb[c0 + j];
#pragma unroll
        for (int kk = 0; kk < 9 * CIN; ++kk) {
            const float v = in[kk];
            const float4* wr = reinterpret_cast<const float4*>(sw + kk * p.Cout + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w4 = wr[q];
                acc[q * 4 + 0] = fmaf(v, w4.x, acc[q * 4 + 0]);
                acc[q * 4 + 1] = fmaf(v, w4.y, acc[q * 4 + 1]);
                acc[q * 4 + 2] = fmaf(v, w4.z, acc[q * 4 + 2]);
                acc[q * 4 + 3] = fmaf(v, w4.w, acc[q * 4 + 3]);
            }
        }
        if (p.act == YRE_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = silu_f(acc[j]);
        }
        st8<TOut>(p.y.ptr, obase + c0, acc);
        st8<TOut>(p.y.ptr, obase + c0 + 8, acc + 8);
    }
}

}  // namespace

int launch_stem(const yre_stem_desc& d, cudaStream_t s) {
    if (!d.x_nchw || !d.w) YRE_FAIL(YRE_EINVAL, "stem: null pointer");
    if (yre_check_view(&d.y, "stem.y")) return YRE_EINVAL;
    if (d.Cin < 1 || d.Cin > 4) YRE_FAIL(YRE_EUNSUPPORTED, "stem: Cin=%d (supported 1..4)", d.Cin);
    if (d.stride != 1 && d.stride != 2) YRE_FAIL(YRE_EUNSUPPORTED, "stem: stride %d", d.stride);
    if (d.y.C % 16 || d.y.c_off % 8 || d.y.C_total % 8) YRE_FAIL(YRE_EUNSUPPORTED, "stem: Cout must be a multiple of 16");
    const int Ho = (d.H + 2 - 3) / d.stride + 1, Wo = (d.W + 2 - 3) / d.stride + 1;
    if (Ho != d.y.H || Wo != d.y.W || d.B != d.y.B) YRE_FAIL(YRE_EINVAL, "stem: output extent mismatch");
    StemParams p;
    p.x = d.x_nchw; p.y = make_dview(d.y); p.w = d.w; p.bias = d.bias;
    p.B = d.B; p.Cin = d.Cin; p.H = d.H; p.W = d.W; p.Ho = Ho; p.Wo = Wo; p.Cout = d.y.C; p.stride = d.stride; p.act = d.act;
    const long long total = (long long)d.B * Ho * Wo;
    const size_t smem = (size_t)(9 * d.Cin + 1) * d.y.C * sizeof(float);
    if (smem > 48 * 1024) YRE_FAIL(YRE_EUNSUPPORTED, "stem: Cout too large for the weight cache");
    dim3 grid(yre_cdiv(total, 128));
#define STEM_GO(T, C) stem_kernel<T, C><<<grid, 128, smem, s>>>(p)
#define STEM_T(T) switch (d.Cin) { case 1: STEM_GO(T, 1); break; case 2: STEM_GO(T, 2); break; case 3: STEM_GO(T, 3); break; default: STEM_GO(T, 4); }
    if (d.y.dtype == YRE_F32) { STEM_T(float) } else { STEM_T(__nv_bfloat16) }
#undef STEM_T
#undef STEM_GO
    YRE_LAUNCH_CHECK("stem");
    return YRE_OK;
}
