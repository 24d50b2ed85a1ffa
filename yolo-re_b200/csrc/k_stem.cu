// K2: first convolution, straight from the fp32 NCHW image (Cin <= 4, 3x3, pad 1, stride 1|2).
//
//   y = act(conv(x, w) + bias)          -- reference: layers.stem1, src/yolo/blocks/conv.py:88-89
//
// HBM-bound (AI ~ 23 FLOP/B): one thread owns 4 consecutive output pixels of a row, keeps their
// 3 x (4*stride+1) x Cin input patch in registers, and walks the output channels 16 at a time with
// the folded weights broadcast from shared memory (one weight read feeds 4 pixels).  Output is written channels-last (optionally as the 4 parity planes the
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

constexpr int PX = 4;   // output pixels per thread (consecutive in x): weights are read once per 4 pixels

template <typename TOut, int CIN, int STRIDE>
__global__ void __launch_bounds__(128) stem_kernel(const StemParams p) {
    extern __shared__ __align__(16) float sw[];               // [9*CIN][Cout] + bias[Cout]
    constexpr int K = 9 * CIN;
    constexpr int NCOL = (PX - 1) * STRIDE + 3;
    for (int i = threadIdx.x; i < K * p.Cout; i += blockDim.x) {
        const int co = i / K, kk = i % K;       // global order [co][tap][ci]
        sw[kk * p.Cout + co] = p.w[i];
    }
    float* sb = sw + K * p.Cout;
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) sb[i] = p.bias ? p.bias[i] : 0.f;
    __syncthreads();

    const int xg = (p.Wo + PX - 1) / PX;                      // pixel groups per output row
    const long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)p.B * p.Ho * xg;
    if (gidx >= total) return;
    const int ox0 = (int)(gidx % xg) * PX;
    const long long t = gidx / xg;
    const int oy = (int)(t % p.Ho), b = (int)(t / p.Ho);

    float in[3][NCOL][CIN];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int iy = oy * STRIDE + dy - 1;
        const bool rok = iy >= 0 && iy < p.H;
#pragma unroll
        for (int cx = 0; cx < NCOL; ++cx) {
            const int ix = ox0 * STRIDE + cx - 1;
            const bool ok = rok && ix >= 0 && ix < p.W;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci)
                in[dy][cx][ci] = ok ? __ldg(p.x + (((long long)b * CIN + ci) * p.H + iy) * p.W + ix) : 0.f;
        }
    }

    long long obase[PX];
#pragma unroll
    for (int q = 0; q < PX; ++q) obase[q] = (ox0 + q < p.Wo) ? dview_pix(p.y, b, oy, ox0 + q) : -1;

    for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        float acc[PX][16];
#pragma unroll
        for (int q = 0; q < PX; ++q)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[q][j] = sb[c0 + j];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) {
                    const float4* wr = reinterpret_cast<const float4*>(sw + ((dy * 3 + dx) * CIN + ci) * p.Cout + c0);
                    float w[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 w4 = wr[g];
                        w[g * 4] = w4.x; w[g * 4 + 1] = w4.y; w[g * 4 + 2] = w4.z; w[g * 4 + 3] = w4.w;
                    }
#pragma unroll
                    for (int q = 0; q < PX; ++q) {
                        const float v = in[dy][q * STRIDE + dx][ci];
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[q][j] = fmaf(v, w[j], acc[q][j]);
                    }
                }
#pragma unroll
        for (int q = 0; q < PX; ++q) {
            if (obase[q] < 0) continue;
            if (p.act == YRE_ACT_SILU) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[q][j] = silu_f(acc[q][j]);
            }
            st8<TOut>(p.y.ptr, obase[q] + c0, acc[q]);
            st8<TOut>(p.y.ptr, obase[q] + c0 + 8, acc[q] + 8);
        }
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
    const long long total = (long long)d.B * Ho * ((Wo + PX - 1) / PX);
    const size_t smem = (size_t)(9 * d.Cin + 1) * d.y.C * sizeof(float);
    if (smem > 48 * 1024) YRE_FAIL(YRE_EUNSUPPORTED, "stem: Cout too large for the weight cache");
    dim3 grid(yre_cdiv(total, 128));
#define STEM_GO(T, C) do { if (d.stride == 2) stem_kernel<T, C, 2><<<grid, 128, smem, s>>>(p); else stem_kernel<T, C, 1><<<grid, 128, smem, s>>>(p); } while (0)
#define STEM_T(T) switch (d.Cin) { case 1: STEM_GO(T, 1); break; case 2: STEM_GO(T, 2); break; case 3: STEM_GO(T, 3); break; default: STEM_GO(T, 4); }
    if (d.y.dtype == YRE_F32) { STEM_T(float) } else { STEM_T(__nv_bfloat16) }
#undef STEM_T
#undef STEM_GO
    YRE_LAUNCH_CHECK("stem");
    return YRE_OK;
}
