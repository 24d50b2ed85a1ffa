// K3 (ADown pre-pool), K4 (SPPELAN pyramid), K5 (nearest x2 upsample), K8 (CBFuse sum) and the
// NCHW<->channels-last conversions at the API boundary.  All are HBM-bound: one thread moves
// 8 consecutive channels (16 B of bf16 / 32 B of fp32) so that a warp touches whole 128 B lines.
#include "yre_common.cuh"
#include <float.h>

namespace {

// -------------------------------------------------------------------------------------------------
// K3: avg_pool2d(x,2,1,0) then, per channel half, [identity | max_pool2d(3,2,1)]
//     reference: src/yolo/blocks/downsample.py:41-44
// One thread per (output cell of the stride-2 grid, 8-channel group).  For the low channel half
// it emits the four average-map pixels (2oy+py, 2ox+px) -- i.e. one cell of each parity plane;
// for the high half it emits max over average rows/cols 2o-1..2o+1.
template <typename T>
__global__ void __launch_bounds__(256, 2) adown_prepool_kernel(DView x, DView lo, DView hi, int Ho, int Wo, int half) {
    // blockIdx.y selects the channel half, so the (very different) low/high paths never share a warp
    const int hgroups = half / 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)x.B * Ho * Wo * hgroups;
    if (idx >= total) return;
    const int g = (int)(idx % hgroups);
    long long t = idx / hgroups;
    const int ox = (int)(t % Wo); t /= Wo;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const int c = g * 8 + (blockIdx.y ? half : 0);
    const int Ha = x.H - 1, Wa = x.W - 1;          // average-map extent

    if (c < half) {
        // x rows 2oy..2oy+2, cols 2ox..2ox+2, streamed two rows at a time
        float ra[3][8], rb[3][8];
        auto load_row = [&](int iy, float (*r)[8]) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int ix = 2 * ox + q;
                if (iy < x.H && ix < x.W) ld8<T>(x.ptr, dview_pix(x, b, iy, ix) + c, r[q]);
                else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) r[q][e] = 0.f;
                }
            }
        };
        auto emit = [&](int py, float (*top)[8], float (*bot)[8]) {
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                const int ay = 2 * oy + py, ax = 2 * ox + px;
                float o[8];
                const bool ok = ay < Ha && ax < Wa;
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    o[e] = ok ? (((top[px][e] + top[px + 1][e]) + bot[px][e]) + bot[px + 1][e]) * 0.25f : 0.f;
                if (lo.layout == YRE_PHASE4) {
                    // every cell of every parity plane is written (zeros where no source pixel exists)
                    if ((ay >> 1) < lo.Hp && (ax >> 1) < lo.Wp) st8<T>(lo.ptr, dview_pix(lo, b, ay, ax) + c, o);
                } else if (ok) {
                    st8<T>(lo.ptr, dview_pix(lo, b, ay, ax) + c, o);
                }
            }
        };
        load_row(2 * oy, ra);
        load_row(2 * oy + 1, rb);
        emit(0, ra, rb);
        load_row(2 * oy + 2, ra);
        emit(1, rb, ra);
    } else {
        float m[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = -FLT_MAX;
        // average rows 2oy-1..2oy+1 need x rows 2oy-1..2oy+2, x cols 2ox-1..2ox+2
        float prev[4][8];
        bool have_prev = false;
        for (int r = 0; r < 4; ++r) {
            const int iy = 2 * oy - 1 + r;
            const bool rowok = iy >= 0 && iy < x.H;
            float cur[4][8];
            if (rowok) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int ix = 2 * ox - 1 + q;
                    if (ix >= 0 && ix < x.W) ld8<T>(x.ptr, dview_pix(x, b, iy, ix) + c, cur[q]);
                    else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) cur[q][e] = 0.f;
                    }
                }
            }
            // average row ay = iy-1 is built from x rows ay (prev) and ay+1 (cur)
            if (have_prev && rowok) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int ax = 2 * ox - 1 + q;
                    if (ax >= 0 && ax < Wa) {
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            m[e] = fmaxf(m[e], (((prev[q][e] + prev[q + 1][e]) + cur[q][e]) + cur[q + 1][e]) * 0.25f);
                    }
                }
            }
            if (rowok) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int e = 0; e < 8; ++e) prev[q][e] = cur[q][e];
            }
            have_prev = rowok;
        }
        st8<T>(hi.ptr, dview_pix(hi, b, oy, ox) + (c - half), m);
    }
}

// -------------------------------------------------------------------------------------------------
// K4: windows 5/9/13 max (== three chained MaxPool2d(5,1,2); reference sppelan.py:44-47)
template <typename T>
__global__ void __launch_bounds__(256) spp_kernel(DView x, DView y5, DView y9, DView y13) {
    const int groups = x.C / 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)x.B * x.H * x.W * groups;
    if (idx >= total) return;
    const int c = (int)(idx % groups) * 8;
    long long t = idx / groups;
    const int px = (int)(t % x.W); t /= x.W;
    const int py = (int)(t % x.H);
    const int b = (int)(t / x.H);
    float m5[8], m9[8], m13[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m5[e] = m9[e] = m13[e] = -FLT_MAX;
    for (int dy = -6; dy <= 6; ++dy) {
        const int iy = py + dy;
        if (iy < 0 || iy >= x.H) continue;
        const int ady = dy < 0 ? -dy : dy;
        for (int dx = -6; dx <= 6; ++dx) {
            const int ix = px + dx;
            if (ix < 0 || ix >= x.W) continue;
            const int adx = dx < 0 ? -dx : dx;
            const int r = ady > adx ? ady : adx;
            float v[8];
            ld8<T>(x.ptr, dview_pix(x, b, iy, ix) + c, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                m13[e] = fmaxf(m13[e], v[e]);
                if (r <= 4) m9[e] = fmaxf(m9[e], v[e]);
                if (r <= 2) m5[e] = fmaxf(m5[e], v[e]);
            }
        }
    }
    st8<T>(y5.ptr, dview_pix(y5, b, py, px) + c, m5);
    st8<T>(y9.ptr, dview_pix(y9, b, py, px) + c, m9);
    st8<T>(y13.ptr, dview_pix(y13, b, py, px) + c, m13);
}

// -------------------------------------------------------------------------------------------------
// K5: nearest x2 (reference: nn.Upsample, parser.py:159-171) written into the consumer's slice
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_kernel(DView x, DView y) {
    const int groups = x.C / 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)y.B * y.H * y.W * groups;
    if (idx >= total) return;
    const int c = (int)(idx % groups) * 8;
    long long t = idx / groups;
    const int ox = (int)(t % y.W); t /= y.W;
    const int oy = (int)(t % y.H);
    const int b = (int)(t / y.H);
    float v[8];
    ld8<T>(x.ptr, dview_pix(x, b, oy >> 1, ox >> 1) + c, v);
    st8<T>(y.ptr, dview_pix(y, b, oy, ox) + c, v);
}

// -------------------------------------------------------------------------------------------------
// K8: CBFuse (reference auxiliary.py:100-110): nearest-resize every source to the target size,
// stack with the target and sum in that order.
struct FuseSrcs { DView v[8]; int n; };
template <typename T>
__global__ void __launch_bounds__(256) cbfuse_kernel(FuseSrcs srcs, DView tgt, DView y) {
    const int groups = y.C / 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)y.B * y.H * y.W * groups;
    if (idx >= total) return;
    const int c = (int)(idx % groups) * 8;
    long long t = idx / groups;
    const int ox = (int)(t % y.W); t /= y.W;
    const int oy = (int)(t % y.H);
    const int b = (int)(t / y.H);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int i = 0; i < srcs.n; ++i) {
        const DView& sv = srcs.v[i];
        // F.interpolate(mode="nearest"): src = floor(dst * in/out)
        const int sy = min((int)floorf(oy * ((float)sv.H / y.H)), sv.H - 1);
        const int sx = min((int)floorf(ox * ((float)sv.W / y.W)), sv.W - 1);
        float v[8];
        ld8<T>(sv.ptr, dview_pix(sv, b, sy, sx) + c, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
    float v[8];
    ld8<T>(tgt.ptr, dview_pix(tgt, b, oy, ox) + c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += v[e];
    st8<T>(y.ptr, dview_pix(y, b, oy, ox) + c, acc);
}

// -------------------------------------------------------------------------------------------------
// NCHW fp32 <-> channels-last view, through a 32x32 shared-memory transpose (channels x pixels)
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_view_kernel(const float* __restrict__ x, DView y) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const long long HW = (long long)y.H * y.W;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 rows per pass
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        const long long p = p0 + tx;
        tile[r][tx] = (c < y.C && p < HW) ? x[((long long)b * y.C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long p = p0 + r;
        const int c = c0 + tx;
        if (p < HW && c < y.C) {
            const int py = (int)(p / y.W), px = (int)(p % y.W);
            Elt<T>::st(y.ptr, dview_pix(y, b, py, px) + c, tile[tx][r]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) view_to_nchw_kernel(DView x, float* __restrict__ y) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const long long HW = (long long)x.H * x.W;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const long long p = p0 + r;
        const int c = c0 + tx;
        float v = 0.f;
        if (p < HW && c < x.C) {
            const int py = (int)(p / x.W), px = (int)(p % x.W);
            v = Elt<T>::ld(x.ptr, dview_pix(x, b, py, px) + c);
        }
        tile[r][tx] = v;     // [pixel][channel]
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        const long long p = p0 + tx;
        if (c < x.C && p < HW) y[((long long)b * x.C + c) * HW + p] = tile[tx][r];
    }
}

int check8(const yre_view& v, const char* what) {
    if (yre_check_view(&v, what)) return YRE_EINVAL;
    if (v.C % 8 || v.c_off % 8 || v.C_total % 8) YRE_FAIL(YRE_EUNSUPPORTED, "%s: channels must be multiples of 8", what);
    return YRE_OK;
}

}  // namespace

int launch_adown_prepool(const yre_view& x, const yre_view& lo, const yre_view& hi, cudaStream_t s) {
    if (check8(x, "adown.x") || check8(lo, "adown.avg_lo") || check8(hi, "adown.max_hi")) return YRE_EINVAL;
    if (x.layout != YRE_NHWC || hi.layout != YRE_NHWC) YRE_FAIL(YRE_EUNSUPPORTED, "adown: x and max_hi must be NHWC");
    if (x.C % 16) YRE_FAIL(YRE_EUNSUPPORTED, "adown: C must be a multiple of 16");
    const int half = x.C / 2;
    const int Ho = (x.H - 1 + 2 - 3) / 2 + 1, Wo = (x.W - 1 + 2 - 3) / 2 + 1;
    if (lo.C != half || hi.C != half || lo.H != x.H - 1 || lo.W != x.W - 1 || hi.H != Ho || hi.W != Wo ||
        lo.B != x.B || hi.B != x.B || lo.dtype != x.dtype || hi.dtype != x.dtype)
        YRE_FAIL(YRE_EINVAL, "adown: output views do not match the input (want avg %dx%d C%d, max %dx%d)", x.H - 1, x.W - 1, half, Ho, Wo);
    // the thread grid walks ceil((H-1)/2) x ceil((W-1)/2) cells, which covers both outputs
    const int Gy = (x.H - 1 + 1) / 2, Gx = (x.W - 1 + 1) / 2;
    if (Gy < Ho || Gx < Wo) YRE_FAIL(YRE_EINVAL, "adown: grid smaller than the pooled output");
    const long long total = (long long)x.B * Gy * Gx * (half / 8);
    dim3 grid(yre_cdiv(total, 256), 2);
    // NB: kernel indexes its cell grid with (Ho,Wo) = (Gy,Gx); cells beyond hi's extent never
    // occur because Gy==Ho, Gx==Wo for every H,W >= 2.
    if (Gy != Ho || Gx != Wo) YRE_FAIL(YRE_EUNSUPPORTED, "adown: H=%d W=%d", x.H, x.W);
    if (x.dtype == YRE_F32) adown_prepool_kernel<float><<<grid, 256, 0, s>>>(make_dview(x), make_dview(lo), make_dview(hi), Gy, Gx, half);
    else adown_prepool_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(make_dview(x), make_dview(lo), make_dview(hi), Gy, Gx, half);
    YRE_LAUNCH_CHECK("adown_prepool");
    return YRE_OK;
}

int launch_spp_maxpool(const yre_view& x, const yre_view& y5, const yre_view& y9, const yre_view& y13, cudaStream_t s) {
    if (check8(x, "spp.x") || check8(y5, "spp.y5") || check8(y9, "spp.y9") || check8(y13, "spp.y13")) return YRE_EINVAL;
    const yre_view* ys[3] = {&y5, &y9, &y13};
    for (auto* y : ys)
        if (y->B != x.B || y->H != x.H || y->W != x.W || y->C != x.C || y->dtype != x.dtype || y->layout != YRE_NHWC)
            YRE_FAIL(YRE_EINVAL, "spp: output views must match the input");
    if (x.layout != YRE_NHWC) YRE_FAIL(YRE_EUNSUPPORTED, "spp: NHWC only");
    const long long total = (long long)x.B * x.H * x.W * (x.C / 8);
    dim3 grid(yre_cdiv(total, 256));
    if (x.dtype == YRE_F32) spp_kernel<float><<<grid, 256, 0, s>>>(make_dview(x), make_dview(y5), make_dview(y9), make_dview(y13));
    else spp_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(make_dview(x), make_dview(y5), make_dview(y9), make_dview(y13));
    YRE_LAUNCH_CHECK("spp_maxpool");
    return YRE_OK;
}

int launch_upsample2x(const yre_view& x, const yre_view& y, cudaStream_t s) {
    if (check8(x, "upsample.x") || check8(y, "upsample.y")) return YRE_EINVAL;
    if (y.B != x.B || y.H != 2 * x.H || y.W != 2 * x.W || y.C != x.C || y.dtype != x.dtype)
        YRE_FAIL(YRE_EINVAL, "upsample2x: output must be 2x the input");
    const long long total = (long long)y.B * y.H * y.W * (y.C / 8);
    dim3 grid(yre_cdiv(total, 256));
    if (x.dtype == YRE_F32) upsample2x_kernel<float><<<grid, 256, 0, s>>>(make_dview(x), make_dview(y));
    else upsample2x_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(make_dview(x), make_dview(y));
    YRE_LAUNCH_CHECK("upsample2x");
    return YRE_OK;
}

int launch_cbfuse_sum(const yre_view* srcs, int n, const yre_view& tgt, const yre_view& y, cudaStream_t s) {
    if (n < 0 || n > 8) YRE_FAIL(YRE_EUNSUPPORTED, "cbfuse: %d sources (max 8)", n);
    if (check8(tgt, "cbfuse.target") || check8(y, "cbfuse.y")) return YRE_EINVAL;
    FuseSrcs fs; fs.n = n;
    for (int i = 0; i < n; ++i) {
        if (check8(srcs[i], "cbfuse.src")) return YRE_EINVAL;
        if (srcs[i].C != y.C || srcs[i].B != y.B || srcs[i].dtype != y.dtype) YRE_FAIL(YRE_EINVAL, "cbfuse: source %d shape/dtype mismatch", i);
        fs.v[i] = make_dview(srcs[i]);
    }
    if (tgt.B != y.B || tgt.H != y.H || tgt.W != y.W || tgt.C != y.C || tgt.dtype != y.dtype) YRE_FAIL(YRE_EINVAL, "cbfuse: target mismatch");
    const long long total = (long long)y.B * y.H * y.W * (y.C / 8);
    dim3 grid(yre_cdiv(total, 256));
    if (y.dtype == YRE_F32) cbfuse_kernel<float><<<grid, 256, 0, s>>>(fs, make_dview(tgt), make_dview(y));
    else cbfuse_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(fs, make_dview(tgt), make_dview(y));
    YRE_LAUNCH_CHECK("cbfuse_sum");
    return YRE_OK;
}

int launch_nchw_to_view(const float* x, const yre_view& y, cudaStream_t s) {
    if (!x) YRE_FAIL(YRE_EINVAL, "nchw_to_view: null input");
    if (yre_check_view(&y, "nchw_to_view.y")) return YRE_EINVAL;
    dim3 grid(yre_cdiv((long long)y.H * y.W, 32), yre_cdiv(y.C, 32), y.B);
    if (y.dtype == YRE_F32) nchw_to_view_kernel<float><<<grid, 256, 0, s>>>(x, make_dview(y));
    else nchw_to_view_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(x, make_dview(y));
    YRE_LAUNCH_CHECK("nchw_to_view");
    return YRE_OK;
}

int launch_view_to_nchw(const yre_view& x, float* y, cudaStream_t s) {
    if (!y) YRE_FAIL(YRE_EINVAL, "view_to_nchw: null output");
    if (yre_check_view(&x, "view_to_nchw.x")) return YRE_EINVAL;
    dim3 grid(yre_cdiv((long long)x.H * x.W, 32), yre_cdiv(x.C, 32), x.B);
    if (x.dtype == YRE_F32) view_to_nchw_kernel<float><<<grid, 256, 0, s>>>(make_dview(x), y);
    else view_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(make_dview(x), y);
    YRE_LAUNCH_CHECK("view_to_nchw");
    return YRE_OK;
}
