// K6: DFL softmax-integral box decode + anchor/stride generation + sigmoid class scoring.
//
//   reference: DetectDFL.forward tail   src/yolo/heads/detect.py:93-108
//              DFL.forward              src/yolo/heads/dfl.py:46-50
//              make_anchors/dist2bbox   src/yolo/heads/anchor.py:26-40, 57-64
//
// HBM-bound: per anchor it reads (64+nc) logits and writes 4+nc floats, nothing is reused.
// A CTA stages TILE anchors: the raw rows are fetched with fully coalesced 16-byte loads into
// shared memory, then one thread per (anchor, box side) does the 16-bin softmax expectation,
// sigmoid runs over the class logits, and the [TILE][4+nc] result leaves as one contiguous,
// coalesced block.  Anchor coordinates come from the anchor's linear index (no anchor tensor,
// no host sync).
#include "yre_common.cuh"

namespace {

constexpr int TILE = 32;
constexpr int MAXL = 8;

struct DecodeParams {
    DView raw[MAXL];
    float stride[MAXL];
    int a_start[MAXL + 1];   // first anchor of every level
    int levels, nc, A, B;
    float dfl_w[16];
    float* y;
};

template <typename T>
__global__ void __launch_bounds__(128) decode_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) float sm[];
    const int CH = 64 + p.nc;               // floats per raw row
    const int PITCH = CH + 4;               // keeps 16B alignment, skews banks
    const int OC = 4 + p.nc;
    float* s_raw = sm;                      // [TILE][PITCH]
    float* s_e = s_raw + TILE * PITCH;      // [TILE][4]
    float* s_out = s_e + TILE * 4;          // [TILE][OC]

    const int tiles_per_img = (p.A + TILE - 1) / TILE;
    const int b = blockIdx.x / tiles_per_img;
    const int a0 = (blockIdx.x % tiles_per_img) * TILE;
    const int na = min(TILE, p.A - a0);
    const int tid = threadIdx.x;

    // ---- per-anchor source offset (level lookup + div/mod once per anchor, not once per load) ----
    __shared__ long long s_src[TILE];
    __shared__ float s_ax[TILE], s_ay[TILE], s_st[TILE];
    if (tid < na) {
        const int a = a0 + tid;
        int l = 0;
        while (l + 1 < p.levels && a >= p.a_start[l + 1]) ++l;
        const int r = a - p.a_start[l];
        const DView& v = p.raw[l];
        const int py = r / v.W, px = r - py * v.W;
        s_src[tid] = dview_pix(v, b, py, px) | ((long long)l << 56);
        s_ax[tid] = (float)px + 0.5f; s_ay[tid] = (float)py + 0.5f; s_st[tid] = p.stride[l];
    }
    __syncthreads();
    // ---- stage raw rows: warp w takes anchors w, w+4, ...; a row (CH floats) is one contiguous burst ----
    const int chunks = CH / 4;
    const int wid = tid >> 5, lane = tid & 31;
    for (int al = wid; al < na; al += 4) {
        const long long src = s_src[al];
        const void* base = p.raw[(int)(src >> 56)].ptr;
        const long long off = src & 0x00ffffffffffffffll;
        for (int ck = lane; ck < chunks; ck += 32)
            *reinterpret_cast<float4*>(s_raw + al * PITCH + ck * 4) = ld4<T>(base, off + ck * 4);
    }
    __syncthreads();

    // ---- DFL expectation: thread = (anchor, side) ----
    {
        const int al = tid >> 2, side = tid & 3;
        if (al < na) {
            const float* z = s_raw + al * PITCH + side * 16;
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(z + q * 4);
                v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
            }
            float mx = v[0];
#pragma unroll
            for (int k = 1; k < 16; ++k) mx = fmaxf(mx, v[k]);
            float den = 0.f, num = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float e = __expf(v[k] - mx);
                den += e;
                num = fmaf(e, p.dfl_w[k], num);
            }
            s_e[al * 4 + side] = num / den;
        }
    }
    __syncthreads();

    // ---- boxes: thread = anchor ----
    if (tid < na) {
        const float ax = s_ax[tid], ay = s_ay[tid], st = s_st[tid];
        const float x1 = ax - s_e[tid * 4 + 0], y1 = ay - s_e[tid * 4 + 1];
        const float x2 = ax + s_e[tid * 4 + 2], y2 = ay + s_e[tid * 4 + 3];
        float* o = s_out + tid * OC;
        o[0] = ((x1 + x2) / 2.f) * st;
        o[1] = ((y1 + y2) / 2.f) * st;
        o[2] = (x2 - x1) * st;
        o[3] = (y2 - y1) * st;
    }
    // ---- class scores: warp per anchor, lanes stride the classes (no div/mod) ----
    for (int al = wid; al < na; al += 4)
        for (int c = lane; c < p.nc; c += 32)
            s_out[al * OC + 4 + c] = __frcp_rn(1.0f + __expf(-s_raw[al * PITCH + 64 + c]));
    __syncthreads();

    // ---- contiguous coalesced store of na*OC floats ----
    float* dst = p.y + ((long long)b * p.A + a0) * OC;
    const int nfl = na * OC;
    if ((OC & 3) == 0) {
        for (int i = tid; i < nfl / 4; i += blockDim.x)
            reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_out)[i];
    } else {
        for (int i = tid; i < nfl; i += blockDim.x) dst[i] = s_out[i];
    }
}

}  // namespace

int launch_decode(const yre_decode_desc& d, cudaStream_t s) {
    if (d.levels < 1 || d.levels > MAXL) YRE_FAIL(YRE_EINVAL, "decode: levels=%d", d.levels);
    if (!d.y) YRE_FAIL(YRE_EINVAL, "decode: null pointer");
    if (d.nc < 1 || d.nc % 4) YRE_FAIL(YRE_EUNSUPPORTED, "decode: nc=%d must be a positive multiple of 4", d.nc);
    DecodeParams p;
    p.levels = d.levels; p.nc = d.nc; p.y = d.y;
    int A = 0;
    for (int l = 0; l < d.levels; ++l) {
        const yre_view& v = d.raw[l];
        if (yre_check_view(&v, "decode.raw")) return YRE_EINVAL;
        if (v.layout != YRE_NHWC || v.C != 64 + d.nc || v.c_off % 4 || v.C_total % 4 || v.dtype != d.raw[0].dtype || v.B != d.raw[0].B)
            YRE_FAIL(YRE_EINVAL, "decode: level %d must be an NHWC [B,H,W,%d] view", l, 64 + d.nc);
        p.raw[l] = make_dview(v);
        p.stride[l] = d.stride[l];
        p.a_start[l] = A;
        A += v.H * v.W;
    }
    p.a_start[d.levels] = A;
    p.A = A; p.B = d.raw[0].B;
    for (int k = 0; k < 16; ++k) p.dfl_w[k] = d.dfl_w[k];
    const int CH = 64 + d.nc, OC = 4 + d.nc;
    const size_t smem = (size_t)(TILE * (CH + 4) + TILE * 4 + TILE * OC) * sizeof(float);
    const int tiles = (A + TILE - 1) / TILE;
    dim3 grid((unsigned)(tiles * p.B));
    if (smem > 200 * 1024) YRE_FAIL(YRE_EUNSUPPORTED, "decode: nc=%d too large for the staging tile", d.nc);
    if (smem > 48 * 1024) {
        YRE_CUDA(cudaFuncSetAttribute(decode_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        YRE_CUDA(cudaFuncSetAttribute(decode_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (d.raw[0].dtype == YRE_F32) decode_kernel<float><<<grid, 128, smem, s>>>(p);
    else decode_kernel<__nv_bfloat16><<<grid, 128, smem, s>>>(p);
    YRE_LAUNCH_CHECK("dfl_decode_score");
    return YRE_OK;
}
