// K7: batched class-aware NMS, bit-exact with the reference.
//
//   reference: non_max_suppression           src/yolo/utils/nms.py:19-94
//              torchvision.ops.nms via _nms   src/yolo/utils/nms.py:97-104
//
// Three launches for the whole batch (the reference loops over images in Python):
//   nms_init    zero the per-image counters
//   nms_filter  HBM-bound pass over pred[B][A][4+nc]: per-anchor first-max class, strict
//               `conf > thr` (+ optional class filter), xywh->xyxy, per-image max coordinate,
//               compaction of (score, anchor) sort keys
//   nms_select  one CTA (1024 threads) per image: bitonic sort of the 64-bit keys (== stable descending
//               score sort, ties -> lower anchor first) -- in shared memory; above 8192 candidates the
//               network is run segment-wise (8192-key segments sorted / merged in shared memory, only the
//               few stages with a compare distance >= 8192 touch global memory) -- then a chunked greedy
//               scan: every chunk of CHUNK sorted candidates is tested against the boxes kept so far
//               (<= max_det of them, two threads per candidate), an IoU bitmask is built inside the chunk
//               (32 warps), and one thread resolves the chunk by walking the set bits of the survivors.
//               The scan stops as soon as max_det boxes are kept, which is exactly `keep[:max_det]` of the
//               reference -- there is no pre-NMS top-k cap.
//
// Bit-exactness notes: every fp32 op is an explicitly rounded intrinsic (no FMA contraction);
// the class offset is `float(cls) * (max_coord + 1)` then `box + offset` as two rounded ops
// (nms.py:79-81); IoU is inter / ((area_i + area_j) - inter) in fp32, compared against the
// threshold as a DOUBLE, as torchvision's CPU kernel does.
#include "yre_common.cuh"

namespace {

constexpr int CHUNK = 512;            // candidates per greedy round
constexpr int NT_SEL = 1024;          // threads of nms_select (two per candidate of a chunk)
constexpr int CW = CHUNK / 64;        // mask words per row
constexpr int SMEM_KEYS = 8192;       // keys sorted in shared memory at a time (one segment)

typedef unsigned long long u64;

__device__ __forceinline__ unsigned f2ord(float f) {   // order-preserving float -> uint
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

struct NmsWs {           // carved out of the caller's workspace
    u64* keys;           // [B][cap]
    int* cls;            // [B][A]
    int* count;          // [B]
    unsigned* maxc;      // [B] (ordered-uint encoding)
    int cap;
};

__host__ __device__ inline int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

__global__ void nms_init_kernel(int* count, unsigned* maxc, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) { count[i] = 0; maxc[i] = 0u; }
}

// ---------------------------------------------------------------------------------------------
// filter: one warp per 32 consecutive anchors of one image; rows are read as coalesced float4
struct FilterParams {
    const float* pred; int B, A, nc; float conf; const int* classes; int n_classes;
    NmsWs ws;
};

__global__ void __launch_bounds__(128) nms_filter_kernel(const FilterParams p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n4 = (4 + p.nc) / 4;                 // float4 per row (nc % 4 == 0)
    const int nparts = n4 - 1;
    const int pitch = nparts | 1;                  // odd pitch: conflict-free per-lane rows
    float* pv = reinterpret_cast<float*>(smraw) + (size_t)warp * 32 * pitch * 2;    // [32][pitch] best value
    int* pi = reinterpret_cast<int*>(pv + 32 * pitch);                               // [32][pitch] best class

    const int tiles = (p.A + 31) / 32;
    const int wglobal = blockIdx.x * (blockDim.x >> 5) + warp;
    if (wglobal >= tiles * p.B) return;
    const int b = wglobal / tiles, a0 = (wglobal % tiles) * 32;
    const int na = min(32, p.A - a0);
    const float4* base = reinterpret_cast<const float4*>(p.pred + ((size_t)b * p.A + a0) * (4 + p.nc));

    const bool vec = (p.nc & 3) == 0;        // rows are 16-byte aligned only when nc % 4 == 0
    // loads are issued 8 at a time into registers before anything is consumed (no load -> use -> load chain)
    for (int base0 = 0; vec && base0 < na * n4; base0 += 8 * 32) {
        float4 buf[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int idx = base0 + q * 32 + lane;
            if (idx < na * n4) buf[q] = base[idx];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int idx = base0 + q * 32 + lane;
            if (idx >= na * n4) continue;
            const int al = idx / n4, pos = idx - al * n4;
            if (pos == 0) continue;
            const float4 v = buf[q];
            float best = v.x; int bi = 0;
            if (v.y > best) { best = v.y; bi = 1; }
            if (v.z > best) { best = v.z; bi = 2; }
            if (v.w > best) { best = v.w; bi = 3; }
            pv[al * pitch + pos - 1] = best;
            pi[al * pitch + pos - 1] = (pos - 1) * 4 + bi;
        }
    }
    __syncwarp();

    bool cand = false;
    float conf = 0.f; int cls = 0;
    float x1 = 0, y1 = 0, x2 = 0, y2 = 0;
    if (lane < na) {
        if (vec) {
            conf = pv[lane * pitch]; cls = pi[lane * pitch];
            for (int q = 1; q < nparts; ++q) {
                const float v = pv[lane * pitch + q];
                if (v > conf) { conf = v; cls = pi[lane * pitch + q]; }     // strict > keeps the first max
            }
        } else {                              // generic class count: one lane walks its own row
            const float* row = p.pred + ((size_t)b * p.A + a0 + lane) * (4 + p.nc);
            conf = row[4];
            for (int q = 1; q < p.nc; ++q) {
                const float v = row[4 + q];
                if (v > conf) { conf = v; cls = q; }
            }
        }
        cand = conf > p.conf;
        if (cand && p.n_classes >= 0) {
            bool ok = false;
            for (int k = 0; k < p.n_classes; ++k) ok |= (p.classes[k] == cls);
            cand = ok;
        }
        if (cand) {
            const float* brow = p.pred + ((size_t)b * p.A + a0 + lane) * (4 + p.nc);
            const float4 bx = vec ? base[lane * n4] : make_float4(brow[0], brow[1], brow[2], brow[3]);
            const float hw = __fmul_rn(bx.z, 0.5f), hh = __fmul_rn(bx.w, 0.5f);   // w/2, h/2 (exact)
            x1 = __fsub_rn(bx.x, hw); y1 = __fsub_rn(bx.y, hh);
            x2 = __fadd_rn(bx.x, hw); y2 = __fadd_rn(bx.y, hh);
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, cand);
    if (ballot == 0u) return;
    // per-image max coordinate over candidate boxes
    unsigned mo = cand ? max(max(f2ord(x1), f2ord(y1)), max(f2ord(x2), f2ord(y2))) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mo = max(mo, __shfl_xor_sync(0xffffffffu, mo, o));
    int slot0 = 0;
    if (lane == 0) {
        atomicMax(p.ws.maxc + b, mo);
        slot0 = atomicAdd(p.ws.count + b, __popc(ballot));
    }
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    if (cand) {
        const int slot = slot0 + __popc(ballot & ((1u << lane) - 1u));
        const int a = a0 + lane;
        p.ws.keys[(size_t)b * p.ws.cap + slot] = ((u64)(~f2ord(conf)) << 32) | (unsigned)a;
        p.ws.cls[(size_t)b * p.A + a] = cls;
    }
}

// ---------------------------------------------------------------------------------------------
struct SelectParams {
    const float* pred; int B, A, nc; double iou; int max_det, agnostic;
    NmsWs ws;
    float* out; int* counts; long long* keep_anchor;
    int keys_in_smem;     // SMEM_KEYS or 0
    const float* scale;   // optional [B][5] = pad_w, pad_h, gain, orig_w, orig_h (scale_boxes fused into the output)
};

__device__ __forceinline__ bool iou_gt(const float4 a, const float aa, const float4 b, const float ab, const double thr) {
    const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    const float inter = __fmul_rn(w, h);
    // inter == 0 -> IoU is +-0 or NaN, neither exceeds a threshold >= 0: skip the IEEE division (exact shortcut)
    if (inter == 0.f && thr >= 0.0) return false;
    const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
    return (double)iou > thr;
}

// bitonic network stages of merge size k with compare distances j = j_hi, j_hi/2, ..., 1 on the np_local keys in d
// (shared memory); `base` = global index of d[0] (the sort direction of an element depends on its GLOBAL index)
__device__ __forceinline__ void bitonic_local(u64* d, int np_local, int base, int k, int j_hi) {
    for (int j = j_hi; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < (np_local >> 1); i += blockDim.x) {
            const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
            const int hi = lo | j;
            const bool asc = ((base + lo) & k) == 0;
            const u64 a = d[lo], b = d[hi];
            if ((a > b) == asc) { d[lo] = b; d[hi] = a; }
        }
        __syncthreads();
    }
}

// one stage (merge size k, distance j >= SMEM_KEYS) over np keys in global memory; loads are batched ahead of the
// compare-exchange so that a thread has eight independent pairs in flight instead of one
__device__ __forceinline__ void bitonic_global_stage(u64* d, int np, int k, int j) {
    constexpr int U = 8;
    const int half = np >> 1;
    for (int i0 = threadIdx.x; i0 < half; i0 += blockDim.x * U) {
        u64 a[U], b[U];
        int lo[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * (int)blockDim.x;
            lo[u] = ((i & ~(j - 1)) << 1) | (i & (j - 1));
            if (i < half) { a[u] = d[lo[u]]; b[u] = d[lo[u] | j]; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * (int)blockDim.x;
            if (i < half) {
                const bool asc = (lo[u] & k) == 0;
                if ((a[u] > b[u]) == asc) { d[lo[u]] = b[u]; d[lo[u] | j] = a[u]; }
            }
        }
    }
    __syncthreads();
}

// Ascending sort of the n keys of one image (padded to np = pow2 with ~0).  Returns where the sorted keys live.
__device__ u64* sort_keys(u64* gkeys, int n, int np, u64* skeys, int keys_in_smem) {
    const int tid = threadIdx.x;
    if (np <= keys_in_smem) {
        for (int i = tid; i < np; i += blockDim.x) skeys[i] = i < n ? gkeys[i] : ~0ull;
        __syncthreads();
        for (int k = 2; k <= np; k <<= 1) bitonic_local(skeys, np, 0, k, k >> 1);
        return skeys;
    }
    const int SEG = keys_in_smem;                       // np is a multiple of SEG (both powers of two)
    for (int i = n + tid; i < np; i += blockDim.x) gkeys[i] = ~0ull;
    __syncthreads();
    // 1. every segment: all merge sizes up to SEG in shared memory
    for (int s0 = 0; s0 < np; s0 += SEG) {
        for (int i = tid; i < SEG; i += blockDim.x) skeys[i] = gkeys[s0 + i];
        __syncthreads();
        for (int k = 2; k <= SEG; k <<= 1) bitonic_local(skeys, SEG, s0, k, k >> 1);
        for (int i = tid; i < SEG; i += blockDim.x) gkeys[s0 + i] = skeys[i];
        __syncthreads();
    }
    // 2. larger merge sizes: the long-distance stages in global memory, the rest per segment in shared memory
    for (int k = SEG << 1; k <= np; k <<= 1) {
        for (int j = k >> 1; j >= SEG; j >>= 1) bitonic_global_stage(gkeys, np, k, j);
        for (int s0 = 0; s0 < np; s0 += SEG) {
            for (int i = tid; i < SEG; i += blockDim.x) skeys[i] = gkeys[s0 + i];
            __syncthreads();
            bitonic_local(skeys, SEG, s0, k, SEG >> 1);
            for (int i = tid; i < SEG; i += blockDim.x) gkeys[s0 + i] = skeys[i];
            __syncthreads();
        }
    }
    return gkeys;
}

__global__ void __launch_bounds__(NT_SEL) nms_select_kernel(const SelectParams p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    // carve shared memory
    float4* cbox = reinterpret_cast<float4*>(smraw);                 // [CHUNK] offset boxes
    float4* ubox = cbox + CHUNK;                                     // [CHUNK] plain boxes
    float4* kbox = ubox + CHUNK;                                     // [max_det] kept offset boxes
    u64* mask = reinterpret_cast<u64*>(kbox + p.max_det);            // [CHUNK][CW]
    u64* skeys = mask + (size_t)CHUNK * CW;                          // [keys_in_smem]
    float* carea = reinterpret_cast<float*>(skeys + p.keys_in_smem); // [CHUNK]
    float* cconf = carea + CHUNK;                                    // [CHUNK]
    float* karea = cconf + CHUNK;                                    // [max_det]
    int* ccls = reinterpret_cast<int*>(karea + p.max_det);           // [CHUNK]
    int* canc = ccls + CHUNK;                                        // [CHUNK]
    int* kpos = canc + CHUNK;                                        // [CHUNK] chunk positions kept this round
    unsigned* alive_w = reinterpret_cast<unsigned*>(kpos + CHUNK);   // [CHUNK/32]
    int* dead = reinterpret_cast<int*>(alive_w + CHUNK / 32);        // [CHUNK] suppressed by an earlier chunk's keeps
    __shared__ int s_newkept;

    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = min(p.ws.count[b], p.A);
    if (n == 0) { if (tid == 0) p.counts[b] = 0; return; }
    const int np = pow2ceil(n);
    const u64* keys = sort_keys(p.ws.keys + (size_t)b * p.ws.cap, n, np, skeys, p.keys_in_smem);

    const float scale = __fadd_rn(ord2f(p.ws.maxc[b]), 1.0f);
    const int rowf = 4 + p.nc;
    int kept = 0;
    for (int c0 = 0; c0 < n && kept < p.max_det; c0 += CHUNK) {
        const int cnt = min(CHUNK, n - c0);
        // ---- 1. load this chunk's candidates (threads 0 .. cnt-1) ----
        if (tid < CHUNK) dead[tid] = 0;
        if (tid < cnt) {
            const u64 key = keys[c0 + tid];
            const int a = (int)(unsigned)(key & 0xffffffffu);
            const float conf = ord2f(~(unsigned)(key >> 32));
            const int cls = p.ws.cls[(size_t)b * p.A + a];
            const float* row = p.pred + ((size_t)b * p.A + a) * rowf;
            const float4 bx = (p.nc & 3) ? make_float4(row[0], row[1], row[2], row[3]) : *reinterpret_cast<const float4*>(row);
            const float hw = __fmul_rn(bx.z, 0.5f), hh = __fmul_rn(bx.w, 0.5f);
            float4 u, mine;
            u.x = __fsub_rn(bx.x, hw); u.y = __fsub_rn(bx.y, hh); u.z = __fadd_rn(bx.x, hw); u.w = __fadd_rn(bx.y, hh);
            if (p.agnostic) mine = u;
            else {
                const float off = __fmul_rn(__int2float_rn(cls), scale);
                mine.x = __fadd_rn(u.x, off); mine.y = __fadd_rn(u.y, off);
                mine.z = __fadd_rn(u.z, off); mine.w = __fadd_rn(u.w, off);
            }
            ubox[tid] = u; cbox[tid] = mine;
            carea[tid] = __fmul_rn(__fsub_rn(mine.z, mine.x), __fsub_rn(mine.w, mine.y));
            cconf[tid] = conf; ccls[tid] = cls; canc[tid] = a;
        }
        __syncthreads();
        // ---- 2. against everything kept by earlier chunks: two threads per candidate, each half of the kept list
        //         (whether ANY kept box suppresses the candidate is all that matters, so the split is exact) ----
        {
            const int c = tid & (CHUNK - 1), part = tid >> 9;
            if (c < cnt && kept > 0) {
                const int mid = kept >> 1;
                const int k0 = part ? mid : 0, k1 = part ? kept : mid;
                const float4 mine = cbox[c];
                const float marea = carea[c];
                for (int k = k0; k < k1; ++k)
                    if (iou_gt(kbox[k], karea[k], mine, marea, p.iou)) { dead[c] = 1; break; }
            }
        }
        __syncthreads();
        if (tid < CHUNK) {
            const bool alive = tid < cnt && !dead[tid];
            const unsigned bal = __ballot_sync(0xffffffffu, alive);
            if ((tid & 31) == 0) alive_w[tid >> 5] = bal;
        }
        __syncthreads();
        // ---- 3. in-chunk suppression bitmask (row i: later boxes j>i it would suppress) ----
        // rows are dealt round-robin to the 32 warps; the 32 lanes of a warp test 32 columns at a time and a
        // ballot yields the mask bits -- balanced work and conflict-free shared-memory reads.
        {
            const int wid = tid >> 5, ln = tid & 31;
            for (int i = wid; i < cnt; i += NT_SEL / 32) {
                if (!((alive_w[i >> 5] >> (i & 31)) & 1u)) continue;          // warp-uniform
                const float4 bi = cbox[i];
                const float ai = carea[i];
                for (int w = i >> 6; w < CW; ++w) {
                    unsigned lo_bits = 0u, hi_bits = 0u;
                    if (w * 64 < cnt) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int j = w * 64 + h * 32 + ln;
                            bool sup = false;
                            if (j > i && j < cnt && ((alive_w[j >> 5] >> (j & 31)) & 1u)) sup = iou_gt(bi, ai, cbox[j], carea[j], p.iou);
                            const unsigned bal2 = __ballot_sync(0xffffffffu, sup);
                            if (h == 0) lo_bits = bal2; else hi_bits = bal2;
                        }
                    }
                    if (ln == 0) mask[(size_t)i * CW + w] = (u64)lo_bits | ((u64)hi_bits << 32);
                }
            }
        }
        __syncthreads();
        // ---- 4. serial resolution of the chunk: walk the set bits of (alive & ~removed) only ----
        if (tid == 0) {
            // `removed` lives in registers: the word loop is fully unrolled so every index is static
            u64 removed[CW];
#pragma unroll
            for (int w = 0; w < CW; ++w) removed[w] = 0ull;
            int nk = 0;
            bool full = false;
#pragma unroll
            for (int w = 0; w < CW; ++w) {
                if (full || w * 64 >= cnt) continue;
                u64 cand = ((u64)alive_w[2 * w] | ((u64)alive_w[2 * w + 1] << 32)) & ~removed[w];
                while (cand) {
                    const int jj = __ffsll((long long)cand) - 1;
                    const int i = w * 64 + jj;
                    kpos[nk++] = i;
                    if (kept + nk >= p.max_det) { full = true; break; }
                    const u64* mrow = mask + (size_t)i * CW;
#pragma unroll
                    for (int v = 0; v < CW; ++v)
                        if (v >= w) removed[v] |= mrow[v];
                    cand &= cand - 1;                    // this candidate is done
                    cand &= ~removed[w];                 // and so is everything it (or an earlier keep) suppresses
                }
            }
            s_newkept = nk;
        }
        __syncthreads();
        // ---- 5. publish the newly kept boxes ----
        const int nk = s_newkept;
        for (int q = tid; q < nk; q += blockDim.x) {
            const int i = kpos[q], k = kept + q;
            kbox[k] = cbox[i]; karea[k] = carea[i];
            float* o = p.out + ((size_t)b * p.max_det + k) * 6;
            float4 u = ubox[i];
            if (p.scale) {        // scale_boxes (scripts/detect.py:74-109), same ops as scale_boxes_kernel (k_preproc.cu)
                const float* sc = p.scale + (size_t)b * 5;
                const float pw = sc[0], ph = sc[1], gain = sc[2], ow = sc[3], oh = sc[4];
                u.x = fminf(fmaxf(__fdiv_rn(__fsub_rn(u.x, pw), gain), 0.f), ow);
                u.y = fminf(fmaxf(__fdiv_rn(__fsub_rn(u.y, ph), gain), 0.f), oh);
                u.z = fminf(fmaxf(__fdiv_rn(__fsub_rn(u.z, pw), gain), 0.f), ow);
                u.w = fminf(fmaxf(__fdiv_rn(__fsub_rn(u.w, ph), gain), 0.f), oh);
            }
            o[0] = u.x; o[1] = u.y; o[2] = u.z; o[3] = u.w; o[4] = cconf[i]; o[5] = __int2float_rn(ccls[i]);
            p.keep_anchor[(size_t)b * p.max_det + k] = canc[i];
        }
        kept += nk;
        __syncthreads();
    }
    if (tid == 0) p.counts[b] = kept;
}

size_t select_smem_bytes(int max_det, int keys_in_smem) {
    return sizeof(float4) * (2 * CHUNK + (size_t)max_det) + sizeof(u64) * ((size_t)CHUNK * CW + keys_in_smem) +
           sizeof(float) * (2 * CHUNK + (size_t)max_det) + sizeof(int) * (4 * CHUNK) + sizeof(unsigned) * (CHUNK / 32) + 64;
}

NmsWs carve(void* ws, int B, int A) {
    NmsWs w;
    w.cap = pow2ceil(A);
    unsigned char* p = reinterpret_cast<unsigned char*>(ws);
    w.keys = reinterpret_cast<u64*>(p); p += sizeof(u64) * (size_t)B * w.cap;
    w.cls = reinterpret_cast<int*>(p); p += sizeof(int) * (size_t)B * A;
    w.count = reinterpret_cast<int*>(p); p += sizeof(int) * (size_t)B;
    w.maxc = reinterpret_cast<unsigned*>(p);
    return w;
}

}  // namespace

extern "C" size_t yre_nms_workspace_bytes(int32_t B, int32_t A) {
    if (B <= 0 || A <= 0) return 0;
    return sizeof(u64) * (size_t)B * pow2ceil(A) + sizeof(int) * (size_t)B * A + 2 * sizeof(int) * (size_t)B + 256;
}

static int check_nms(const yre_nms_desc& d, bool outputs) {
    if (!d.pred || !d.workspace) YRE_FAIL(YRE_EINVAL, "nms: null pointer");
    if (outputs && (!d.out || !d.counts || !d.keep_anchor)) YRE_FAIL(YRE_EINVAL, "nms: null output pointer");
    if (d.B <= 0 || d.A <= 0 || d.nc <= 0) YRE_FAIL(YRE_EINVAL, "nms: bad extent B=%d A=%d nc=%d", d.B, d.A, d.nc);
    if (d.max_det <= 0 || d.max_det > 4096) YRE_FAIL(YRE_EUNSUPPORTED, "nms: max_det=%d (supported 1..4096)", d.max_det);
    if (d.workspace_bytes < yre_nms_workspace_bytes(d.B, d.A)) YRE_FAIL(YRE_EINVAL, "nms: workspace too small");
    if (d.n_classes > 0 && !d.classes) YRE_FAIL(YRE_EINVAL, "nms: classes pointer missing");
    if ((reinterpret_cast<uintptr_t>(d.pred) & 15) || (reinterpret_cast<uintptr_t>(d.workspace) & 15))
        YRE_FAIL(YRE_EINVAL, "nms: pred/workspace must be 16-byte aligned");
    return YRE_OK;
}

static int launch_nms_filter(const yre_nms_desc& d, const NmsWs& ws, cudaStream_t s) {
    nms_init_kernel<<<yre_cdiv(d.B, 256), 256, 0, s>>>(ws.count, ws.maxc, d.B);
    YRE_LAUNCH_CHECK("nms_init");

    FilterParams fp;
    fp.pred = d.pred; fp.B = d.B; fp.A = d.A; fp.nc = d.nc; fp.conf = d.conf_thres;
    fp.classes = d.classes; fp.n_classes = d.n_classes; fp.ws = ws;
    const size_t per_warp = (d.nc % 4 == 0) ? (size_t)32 * ((d.nc / 4) | 1) * 8 : 16;
    int warps = 4;
    while (warps > 1 && per_warp * warps > 40 * 1024) warps >>= 1;
    if (per_warp * warps > 48 * 1024) YRE_FAIL(YRE_EUNSUPPORTED, "nms: nc=%d too large", d.nc);
    const int tiles = yre_cdiv(d.A, 32) * d.B;
    nms_filter_kernel<<<yre_cdiv(tiles, warps), warps * 32, per_warp * warps, s>>>(fp);
    YRE_LAUNCH_CHECK("nms_filter");
    return YRE_OK;
}

extern "C" int yre_nms_filter_only(const yre_nms_desc* d, yre_stream_t s) {
    if (!d) YRE_FAIL(YRE_EINVAL, "nms: null descriptor");
    if (int e = check_nms(*d, false)) return e;
    return launch_nms_filter(*d, carve(d->workspace, d->B, d->A), (cudaStream_t)s);
}

int launch_nms(const yre_nms_desc& d, cudaStream_t s) {
    if (int e = check_nms(d, true)) return e;
    NmsWs ws = carve(d.workspace, d.B, d.A);
    if (int e = launch_nms_filter(d, ws, s)) return e;

    SelectParams sp;
    sp.pred = d.pred; sp.B = d.B; sp.A = d.A; sp.nc = d.nc; sp.iou = d.iou_thres; sp.max_det = d.max_det;
    sp.agnostic = d.agnostic; sp.ws = ws; sp.out = d.out; sp.counts = d.counts;
    sp.keep_anchor = reinterpret_cast<long long*>(d.keep_anchor);
    sp.keys_in_smem = SMEM_KEYS;
    sp.scale = d.scale;
    const size_t smem = select_smem_bytes(d.max_det, SMEM_KEYS);
    static YrePerDeviceOnce once;
    if (int e = once.run([]() -> int {
            YRE_CUDA(cudaFuncSetAttribute(nms_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            return YRE_OK;
        })) return e;
    if (smem > 200 * 1024) YRE_FAIL(YRE_EUNSUPPORTED, "nms: shared memory budget exceeded");
    nms_select_kernel<<<d.B, NT_SEL, smem, s>>>(sp);
    YRE_LAUNCH_CHECK("nms_select");
    return YRE_OK;
}
