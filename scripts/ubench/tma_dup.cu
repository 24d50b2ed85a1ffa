// Probe: does a tiled tensor map accept global stride 0 (a duplicated dimension)?  If it does, a 2x nearest-neighbour
// upsample is a TMA box {C, 2, w, 2, h} over the LOW-RES tensor with the two "2" dimensions at stride 0.
//   nvcc -arch=sm_100a -o tma_dup tma_dup.cu -lcuda && ./tma_dup
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void k(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* out, int n, int c0, int x0, int r0) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    unsigned b = (unsigned)__cvta_generic_to_shared(&bar), d = (unsigned)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n * 2));
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                     ::"r"(d), "l"(reinterpret_cast<unsigned long long>(&tm)), "r"(b), "r"(c0), "r"(0), "r"(x0), "r"(0), "r"(r0) : "memory");
    }
    unsigned ok = 0;
    int spins = 0;
    while (!ok && spins < (1 << 22)) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
        ++spins;
    }
    __syncthreads();
    const __nv_bfloat16* s = reinterpret_cast<const __nv_bfloat16*>(sm);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = ok ? s[i] : __float2bfloat16(-1.f);
}

int main() {
    const int B = 2, H = 6, W = 8, C = 64;            // low-res tensor [B][H][W][C]
    std::vector<__nv_bfloat16> h((size_t)B * H * W * C);
    for (size_t i = 0; i < h.size(); ++i) h[i] = __float2bfloat16((float)(i % 251));
    __nv_bfloat16 *dx, *dout;
    cudaMalloc(&dx, h.size() * 2);
    cudaMemcpy(dx, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    const int bw = 4, bh = 4;                          // low-res box -> 8 x 8 high-res pixels
    const int n = C * 2 * bw * 2 * bh;
    cudaMalloc(&dout, n * 2);
    CUtensorMap tm;
    cuuint64_t gdim[5] = {(cuuint64_t)C, 2, (cuuint64_t)W, 2, (cuuint64_t)H * B};
    cuuint32_t box[5] = {(cuuint32_t)C, 2, (cuuint32_t)bw, 2, (cuuint32_t)bh};
    cuuint32_t est[5] = {1, 1, 1, 1, 1};
    int rc_all = 1;
    for (int variant = 0; variant < 2; ++variant) {
        cuuint64_t gstr[4] = {0, (cuuint64_t)C * 2, 0, (cuuint64_t)W * C * 2};
        CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dx, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            variant ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("variant %d (swizzle %s): encode -> %d\n", variant, variant ? "128B" : "none", (int)r);
        if (r != CUDA_SUCCESS) continue;
        const int x0 = 2, r0 = 1 * H + 1;              // image 1, low-res row 1, col 2
        k<<<1, 128, n * 2 + 1024>>>(tm, dout, n, 0, x0, r0);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  kernel -> %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 2;
        std::vector<__nv_bfloat16> o(n);
        cudaMemcpy(o.data(), dout, n * 2, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < 2 * bh; ++y)
            for (int x = 0; x < 2 * bw; ++x)
                for (int c = 0; c < C; ++c) {
                    const int pix = y * 2 * bw + x;
                    int idx = pix * C + c;
                    if (variant) {                     // 128B swizzle: 16-byte chunk index XOR (row & 7)
                        const int chunk = (c / 8) ^ (pix & 7);
                        idx = pix * C + chunk * 8 + (c % 8);
                    }
                    const size_t src = (((size_t)1 * H + (1 + y / 2)) * W + (x0 + x / 2)) * C + c;
                    if (__bfloat162float(o[idx]) != __bfloat162float(h[src])) ++bad;
                }
        printf("  mismatches: %d of %d\n", bad, n);
        if (bad == 0) rc_all = 0;
    }
    return rc_all;
}
