// Micro-benchmark (GPU box): cycles per iteration of the NMS "resolve" loop (ffs -> row load -> OR) for
//   variant 0: one thread, 64-bit words, 8 loads per keep      variant 1: one warp, 32-bit words, 2 loads per keep
// with the rest of a 1024-thread CTA parked at __syncthreads, or a 32-thread CTA.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__global__ void k(int variant, long long* out, int nkeep) {
    __shared__ unsigned mask32[512 * 16];
    __shared__ unsigned alive_w[16];
    __shared__ int kpos[512];
    for (int i = threadIdx.x; i < 512 * 16; i += blockDim.x) mask32[i] = 0u;      // nothing suppresses anything: 512 keeps
    if (threadIdx.x < 16) alive_w[threadIdx.x] = 0xffffffffu;
    __syncthreads();
    long long t0 = clock64();
    int nk = 0;
    if (variant == 0 && threadIdx.x == 0) {
        const u64* mask = reinterpret_cast<const u64*>(mask32);
        u64 removed[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) removed[w] = 0;
        bool full = false;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (full) continue;
            u64 cand = ((u64)alive_w[2 * w] | ((u64)alive_w[2 * w + 1] << 32)) & ~removed[w];
            while (cand) {
                const int i = w * 64 + __ffsll((long long)cand) - 1;
                kpos[nk++] = i;
                if (nk >= nkeep) { full = true; break; }
                const u64* mrow = mask + (size_t)i * 8;
#pragma unroll
                for (int v = 0; v < 8; ++v) if (v >= w) removed[v] |= mrow[v];
                cand &= cand - 1; cand &= ~removed[w];
            }
        }
    } else if (variant >= 1 && threadIdx.x < 32) {
        const int ln = threadIdx.x, nw = 16;
        unsigned removed_own = 0u;
        bool full = false;
        for (int w = 0; w < nw && !full; ++w) {
            unsigned rem_w = __shfl_sync(0xffffffffu, removed_own, w);
            unsigned cand = alive_w[w] & ~rem_w;
            while (cand) {
                const int i = w * 32 + __ffs((int)cand) - 1;
                if (ln == 0) kpos[nk] = i;
                ++nk;
                if (nk >= nkeep) { full = true; break; }
                const unsigned* mrow = mask32 + i * 16;
                rem_w |= mrow[w];
                if (variant == 1) { if (ln > w && ln < nw) removed_own |= mrow[ln]; }
                cand &= cand - 1; cand &= ~rem_w;
            }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = nk; out[2] = kpos[0]; }
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    for (int threads : {32, 1024})
        for (int variant : {0, 1, 2}) {
            k<<<1, threads>>>(variant, d, 300); cudaDeviceSynchronize();
            k<<<1, threads>>>(variant, d, 300); cudaDeviceSynchronize();
            long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
            printf("threads %4d variant %d: %lld cycles for %lld keeps = %.1f cycles/keep (%s)\n", threads, variant, h[0], h[1], (double)h[0] / h[1], cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
