// common.cuh -- shared device helpers for the niwqg_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include "fft_core.cuh"

#define NIWQG_PW_THREADS 256     // pointwise / reduction CTA size
#define NIWQG_PW_BLOCKS (148 * 4)  // persistent grid-stride grid: 4 CTAs per SM on 148 SMs

// signed wavenumber index of the kernel-family grid: [0..N/2-1, -N/2..-1]  (niwqg/Kernel.py:242-244)
__host__ __device__ __forceinline__ int sidx(int i, int N) { return i < (N >> 1) ? i : i - N; }

// Spectral-array geometry.  Natural layout (one GPU): [N rows ky][N columns kx].  Slab layout (P ranks): every
// rank holds all N rows of ncl = N/P columns, chosen so that column kx and its conjugate partner N-kx live on the
// same rank (the (K,-K) pair kernels stay local): local columns [0,h) are kx = rank*h + lc, local columns [h,2h)
// are their mirrors N - (rank*h + lc - h); rank 0, whose first column kx=0 is its own mirror, holds the Nyquist
// column N/2 (also its own mirror) in slot h instead.
struct Grid {
    int N;          // global grid edge
    double dk;      // 2 pi / L
    int ncl;        // local spectral columns (N when not decomposed)
    int h;          // ncl / 2
    int rank;       // slab rank
    int sym;        // 1 = slab layout, 0 = natural
};
__host__ __device__ __forceinline__ int grid_kx(const Grid& g, int lc) {
    if (!g.sym) return lc;
    if (lc < g.h) return g.rank * g.h + lc;
    if (g.rank == 0 && lc == g.h) return g.N >> 1;
    return g.N - (g.rank * g.h + lc - g.h);
}
__host__ __device__ __forceinline__ int grid_partner(const Grid& g, int lc) {
    if (!g.sym) return (g.N - lc) & (g.N - 1);
    if (g.rank == 0 && (lc == 0 || lc == g.h)) return lc;
    return lc < g.h ? lc + g.h : lc - g.h;
}
// owner rank and local column of global column kx in the slab layout over P ranks (h = N / (2 P))
__host__ __device__ __forceinline__ void grid_owner(int N, int h, int kx, int& rank, int& lc) {
    const int H = N >> 1;
    if (kx < H) { rank = kx / h; lc = kx % h; }
    else if (kx == H) { rank = 0; lc = h; }
    else { const int m = N - kx; rank = m / h; lc = h + m % h; }
}

// Deterministic block reduction of K partial sums; block result lands in partials[blockIdx.x*K + k].
template <int K>
__device__ __forceinline__ void block_reduce_store(double (&s)[K], double* __restrict__ partials) {
    __shared__ double sh[NIWQG_PW_THREADS / 32][K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double x = s[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
        if (lane == 0) sh[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double x = 0.0;
#pragma unroll
        for (int w = 0; w < NIWQG_PW_THREADS / 32; ++w) x += sh[w][threadIdx.x];
        partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * K + threadIdx.x] = x;
    }
}

template <int K>
__device__ __forceinline__ void block_reduce_max_store(double (&s)[K], double* __restrict__ partials) {
    __shared__ double shm[NIWQG_PW_THREADS / 32][K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double x = s[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x = fmax(x, __shfl_down_sync(0xffffffffu, x, off));
        if (lane == 0) shm[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double x = shm[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < NIWQG_PW_THREADS / 32; ++w) x = fmax(x, shm[w][threadIdx.x]);
        partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * K + threadIdx.x] = x;
    }
}
