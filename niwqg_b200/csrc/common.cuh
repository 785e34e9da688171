// common.cuh -- shared device helpers for the niwqg_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include "fft_core.cuh"

#define NIWQG_PW_THREADS 256     // pointwise / reduction CTA size
#define NIWQG_PW_BLOCKS (148 * 4)  // persistent grid-stride grid: 4 CTAs per SM on 148 SMs

// signed wavenumber index of the kernel-family grid: [0..N/2-1, -N/2..-1]  (niwqg/Kernel.py:242-244)
__host__ __device__ __forceinline__ int sidx(int i, int N) { return i < (N >> 1) ? i : i - N; }

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
