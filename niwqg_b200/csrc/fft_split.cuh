// fft_split.cuh -- 2-D transforms of the largest grids (N >= 2048, natural single-GPU layout) as three streaming
// launches, each at the rate of a pass whose lines fit ONE tile:
//
//   x direction:  N = 2 * Nh.  The PHYSICAL side of every row is stored de-interleaved ([even x | odd x], k_deint), so the
//                 two Nh-point transforms of a row are plain contiguous lines: one-tile row kernel (k_fft_pass, C = 1)
//                 on 2N lines.  The remaining radix-2 butterfly (pairs kx, kx + Nh) rides on the streaming kernel below.
//   y direction:  N = 16 * M.  M-point transforms of W = 4096/M adjacent columns fit one tile (k_fft_colsub2, full
//                 >= 128 B rows, no cluster, no DSMEM); the radix-16 butterfly across the 16 row blocks is a pure
//                 streaming kernel (k_split_p): one thread = 16 rows x 1 column in registers, lanes l and l+16 of a
//                 warp hold the columns kx and kx + Nh and swap them with shuffles for the x butterfly.
//
//   forward  (physical -> spectral):  rows (Nh)  ->  k_fft_colsub2<DIT> (rows 16 j + r -> block r)  ->  k_split_p<DIT>
//   inverse  (spectral -> physical):  k_split_p<DIF> (prologue, conj)  ->  k_fft_colsub2<DIF>  ->  rows (Nh), conj + 1/N^2
//
// The streaming kernel always sits on the SPECTRAL side, so several transforms of one spectrum (phi, phix, phiy) share
// its loads (NOUT outputs), and the cluster kernels with their distributed-shared-memory exchange (8192^2: row pass
// 0.52 ms, column pass 0.69-0.77 ms per GiB) are replaced by passes that run at 0.32-0.37 ms per GiB.
// Reference operation: numpy.fft.fft2 / ifft2 as bound by niwqg/Kernel.py:565-566.
#pragma once
#include "fft2d.cuh"

struct SplitPArgs {
    const cd* in;
    cd* out[3];
    int pro[3];        // spectral prologue of every output (inverse side only)
    int nout;
    int conj_in;
    double scale;      // forward side: output multiplier (1)
    double dk;
    const cd* twc;     // w_N^t, t = 0..N-1
};

template <int N, int PRO>
__device__ __forceinline__ cd split_prologue(double dk, int row, int col, cd x) {
    FftArgs a{};
    a.dk = dk;
    return fft_prologue_one<N, PRO>(a, row, col, x);
}

// Inverse side (decimation in frequency in both directions): thread (m, n) loads s[m + M r][n], r = 0..15
//   x:  a = s[n] + s[n + Nh]   (-> even x),   b = (s[n] - s[n + Nh]) w_N^n   (-> odd x)
//   y:  Y_q[m] = (sum_r t_r w_16^{r q}) w_N^{m q}, stored as row q M + m (block q = input of the M-point transforms)
// Forward side (decimation in time): thread (k, n) loads E_r[k][n] (row r M + k), r = 0..15
//   x:  S = E[n] + w_N^n E[n + Nh] (-> kx = n),   D = E[n] - w_N^n E[n + Nh] (-> kx = n + Nh)
//   y:  X[k + M q] = sum_r (w_N^{k r} t_r) w_16^{r q}, stored as row k + M q (natural)
// Both read and write the same 16 rows of the same two columns: in place is safe.
template <int N, bool DIT>
__global__ void __launch_bounds__(256, 2) k_split_p(SplitPArgs a) {
    constexpr int R = 16, M = N / R, Nh = N / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int side = lane >> 4;
    const int n = blockIdx.x * 128 + warp * 16 + (lane & 15);        // column of the pair's first half, n < Nh
    const int col = n + side * Nh;
    const int m = blockIdx.y;                                         // row inside a block of M rows
    const cd wx = a.twc[n];                                           // w_N^n
    const cd wy = a.twc[m];                                           // w_N^m
    for (int o = 0; o < a.nout; ++o) {
        cd v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = a.in[(size_t)(DIT ? r * M + m : m + M * r) * N + col];
        if (!DIT) {
#define NIWQG_PRO_CASE(P)                                                                     \
    case P:                                                                                   \
        _Pragma("unroll") for (int r = 0; r < R; ++r) v[r] = split_prologue<N, P>(a.dk, m + M * r, col, v[r]); \
        break;
            switch (a.pro[o]) {
                NIWQG_PRO_CASE(PRO_IK)
                NIWQG_PRO_CASE(PRO_IL)
                NIWQG_PRO_CASE(PRO_NEG_WV2)
                NIWQG_PRO_CASE(PRO_WV4)
                NIWQG_PRO_CASE(PRO_UV)
                default: break;
            }
#undef NIWQG_PRO_CASE
            if (a.conj_in) {
#pragma unroll
                for (int r = 0; r < R; ++r) v[r].y = -v[r].y;
            }
        }
        // the radix-2 butterfly along x between lanes l and l + 16
        if (DIT) {
            if (side) {
#pragma unroll
                for (int r = 0; r < R; ++r) v[r] = cmul(v[r], wx);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const cd o2 = make_double2(__shfl_xor_sync(0xffffffffu, v[r].x, 16), __shfl_xor_sync(0xffffffffu, v[r].y, 16));
                v[r] = side ? csub(o2, v[r]) : cadd(v[r], o2);     // E - w O on the second half, E + w O on the first
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const cd o2 = make_double2(__shfl_xor_sync(0xffffffffu, v[r].x, 16), __shfl_xor_sync(0xffffffffu, v[r].y, 16));
                v[r] = side ? cmul(csub(o2, v[r]), wx) : cadd(v[r], o2);
            }
        }
        cd u[R];
        if (DIT) {
            fftc::apply_twiddles<R, 1>(v, wy);                       // E_r[k] *= w_N^{k r}
            fftc::dft<R, 1>(v);
#pragma unroll
            for (int p = 0; p < R; ++p) u[fftc::outidx<R>(p)] = v[p];
        } else {
            fftc::dft<R, 1>(v);
#pragma unroll
            for (int p = 0; p < R; ++p) u[fftc::outidx<R>(p)] = v[p];
            fftc::apply_twiddles<R, 1>(u, wy);                       // Y_q[m] *= w_N^{m q}
        }
        cd* out = a.out[o];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            cd x = u[q];
            if (DIT) { x.x *= a.scale; x.y *= a.scale; }
            out[(size_t)(DIT ? m + M * q : q * M + m) * N + col] = x;
        }
    }
}

// One-tile M-point column transforms of W adjacent columns, N = R * M rows.
//   DIF (inverse side): reads block q (rows q M + j), stores X[R k + q]   (= k_fft_colsub)
//   DIT (forward side): reads the decimated rows R j + q, stores block q (row q M + k)
template <int M, int W, int R, bool DIT>
__global__ void __launch_bounds__(W * M / 16, 2) k_fft_colsub2(FftArgs a) {
    using TL = Tile<M, W, R, true>;
    constexpr int N = TL::N;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd* smem = reinterpret_cast<cd*>(smem_raw);
    cd* smtw = smem + (size_t)W * TL::LINE;
    const int tid = threadIdx.x, w = tid % W, j = tid / W;
    const int q = (int)(blockIdx.x % R), group = (int)(blockIdx.x / R);
    const int line = group * W + w;
    const cd* in = (const cd*)a.in;
    cd v[fftc::E];
#pragma unroll
    for (int e = 0; e < fftc::E; ++e) {
        const int jj = j + e * TL::TPF;
        v[e] = in[(size_t)(DIT ? R * jj + q : q * M + jj) * N + line];
    }
    for (int t = tid; t < TL::TWLEN; t += TL::T) smtw[t] = a.tw[t];
    fft_stages<M, W, R, true, true, true, 1>(v, j, w, q, smem, smtw, a, line, 0);   // a.deint_out = DIT selects the store map
}

template <int N>
static cudaError_t split_set_attrs() {
    constexpr int M = N / 16, W = 4096 / M;
    using TL = Tile<M, W, 16, true>;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_fft_colsub2<M, W, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL::SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_fft_colsub2<M, W, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL::SMEM);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    return cudaSuccess;
}

// the M-point column transforms; `a` carries in / out / tw (M-point stage twiddles) / scale / epi
template <int N, bool DIT>
static cudaError_t launch_split_colsub(const FftArgs& a, cudaStream_t st) {
    constexpr int M = N / 16, W = 4096 / M;
    using TL = Tile<M, W, 16, true>;
    cudaError_t e = split_set_attrs<N>();
    if (e != cudaSuccess) return e;
    FftArgs b = a;
    b.deint_out = DIT ? 1 : 0;
    k_fft_colsub2<M, W, 16, DIT><<<dim3((N / W) * 16, 1), TL::T, TL::SMEM, st>>>(b);
    return cudaGetLastError();
}

// forward row pass of the split path (2N lines of N/2) that forms its input with a loader (LD_*): in / in2 / in3 -> out
template <int N, int LD>
static cudaError_t launch_split_rows_loader(const FftArgs& a, cudaStream_t st) {
    constexpr int Nh = N / 2, W = 4096 / Nh;
    using TL = Tile<Nh, W, 1, false>;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_fft_pass<Nh, W, 1, false, true, LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL::SMEM);
        if (e != cudaSuccess) return e;
        if (dev < 64) attr_set[dev] = true;
    }
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    FftArgs b = a;
    b.tma_in = 0; b.pf_groups = 0;
    k_fft_pass<Nh, W, 1, false, true, LD><<<dim3(a.nlines / W, 1, 1), TL::T, TL::SMEM, st>>>(b, tmap);
    return cudaGetLastError();
}

template <int N, bool DIT>
static cudaError_t launch_split_p(const SplitPArgs& a, cudaStream_t st) {
    k_split_p<N, DIT><<<dim3(N / 2 / 128, N / 16, 1), 256, 0, st>>>(a);
    return cudaGetLastError();
}
