// fft2d.cuh -- batched fp64 complex 2-D FFT passes for sm_100a.
//
// A 2-D transform is two launches of k_fft_pass: a ROW pass (contiguous lines)
// and a COL pass (lines strided by the row pitch, W adjacent columns per CTA so
// every global access is W*16 contiguous bytes).  Each CTA keeps its whole tile
// (W lines x N points, 16 points per thread) in registers, so both passes are
// in-place safe and every point crosses HBM exactly once per pass.
// Inverse transforms run the forward kernel on conjugated data
// (ifft(x) = conj(fft(conj(x)))/N^2): conj on the first pass's load, conj and the
// 1/N^2 scale on the second pass's store -- numpy's convention
// (niwqg/Kernel.py:565-566).
#pragma once
#include "common.cuh"

// prologue modes: what the first pass multiplies the loaded spectral value by
enum {
    PRO_NONE = 0,
    PRO_REAL_IN,   // input array is double (imaginary part 0)
    PRO_IK,        // * i k          (phix, Kernel.py:610)
    PRO_IL,        // * i l          (phiy)
    PRO_NEG_WV2,   // * -(k^2+l^2)   (lapphi, Kernel.py:685)
    PRO_WV4,       // * (k^2+l^2)^2  (lap2phi, Kernel.py:688)
    PRO_UV,        // * (-il' + i*ik') with Nyquist lines zeroed: packs u + i v of a Hermitian ph (Kernel.py:681)
};
enum { EPI_NONE = 0, EPI_REAL_OUT };   // EPI_REAL_OUT: store the real part into a double array

struct FftArgs {
    const void* in;
    void* out;
    const cd* tw;
    int pro, epi;
    int conj_in, conj_out;
    double scale;
    double dk;
};

template <int N>
__device__ __forceinline__ cd fft_load(const FftArgs& a, size_t mbase, int row, int col) {
    const size_t idx = mbase + (size_t)row * N + col;
    cd x;
    if (a.pro == PRO_REAL_IN) {
        x = make_double2(((const double*)a.in)[idx], 0.0);
    } else {
        x = ((const cd*)a.in)[idx];
        if (a.pro != PRO_NONE) {
            const double k = a.dk * (double)sidx(col, N), l = a.dk * (double)sidx(row, N);
            switch (a.pro) {
                case PRO_IK: x = make_double2(-k * x.y, k * x.x); break;
                case PRO_IL: x = make_double2(-l * x.y, l * x.x); break;
                case PRO_NEG_WV2: { double w = -(k * k + l * l); x = make_double2(w * x.x, w * x.y); } break;
                case PRO_WV4: { double w = k * k + l * l; w = w * w; x = make_double2(w * x.x, w * x.y); } break;
                case PRO_UV: {
                    const double kz = (col == (N >> 1)) ? 0.0 : k, lz = (row == (N >> 1)) ? 0.0 : l;
                    // (-i lz + i*(i kz)) * x = (-kz - i lz) * x
                    x = make_double2(-kz * x.x + lz * x.y, -kz * x.y - lz * x.x);
                } break;
                default: break;
            }
        }
    }
    if (a.conj_in) x.y = -x.y;
    return x;
}

template <int N>
__device__ __forceinline__ void fft_store(const FftArgs& a, size_t mbase, int row, int col, cd x) {
    const size_t idx = mbase + (size_t)row * N + col;
    if (a.conj_out) x.y = -x.y;
    x.x *= a.scale;
    x.y *= a.scale;
    if (a.epi == EPI_REAL_OUT) ((double*)a.out)[idx] = x.x;
    else ((cd*)a.out)[idx] = x;
}

template <int N, int W, bool COL> struct Tile {
    static constexpr int TPF = N / fftc::E;      // threads per transform
    static constexpr int T = W * TPF;            // threads per CTA
    static constexpr int LINE = fftc::phys_len(N);
    static constexpr size_t SMEM = (size_t)W * LINE * sizeof(cd);
    __device__ static __forceinline__ int slot(int w, int o) {
        return COL ? fftc::phys(o) * W + w : w * LINE + fftc::phys(o);
    }
};

template <int N, int W, bool COL, int NS>
__device__ __forceinline__ void fft_stages(cd (&v)[fftc::E], int j, int w, cd* smem, const FftArgs& a,
                                           int line, size_t mbase) {
    using TL = Tile<N, W, COL>;
    constexpr int R = fftc::StageRadix<N, NS>::R;
    constexpr int S = fftc::E / R;
    constexpr bool LAST = (NS * R == N);
    fftc::stage_compute<N, NS>(v, j, a.tw);
    if (LAST) {
#pragma unroll
        for (int u = 0; u < S; ++u)
#pragma unroll
            for (int p = 0; p < R; ++p) {
                const int o = fftc::stage_out_index<N, NS>(j, u, p);
                fft_store<N>(a, mbase, COL ? o : line, COL ? line : o, v[u + p * S]);
            }
    } else {
#pragma unroll
        for (int u = 0; u < S; ++u)
#pragma unroll
            for (int p = 0; p < R; ++p) smem[TL::slot(w, fftc::stage_out_index<N, NS>(j, u, p))] = v[u + p * S];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < fftc::E; ++e) v[e] = smem[TL::slot(w, j + e * TL::TPF)];
        __syncthreads();
        fft_stages<N, W, COL, LAST ? NS : NS * R>(v, j, w, smem, a, line, mbase);
    }
}

template <int N, int W, bool COL>
__global__ void __launch_bounds__(W * N / 16) k_fft_pass(FftArgs a) {
    using TL = Tile<N, W, COL>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* smem = reinterpret_cast<cd*>(smem_raw);
    const int tid = threadIdx.x;
    int w, j;
    if (COL) { w = tid % W; j = tid / W; } else { j = tid % TL::TPF; w = tid / TL::TPF; }
    const int line = blockIdx.x * W + w;
    const size_t mbase = (size_t)blockIdx.y * N * N;
    cd v[fftc::E];
#pragma unroll
    for (int e = 0; e < fftc::E; ++e) {
        const int o = j + e * TL::TPF;
        v[e] = fft_load<N>(a, mbase, COL ? o : line, COL ? line : o);
    }
    fft_stages<N, W, COL, 1>(v, j, w, smem, a, line, mbase);
}

// tile widths: 8192 points per CTA for N >= 1024 (512 threads, 128 regs/thread), 8 lines below
template <int N> struct TileW { static constexpr int W = (N >= 1024) ? (8192 / N) : 8; };

template <int N, bool COL>
static cudaError_t launch_pass_n(const FftArgs& a, int batch, cudaStream_t st) {
    constexpr int W = TileW<N>::W;
    using TL = Tile<N, W, COL>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_fft_pass<N, W, COL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)TL::SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid(N / W, batch);
    k_fft_pass<N, W, COL><<<grid, TL::T, TL::SMEM, st>>>(a);
    return cudaGetLastError();
}

template <bool COL>
static cudaError_t launch_pass(int N, const FftArgs& a, int batch, cudaStream_t st) {
    switch (N) {
        case 32: return launch_pass_n<32, COL>(a, batch, st);
        case 64: return launch_pass_n<64, COL>(a, batch, st);
        case 128: return launch_pass_n<128, COL>(a, batch, st);
        case 256: return launch_pass_n<256, COL>(a, batch, st);
        case 512: return launch_pass_n<512, COL>(a, batch, st);
        case 1024: return launch_pass_n<1024, COL>(a, batch, st);
        case 2048: return launch_pass_n<2048, COL>(a, batch, st);
        case 4096: return launch_pass_n<4096, COL>(a, batch, st);
        case 8192: return launch_pass_n<8192, COL>(a, batch, st);
        default: return cudaErrorInvalidValue;
    }
}
