// fft2d.cuh -- batched fp64 complex 2-D FFT passes for sm_100a.
//
// A 2-D transform is two launches of k_fft_pass: a ROW pass (contiguous lines)
// and a COL pass (lines strided by the row pitch, W adjacent columns per CTA so
// every global access is W*16 contiguous bytes).  Every point crosses HBM
// exactly once per pass (one read, one write).
//
// Tiles are 4096 points per CTA for N >= 512: 256 threads x 16 points in
// registers, <= 128 registers per thread, 68 KB of shared memory, so TWO CTAs are
// resident per SM and one CTA's global loads overlap the other's butterflies.
// A line longer than the tile (column pass of N >= 2048 with W = 4, row pass of
// N = 8192) is split over a thread-block CLUSTER of C CTAs, decimation in time:
//
//     CTA c transforms the sub-sequence x[C*m + c], m = 0..M-1 (M = N/C) locally and parks
//     E_c[k] in its own shared memory; after a cluster barrier CTA c' owns the outputs
//     k in [c' M/C, (c'+1) M/C): it gathers the C partial results of every such k through
//     DISTRIBUTED SHARED MEMORY, multiplies by w_N^{r k}, runs the radix-C butterfly and
//     stores X[k + M q], q = 0..C-1.
//
// In the column pass the decimated rows C*m + c are whole rows (W*16 B contiguous
// each), so loads stay coalesced.  In the row pass (C = 2) the two CTAs of a line read
// the even / odd elements: each uses half of every 32 B sector, the pair runs at the
// same time, so the line crosses HBM once and L2 serves the second half.
//
// Latency: every CTA first issues L2 prefetches (prefetch.global.L2) for the tile that
// the CTA one resident wave ahead will load, so DRAM latency is paid while the tiles in
// between are transformed and a CTA's own loads are mostly L2 hits.
//
// Inverse transforms run the forward kernel on conjugated data
// (ifft(x) = conj(fft(conj(x)))/N^2): conj on the first pass's load, conj and the
// 1/N^2 scale on the second pass's store -- numpy's convention
// (niwqg/Kernel.py:565-566).
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

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
    const cd* tw;     // stage twiddle blocks of the local length M (fftc::tw_offset)
    const cd* twc;    // w_N^t, t = 0..N-1: cluster twiddles (C > 1 only)
    int pro, epi;
    int conj_in, conj_out;
    double scale;
    double dk;
    int pf_groups;    // L2 prefetch distance in line groups (0 = off)
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

// M = local transform length, W = lines per CTA, C = CTAs per line group (cluster size), N = M*C
template <int M, int W, int C, bool COL> struct Tile {
    static constexpr int N = M * C;
    static constexpr int TPF = M / fftc::E;      // threads per local transform
    static constexpr int T = W * TPF;            // threads per CTA
    static constexpr int LINE = fftc::phys_len(M);
    static constexpr size_t SMEM = (size_t)W * LINE * sizeof(cd);
    static constexpr int MINB = (T <= 256) ? 2 : 1;   // resident CTAs per SM the register budget is cut for
    __device__ static __forceinline__ int slot(int w, int o) {
        return COL ? fftc::phys(o) * W + w : w * LINE + fftc::phys(o);
    }
};

// what happens to the result of the last local stage
template <int M, int W, int C, bool COL>
__device__ __forceinline__ void fft_emit(const FftArgs& a, size_t mbase, int line, int w, int c, int k, cd x, cd* smem) {
    using TL = Tile<M, W, C, COL>;
    if constexpr (C == 1) {
        fft_store<TL::N>(a, mbase, COL ? k : line, COL ? line : k, x);
    } else {
        smem[TL::slot(w, k)] = x;   // E_c[k]; the cluster twiddle w_N^{c k} is applied by the gathering CTA
    }
}

template <int M, int W, int C, bool COL, int NS>
__device__ __forceinline__ void fft_stages(cd (&v)[fftc::E], int j, int w, int c, cd* smem, const FftArgs& a,
                                           int line, size_t mbase) {
    using TL = Tile<M, W, C, COL>;
    constexpr int R = fftc::StageRadix<M, NS>::R;
    constexpr int S = fftc::E / R;
    constexpr bool LAST = (NS * R == M);
    fftc::stage_compute<M, NS>(v, j, a.tw);
    if (LAST) {
#pragma unroll
        for (int u = 0; u < S; ++u)
#pragma unroll
            for (int p = 0; p < R; ++p)
                fft_emit<M, W, C, COL>(a, mbase, line, w, c, fftc::stage_out_index<M, NS>(j, u, p), v[u + p * S], smem);
    } else {
#pragma unroll
        for (int u = 0; u < S; ++u)
#pragma unroll
            for (int p = 0; p < R; ++p) smem[TL::slot(w, fftc::stage_out_index<M, NS>(j, u, p))] = v[u + p * S];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < fftc::E; ++e) v[e] = smem[TL::slot(w, j + e * TL::TPF)];
        __syncthreads();
        fft_stages<M, W, C, COL, LAST ? NS : NS * R>(v, j, w, c, smem, a, line, mbase);
    }
}

// split cluster barrier (arrive / wait), with and without memory ordering
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed_after(double dep) {
    asm volatile("{\n\t.reg .f64 t;\n\tmov.f64 t, %0;\n\tbarrier.cluster.arrive.relaxed.aligned;\n\t}" ::"d"(dep) : "memory");
}
__device__ __forceinline__ void cluster_wait_relaxed() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }

// L2 prefetch of the tile that the CTA `pf_groups` line groups ahead will load: the DRAM latency of that tile
// is paid while the tiles in between are transformed, so a CTA's own loads are (mostly) L2 hits.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int M, int W, int C, bool COL>
__device__ __forceinline__ void fft_prefetch(const FftArgs& a, int group, int c, int tid) {
    using TL = Tile<M, W, C, COL>;
    constexpr int N = TL::N;
    const int g = group + a.pf_groups;
    if (COL || a.pf_groups <= 0 || g >= N / W) return;   // column pass: the per-row requests cost more L1 wavefronts than they save
    const size_t esz = (a.pro == PRO_REAL_IN) ? sizeof(double) : sizeof(cd);
    const char* base = (const char*)a.in + (size_t)blockIdx.y * N * N * esz;
    if (COL) {
        // rows C m + c, columns [g W, g W + W): one line-sized request per row
        for (int m = tid; m < M; m += TL::T) prefetch_l2(base + ((size_t)(C * m + c) * N + (size_t)g * W) * esz);
    } else {
        // ROW: lines [g W, g W + W) are contiguous; a cluster CTA reads every C-th element of the whole line,
        // so rank c prefetches the c-th 1/C of it
        const char* p0 = base + ((size_t)g * W * N + (size_t)c * M) * esz;
        const size_t bytes = (size_t)W * M * esz;
        for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)TL::T * 128) prefetch_l2(p0 + off);
    }
}

template <int M, int W, int C, bool COL>
__global__ void __launch_bounds__(W * M / 16, Tile<M, W, C, COL>::MINB) k_fft_pass(FftArgs a) {
    using TL = Tile<M, W, C, COL>;
    constexpr int N = TL::N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* smem = reinterpret_cast<cd*>(smem_raw);
    const int tid = threadIdx.x;
    int w, j;
    if (COL) { w = tid % W; j = tid / W; } else { j = tid % TL::TPF; w = tid / TL::TPF; }
    const int c = (C > 1) ? (int)(blockIdx.x % C) : 0;     // rank in the (C,1,1) cluster
    const int group = blockIdx.x / C;
    const int line = group * W + w;
    const size_t mbase = (size_t)blockIdx.y * N * N;
    cd v[fftc::E];
    fft_prefetch<M, W, C, COL>(a, group, c, tid);
#pragma unroll
    for (int e = 0; e < fftc::E; ++e) {
        const int n = C * (j + e * TL::TPF) + c;           // decimated sub-sequence of CTA c
        v[e] = fft_load<N>(a, mbase, COL ? n : line, COL ? line : n);
    }
    fft_stages<M, W, C, COL, 1>(v, j, w, c, smem, a, line, mbase);
    if constexpr (C > 1) {
        // radix-C butterfly across the cluster: this CTA owns k in [c M/C, (c+1) M/C) of every line of the group
        cg::cluster_group cluster = cg::this_cluster();
        cluster_arrive_release();                          // my E_c[k] are parked
        constexpr int PPT = fftc::E / C;                   // butterflies per thread
        cd wk[PPT];
#pragma unroll
        for (int i = 0; i < PPT; ++i) {                    // cluster twiddles while the peers catch up
            const int g = tid + i * TL::T;
            wk[i] = a.twc[c * (M / C) + (COL ? g / W : g)];
        }
        cluster_wait_acquire();
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int g = tid + i * TL::T;
            const int wq = COL ? g % W : 0, k = c * (M / C) + (COL ? g / W : g);
#pragma unroll
            for (int r = 0; r < C; ++r) {
                const cd* src = cluster.map_shared_rank(smem, r);
                v[i * C + r] = src[TL::slot(wq, k)];
            }
        }
        double dep = 0.0;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            fftc::apply_twiddles<C, 1>(v + i * C, wk[i]);  // E_r[k] *= w_N^{r k}
            fftc::dft<C, 1>(v + i * C);
            dep += v[i * C].x;
        }
        // every value gathered from the peers has been consumed (dep depends on all of them), so the peers may
        // retire; no memory ordering is needed, hence no fence that would wait for the global stores below
        cluster_arrive_relaxed_after(dep);
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int g = tid + i * TL::T;
            const int wq = COL ? g % W : 0, k = c * (M / C) + (COL ? g / W : g);
            const int ln = group * W + wq;
#pragma unroll
            for (int p = 0; p < C; ++p) {
                const int n = k + M * fftc::outidx<C>(p);
                fft_store<N>(a, mbase, COL ? n : ln, COL ? ln : n, v[i * C + p]);
            }
        }
        cluster_wait_relaxed();     // nobody may exit while a peer still reads its shared memory
    }
}

// ---- pass geometry: (M, W, C) per grid size
template <int N, bool COL> struct PassCfg {
    // column pass: W = 4 adjacent columns (64 B per row access), 1024-point local transforms for N >= 1024
    // row pass: whole line per CTA up to 4096 points, two CTAs per line at 8192
    static constexpr int M = COL ? (N >= 1024 ? 1024 : N) : (N > 4096 ? 4096 : N);
    static constexpr int C = N / M;
    static constexpr int W = (M >= 512) ? 4096 / M : 8;
};

template <int N, bool COL>
static cudaError_t launch_pass_n(const FftArgs& a, int batch, cudaStream_t st) {
    constexpr int M = PassCfg<N, COL>::M, W = PassCfg<N, COL>::W, C = PassCfg<N, COL>::C;
    using TL = Tile<M, W, C, COL>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_fft_pass<M, W, C, COL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)TL::SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((N / W) * C, batch, 1);
    cfg.blockDim = dim3(TL::T, 1, 1);
    cfg.dynamicSmemBytes = TL::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = (C > 1) ? 1 : 0;
    FftArgs b = a;
    b.pf_groups = (a.pf_groups > 0 && (N / W) * C > 2 * a.pf_groups) ? (a.pf_groups + C - 1) / C : 0;   // CTAs -> groups
    return cudaLaunchKernelEx(&cfg, k_fft_pass<M, W, C, COL>, b);
}

// local transform length of a pass (the stage twiddle table to bind)
static inline int pass_local_len(int N, bool col) { return col ? (N >= 1024 ? 1024 : N) : (N > 4096 ? 4096 : N); }

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
