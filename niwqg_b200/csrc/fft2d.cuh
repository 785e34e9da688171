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
#include <cuda.h>      // CUtensorMap (TMA descriptors); the encoder is fetched through cudaGetDriverEntryPoint
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
    PRO_IK_CONJ,   // * conj(i k) = -i k (slab layout: the row pass is the second pass of an inverse transform)
    PRO_IL_CONJ,   // * conj(i l) = -i l: the i l factor applied in the SECOND pass of an inverse transform, whose
                   //   intermediate data lives in the conjugated domain (ifft = conj fft conj)
};
enum { EPI_NONE = 0, EPI_REAL_OUT };   // EPI_REAL_OUT: store the real part into a double array

struct FftArgs {
    const void* in;
    void* out;
    const cd* tw;     // stage twiddle blocks of the local length M (fftc::tw_offset)
    const cd* twc;    // w_N^t, t = 0..N-1: cluster twiddles (C > 1 only)
    int pro, epi;
    int conj_in, conj_out;
    double scale, scale_im;   // output multipliers of the real / imaginary part (scale_im = -scale: conjugate)
    double dk;
    int pf_groups;    // L2 prefetch distance in line groups (0 = off)
    // geometry: natural [lines][N] arrays unless stated otherwise
    Grid g;           // spectral column map (prologue wavenumbers, exchange layout)
    int nlines;       // lines transformed by this pass: rows (row pass) or local columns (column pass)
    int pitch;        // column pass: elements between consecutive rows (N, or ncl in the slab layout)
    int xmap_in, xmap_out;   // row pass of a slab transform: load / store through the all-to-all exchange layout
    int xchunk;       // exchange layout: elements per peer chunk (nyl * ncl)
    int deint_in, deint_out;   // row pass split over a cluster: the PHYSICAL side of the line is stored de-interleaved
                      // (position c*M + m holds x = C*m + c), so both cluster kernels touch contiguous memory
    int push;         // slab: the pass's stores go straight into the owners' receive buffers over NVLink (peer[])
    int nyl_shift;    // log2(rows per rank)
    int panel;        // inverse slab transform, pushed exchange: the receive buffer is laid out in panels of 4 columns,
                      // [source rank][panel][row residue mod C][row / C][4 columns] with C = panel = cluster size of the
                      // pushing column pass, so that what one CTA pushes to one peer is contiguous (8 KB runs instead of
                      // 64 B segments: 409 -> ~620 GB/s over NVLink, profiles/r02_nvlink_push_patterns_8gpu.txt); 0 = off
    cd* peer[8];      // receive buffer of every rank (peer[rank] = own), CUDA-IPC mapped
    size_t mstride;   // elements between ensemble members
    int tma_in;       // column pass, natural layout, one tile per line group: the tile is fetched by TMA (cp.async.bulk.tensor)
                      // straight into shared memory instead of 16 B per-thread loads of half-used 128 B lines
    int one_cta_per_sm;   // request enough shared memory that only ONE CTA of this launch fits on an SM: leaves the other
                      // half of every SM to the pass that runs concurrently on the other lane (slab overlap)
    int variant;      // tuning: bit 0 / bit 1 = column / row cluster passes use the decimation-in-time (pull) kernel;
                      // bit 2 = the push kernel uses plain remote stores + a cluster barrier instead of st.async
    // row pass with a physical-product loader (k_fft_pass<..., LD>): further operands in the layout of `in`
    const cd *in2, *in3;
    double ld_scale;  // LD_WAVEPV: jscale of the Jacobian partner
};
// loaders: what a forward row pass forms from its operands while loading them (instead of a pointwise kernel + a re-read)
enum {
    LD_NONE = 0,
    LD_WAVEPV,    // |phi|^2 + i jscale i J(phi*, phi) from in = phi, in2 = phix, in3 = phiy (k_phys_wavepv)
    LD_UQVQ,      // u q + i v q from in = u + i v, in2 = q + i qw (the P1 product of k_phys_rhs, Kernel.py:479-483)
};

// element n of line `line`: offset inside one member's array
//   column pass: n-th row of local column `line`;  row pass: natural [line][n], or - on the exchange side of a slab
//   transform - [owner(n)][line][lc(n)] (the layout ncclAlltoAll moves between the row slabs and the column slabs)
// NAT: natural single-GPU layout - pitch N, no exchange maps, no pushes: all strides are compile-time constants
// offset inside one source rank's chunk of the panel layout: local row yl (of nyl = 2^nyl_shift), local column lc
__device__ __forceinline__ size_t panel_offset(const FftArgs& a, int yl, int lc) {
    int cs = 0;
    while ((1 << cs) < a.panel) ++cs;
    const int rp = ((yl & (a.panel - 1)) << (a.nyl_shift - cs)) | (yl >> cs);      // residue-major row order
    return ((((size_t)(lc >> 2)) << a.nyl_shift) + rp) * 4 + (lc & 3);
}

template <int N, bool COL, bool NAT>
__device__ __forceinline__ size_t fft_index(const FftArgs& a, int xmap, int line, int n) {
    if (NAT) return COL ? (size_t)n * N + line : (size_t)line * N + n;
    if (COL) return (size_t)n * a.pitch + line;
    if (!xmap) return (size_t)line * N + n;
    int r, lc;
    grid_owner(N, a.g.h, n, r, lc);
    if (a.panel) return (size_t)r * a.xchunk + panel_offset(a, line, lc);
    return (size_t)r * a.xchunk + (size_t)line * a.g.ncl + lc;
}
// (ky, kx) of element n of line `line` on the spectral side (prologue wavenumbers)
template <bool COL>
__device__ __forceinline__ void fft_kykx(const FftArgs& a, int line, int n, int& ky, int& kx) {
    if (COL) { ky = n; kx = a.g.sym ? grid_kx(a.g, line) : line; } else { ky = line; kx = n; }
}

// spectral multiplier of the prologue modes at (row, col) applied to x
template <int N, int PRO>
__device__ __forceinline__ cd fft_prologue_one(const FftArgs& a, int row, int col, cd x) {
    const double k = a.dk * (double)sidx(col, N), l = a.dk * (double)sidx(row, N);
    if (PRO == PRO_IK) return make_double2(-k * x.y, k * x.x);
    if (PRO == PRO_IL) return make_double2(-l * x.y, l * x.x);
    if (PRO == PRO_IL_CONJ) return make_double2(l * x.y, -l * x.x);
    if (PRO == PRO_IK_CONJ) return make_double2(k * x.y, -k * x.x);
    if (PRO == PRO_NEG_WV2) { const double w = -(k * k + l * l); return make_double2(w * x.x, w * x.y); }
    if (PRO == PRO_WV4) { double w = k * k + l * l; w = w * w; return make_double2(w * x.x, w * x.y); }
    if (PRO == PRO_UV) {
        const double kz = (col == (N >> 1)) ? 0.0 : k, lz = (row == (N >> 1)) ? 0.0 : l;
        // (-i lz + i*(i kz)) * x = (-kz - i lz) * x
        return make_double2(-kz * x.x + lz * x.y, -kz * x.y - lz * x.x);
    }
    return x;
}

template <int N, bool COL, bool NAT>
__device__ __forceinline__ void fft_store(const FftArgs& a, size_t mbase, int line, int n, cd x) {
    x.x *= a.scale;
    x.y *= a.scale_im;            // = -scale when the output is conjugated
    if (!NAT && a.push) {
        // the all-to-all of a slab transform fused into this pass: the element lands in its owner's receive buffer,
        // chunk [my rank][row within the owner's slab][local column] - exactly what ncclAlltoAll would deliver
        int r;
        size_t off;
        if (COL) {               // inverse transform, column pass: row n belongs to rank n / nyl
            r = n >> a.nyl_shift;
            const int yl = n - (r << a.nyl_shift);
            off = a.panel ? panel_offset(a, yl, line) : (size_t)yl * a.g.ncl + line;
        } else {                 // forward transform, row pass: column n belongs to owner(n)
            int lc;
            grid_owner(N, a.g.h, n, r, lc);
            off = (size_t)line * a.g.ncl + lc;
        }
        a.peer[r][(size_t)a.g.rank * a.xchunk + off] = x;
        return;
    }
    const size_t idx = mbase + fft_index<N, COL, NAT>(a, a.xmap_out, line, n);
    if (a.epi == EPI_REAL_OUT) ((double*)a.out)[idx] = x.x;
    else ((cd*)a.out)[idx] = x;
}

// M = local transform length, W = lines per CTA, C = CTAs per line group (cluster size), N = M*C
template <int M, int W, int C, bool COL> struct Tile {
    static constexpr int N = M * C;
    static constexpr int TPF = M / fftc::E;      // threads per local transform
    static constexpr int T = W * TPF;            // threads per CTA
    static constexpr int LINE = fftc::phys_len(M);
    static constexpr int TWLEN = fftc::tw_table_len(M) + 1;           // stage twiddles, copied to shared memory
    static constexpr size_t SMEM = ((size_t)W * LINE + TWLEN + 1) * sizeof(cd);   // + one slot for an mbarrier
    static constexpr int MINB = (T <= 256) ? 2 : 1;   // resident CTAs per SM the register budget is cut for
    __device__ static __forceinline__ int slot(int w, int o) {
        return COL ? fftc::phys(o) * W + w : w * LINE + fftc::phys(o);
    }
};

// what happens to the result of the last local stage
template <int M, int W, int C, bool COL, bool NAT, bool DIF>
__device__ __forceinline__ void fft_emit(const FftArgs& a, size_t mbase, int line, int w, int c, int k, cd x, cd* smem) {
    using TL = Tile<M, W, C, COL>;
    if constexpr (C == 1) {
        fft_store<TL::N, COL, NAT>(a, mbase, line, k, x);
    } else if constexpr (DIF) {
        const int n = a.deint_out ? c * M + k : C * k + c;    // decimation in frequency: CTA c produced the outputs
                                                               // congruent to c mod C (stored as one contiguous block
                                                               // when the physical side is de-interleaved)
        fft_store<TL::N, COL, NAT>(a, mbase, line, n, x);
    } else {
        smem[TL::slot(w, k)] = x;   // E_c[k]; the cluster twiddle w_N^{c k} is applied by the gathering CTA
    }
}

template <int M, int W, int C, bool COL, bool NAT, bool DIF, int NS>
__device__ __forceinline__ void fft_stages(cd (&v)[fftc::E], int j, int w, int c, cd* smem, const cd* tw,
                                           const FftArgs& a, int line, size_t mbase) {
    using TL = Tile<M, W, C, COL>;
    constexpr int R = fftc::StageRadix<M, NS>::R;
    constexpr int S = fftc::E / R;
    constexpr bool LAST = (NS * R == M);
    fftc::stage_compute<M, NS>(v, j, tw);
    if (LAST) {
#pragma unroll
        for (int u = 0; u < S; ++u)
#pragma unroll
            for (int p = 0; p < R; ++p)
                fft_emit<M, W, C, COL, NAT, DIF>(a, mbase, line, w, c, fftc::stage_out_index<M, NS>(j, u, p), v[u + p * S], smem);
    } else {
#pragma unroll
        for (int u = 0; u < S; ++u)
#pragma unroll
            for (int p = 0; p < R; ++p) smem[TL::slot(w, fftc::stage_out_index<M, NS>(j, u, p))] = v[u + p * S];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < fftc::E; ++e) v[e] = smem[TL::slot(w, j + e * TL::TPF)];
        __syncthreads();
        fft_stages<M, W, C, COL, NAT, DIF, LAST ? NS : NS * R>(v, j, w, c, smem, tw, a, line, mbase);
    }
}

// split cluster barrier (arrive / wait), with and without memory ordering
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed_after(double dep) {
    asm volatile("{\n\t.reg .f64 t;\n\tmov.f64 t, %0;\n\tbarrier.cluster.arrive.relaxed.aligned;\n\t}" ::"d"(dep) : "memory");
}
__device__ __forceinline__ void cluster_wait_relaxed() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }

// ---- asynchronous distributed-shared-memory stores (st.async): the store lands in a peer CTA's shared memory and
// credits its byte count to the peer's mbarrier, so the producer needs no fence and the consumer waits only for its
// own incoming bytes
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cluster_map_u32(unsigned local_addr, int rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_async_cd(unsigned remote_addr, cd x, unsigned remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
                 ::"r"(remote_addr), "d"(x.x), "d"(x.y), "r"(remote_bar) : "memory");
}

// ---- TMA: 2-D tiled bulk tensor load global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_2d(unsigned dst_smem, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst_smem), "l"((unsigned long long)map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// linear bulk copy global -> shared (cp.async.bulk, SASS UBLKCP), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(unsigned dst_smem, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}


// L2 prefetch of the tile that the CTA `pf_groups` line groups ahead will load: the DRAM latency of that tile
// is paid while the tiles in between are transformed, so a CTA's own loads are (mostly) L2 hits.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int M, int W, int C, bool COL>
__device__ __forceinline__ void fft_prefetch(const FftArgs& a, int group, int c, int tid) {
    using TL = Tile<M, W, C, COL>;
    constexpr int N = TL::N;
    const int g = group + a.pf_groups;
    if (COL || a.xmap_in || a.pf_groups <= 0 || g >= a.nlines / W) return;   // column pass: the per-row requests cost more L1 wavefronts than they save
    const size_t esz = (a.pro == PRO_REAL_IN) ? sizeof(double) : sizeof(cd);
    const char* base = (const char*)a.in + (size_t)blockIdx.y * a.mstride * esz;
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

template <int M, int W, int C, bool COL, bool NAT, int LD = LD_NONE>
__global__ void __launch_bounds__(W * M / 16, Tile<M, W, C, COL>::MINB)
k_fft_pass(FftArgs a, const __grid_constant__ CUtensorMap tmap) {
    using TL = Tile<M, W, C, COL>;
    constexpr int N = TL::N;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd* smem = reinterpret_cast<cd*>(smem_raw);
    const int tid = threadIdx.x;
    int w, j;
    if (COL) { w = tid % W; j = tid / W; } else { j = tid % TL::TPF; w = tid / TL::TPF; }
    const int c = (C > 1) ? (int)(blockIdx.x % C) : 0;     // rank in the (C,1,1) cluster
    const int group = blockIdx.x / C;
    const int line = group * W + w;
    const size_t mbase = (size_t)blockIdx.y * a.mstride;
    // stage twiddles -> shared memory: L1 is invalidated by every cluster-scope acquire on this SM, and a global
    // twiddle load sits on the critical path of every stage
    cd* smtw = smem + (size_t)W * TL::LINE;
    cd v[fftc::E];
    fft_prefetch<M, W, C, COL>(a, group, c, tid);
    bool tma_done = false;
    if constexpr (LD == LD_WAVEPV) {
        // forward row pass of the wave-PV pair: the tile is FORMED from phi, phix, phiy while it is loaded
        // (CoupledModel.py:75-90; same arithmetic as k_phys_wavepv), four points (12 loads) in flight per thread
        static_assert(!COL && NAT && C == 1, "loader passes are one-tile row passes");
        const size_t i0 = mbase + (size_t)group * W * N + (size_t)w * M + j;
        const cd* __restrict__ p0 = (const cd*)a.in + i0;
        const cd* __restrict__ p1 = a.in2 + i0;
        const cd* __restrict__ p2 = a.in3 + i0;
#pragma unroll
        for (int b = 0; b < fftc::E; b += 4) {
            cd x0[4], x1[4], x2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                x0[i] = __ldg(p0 + (b + i) * TL::TPF); x1[i] = __ldg(p1 + (b + i) * TL::TPF); x2[i] = __ldg(p2 + (b + i) * TL::TPF);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double im = x1[i].x * x2[i].y - x1[i].y * x2[i].x;
                v[b + i] = make_double2(x0[i].x * x0[i].x + x0[i].y * x0[i].y, a.ld_scale * (-2.0 * im));
            }
        }
        for (int t = tid; t < TL::TWLEN; t += TL::T) smtw[t] = a.tw[t];
        tma_done = true;
    }
    if constexpr (LD == LD_UQVQ) {
        static_assert(!COL && NAT && C == 1, "loader passes are one-tile row passes");
        const size_t i0 = mbase + (size_t)group * W * N + (size_t)w * M + j;
        const cd* __restrict__ p0 = (const cd*)a.in + i0;
        const cd* __restrict__ p1 = a.in2 + i0;
#pragma unroll
        for (int b = 0; b < fftc::E; b += 8) {
            cd x0[8], x1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { x0[i] = __ldg(p0 + (b + i) * TL::TPF); x1[i] = __ldg(p1 + (b + i) * TL::TPF); }
#pragma unroll
            for (int i = 0; i < 8; ++i) v[b + i] = make_double2(x0[i].x * x1[i].x, x0[i].y * x1[i].x);
        }
        for (int t = tid; t < TL::TWLEN; t += TL::T) smtw[t] = a.tw[t];
        tma_done = true;
    }
    if constexpr (LD == LD_NONE && COL && NAT && C == 1 && M >= 256) {
        if (a.tma_in & 1) {
            // the whole W x M tile by TMA: boxes of 256 rows x (W*16) bytes land densely ([row][W]) in the exchange buffer
            const unsigned bar = smem_u32(smtw + TL::TWLEN);
            if (tid == 0) {
                mbar_init(bar, 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(bar, (unsigned)(W * M * sizeof(cd)));
#pragma unroll
                for (int b4 = 0; b4 < M / 256; ++b4)
                    tma_load_2d(smem_u32(smem + (size_t)b4 * 256 * W), &tmap, group * W * 2, (int)blockIdx.y * N + b4 * 256, bar);
            }
            for (int t = tid; t < TL::TWLEN; t += TL::T) smtw[t] = a.tw[t];
            mbar_wait(bar, 0);
#pragma unroll
            for (int e = 0; e < fftc::E; ++e) v[e] = smem[(size_t)(j + e * TL::TPF) * W + w];
            __syncthreads();        // the landing area is the exchange buffer of the stages
            tma_done = true;
        }
    }
    if constexpr (LD == LD_NONE && !COL && NAT && C == 1 && M * W == 4096) {
        if ((a.tma_in & 2) && a.pro != PRO_REAL_IN) {
            // row pass, one tile: the W lines of the group are one contiguous 64 KB block - ONE bulk copy into the
            // exchange buffer instead of 16 LDG.128 per thread (A/B against the register loads: profiles/r02_*)
            const unsigned bar = smem_u32(smtw + TL::TWLEN);
            if (tid == 0) {
                mbar_init(bar, 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(bar, (unsigned)(W * M * sizeof(cd)));
                bulk_load(smem_u32(smem), (const cd*)a.in + mbase + (size_t)group * W * N, (unsigned)(W * M * sizeof(cd)), bar);
            }
            for (int t = tid; t < TL::TWLEN; t += TL::T) smtw[t] = a.tw[t];
            mbar_wait(bar, 0);
#pragma unroll
            for (int e = 0; e < fftc::E; ++e) v[e] = smem[(size_t)w * M + j + e * TL::TPF];
            __syncthreads();        // the landing area is the exchange buffer of the stages
            tma_done = true;
        }
    }
    if (!tma_done) {
// all 16 loads are issued back to back (nothing between them depends on loaded data); the prologue multiply and
        // the conjugation of an inverse transform run afterwards, behind ONE uniform branch
        if (a.pro == PRO_REAL_IN) {
            const double* in = (const double*)a.in + mbase;
#pragma unroll
            for (int e = 0; e < fftc::E; ++e) {
                const int m_ = j + e * TL::TPF;
                const int n = (C > 1 && a.deint_in) ? c * M + m_ : C * m_ + c;   // decimated sub-sequence of CTA c
                v[e] = make_double2(in[fft_index<N, COL, NAT>(a, a.xmap_in, line, n)], 0.0);
            }
        } else {
            const cd* in = (const cd*)a.in + mbase;
#pragma unroll
            for (int e = 0; e < fftc::E; ++e) {
                const int m_ = j + e * TL::TPF;
                const int n = (C > 1 && a.deint_in) ? c * M + m_ : C * m_ + c;
                v[e] = in[fft_index<N, COL, NAT>(a, a.xmap_in, line, n)];
            }
        }
        // stage twiddles -> shared memory, AFTER the data loads are out (this store waits for its own global load)
        for (int t = tid; t < TL::TWLEN; t += TL::T) smtw[t] = a.tw[t];
    }
#define NIWQG_PRO_CASE(P)                                                                     \
    case P:                                                                                   \
        _Pragma("unroll") for (int e = 0; e < fftc::E; ++e) {                                 \
            const int n = C * (j + e * TL::TPF) + c;                                          \
            int ky_, kx_;                                                                     \
            fft_kykx<COL>(a, line, n, ky_, kx_);                                              \
            v[e] = fft_prologue_one<N, P>(a, ky_, kx_, v[e]);                                 \
        }                                                                                     \
        break;
    if (a.pro > PRO_REAL_IN) {
        switch (a.pro) {
            NIWQG_PRO_CASE(PRO_IK)
            NIWQG_PRO_CASE(PRO_IL)
            NIWQG_PRO_CASE(PRO_NEG_WV2)
            NIWQG_PRO_CASE(PRO_WV4)
            NIWQG_PRO_CASE(PRO_UV)
            NIWQG_PRO_CASE(PRO_IL_CONJ)
            NIWQG_PRO_CASE(PRO_IK_CONJ)
            default: break;
        }
    }
#undef NIWQG_PRO_CASE
    if (a.conj_in) {
#pragma unroll
        for (int e = 0; e < fftc::E; ++e) v[e].y = -v[e].y;
    }
    fft_stages<M, W, C, COL, NAT, false, 1>(v, j, w, c, smem, smtw, a, line, mbase);
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
                fft_store<N, COL, NAT>(a, mbase, ln, n, v[i * C + p]);
            }
        }
        cluster_wait_relaxed();     // nobody may exit while a peer still reads its shared memory
    }
}

// ---------------------------------------------------------------------------------------------------
// Cluster split, decimation in FREQUENCY with PUSHES (the default for C > 1):
//   CTA c loads x[m + M r], r = 0..C-1, for its m in [c M/C, (c+1) M/C) (whole 64 B row segments in the column
//   pass, two contiguous chunks in the row pass), runs the radix-C butterfly, multiplies Y_q[m] by w_N^{m q} and
//   STORES it into CTA q's shared memory (st.shared::cluster: fire and forget, nobody waits for a remote load).
//   After ONE cluster barrier CTA q holds its whole sub-sequence Y_q[0..M) and transforms it locally:
//   X[C k + q] = FFT_M(Y_q)[k].  No CTA touches a peer's memory after the barrier, so no exit barrier is needed.
// ASYNC: the pushes are st.async stores crediting the receiver's mbarrier (no producer fence, no cluster rendezvous
// after the pushes; one cluster barrier at kernel start makes every mbarrier visible before anybody pushes).
template <int M, int W, int C, bool COL, bool NAT, bool ASYNC>
__global__ void __launch_bounds__(W * M / 16, Tile<M, W, C, COL>::MINB) k_fft_pass_dif(FftArgs a) {
    using TL = Tile<M, W, C, COL>;
    constexpr int N = TL::N;
    static_assert(C > 1 && C <= 8, "cluster sizes 2, 4, 8");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd* smem = reinterpret_cast<cd*>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    int w, j;
    if (COL) { w = tid % W; j = tid / W; } else { j = tid % TL::TPF; w = tid / TL::TPF; }
    const int c = (int)(blockIdx.x % C);
    const int group = blockIdx.x / C;
    const int line = group * W + w;
    const size_t mbase = (size_t)blockIdx.y * a.mstride;
    cd* smtw = smem + (size_t)W * TL::LINE;
    const unsigned bar = smem_u32(smtw + TL::TWLEN);
    if constexpr (ASYNC) {
        if (tid == 0) {
            mbar_init(bar, 1);
            // every peer sends its (M/C) x W block of my sub-sequence; my own block is written with plain stores
            mbar_expect_tx(bar, (unsigned)((C - 1) * (M / C) * W * sizeof(cd)));
        }
        cluster_arrive_release();          // ... and the loads below are in flight while the peers get here
    }
    fft_prefetch<M, W, C, COL>(a, group, c, tid);
    constexpr int PPT = fftc::E / C;                        // radix-C butterflies per thread
    cd v[fftc::E];
    // butterfly i of this thread: line wq of the group, sub-sequence index m; its inputs are x[m + M r]
#define NIWQG_BF(i)                                                        \
    const int g_ = tid + (i) * TL::T;                                      \
    const int wq = COL ? g_ % W : 0;                                       \
    const int m = c * (M / C) + (COL ? g_ / W : g_);                       \
    const int ln = group * W + wq;
    if (a.pro == PRO_REAL_IN) {
        const double* in = (const double*)a.in + mbase;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            NIWQG_BF(i)
#pragma unroll
            for (int r = 0; r < C; ++r) {
                const int n = m + M * r;
                v[i * C + r] = make_double2(in[fft_index<N, COL, NAT>(a, a.xmap_in, ln, n)], 0.0);
            }
        }
    } else {
        const cd* in = (const cd*)a.in + mbase;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            NIWQG_BF(i)
#pragma unroll
            for (int r = 0; r < C; ++r) {
                const int n = m + M * r;
                v[i * C + r] = in[fft_index<N, COL, NAT>(a, a.xmap_in, ln, n)];
            }
        }
    }
    cd wk[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        NIWQG_BF(i)
        wk[i] = a.twc[m];                                   // w_N^m
        (void)wq; (void)ln;
    }
    // stage twiddles -> shared memory, AFTER the data loads are out (this store waits for its own global load)
    for (int t = tid; t < TL::TWLEN; t += TL::T) smtw[t] = a.tw[t];
#define NIWQG_PRO_CASE(P)                                                                         \
    case P:                                                                                       \
        _Pragma("unroll") for (int i = 0; i < PPT; ++i) {                                         \
            NIWQG_BF(i)                                                                           \
            _Pragma("unroll") for (int r = 0; r < C; ++r) {                                       \
                const int n = m + M * r;                                                          \
                int ky_, kx_;                                                                     \
                fft_kykx<COL>(a, ln, n, ky_, kx_);                                                \
                v[i * C + r] = fft_prologue_one<N, P>(a, ky_, kx_, v[i * C + r]);                 \
            }                                                                                     \
        }                                                                                         \
        break;
    if (a.pro > PRO_REAL_IN) {
        switch (a.pro) {
            NIWQG_PRO_CASE(PRO_IK)
            NIWQG_PRO_CASE(PRO_IL)
            NIWQG_PRO_CASE(PRO_NEG_WV2)
            NIWQG_PRO_CASE(PRO_WV4)
            NIWQG_PRO_CASE(PRO_UV)
            NIWQG_PRO_CASE(PRO_IL_CONJ)
            NIWQG_PRO_CASE(PRO_IK_CONJ)
            default: break;
        }
    }
#undef NIWQG_PRO_CASE
    if (a.conj_in) {
#pragma unroll
        for (int e = 0; e < fftc::E; ++e) v[e].y = -v[e].y;
    }
    if constexpr (ASYNC) cluster_wait_acquire();            // all mbarriers of the cluster are initialised
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        NIWQG_BF(i)
        (void)ln;
        fftc::dft<C, 1>(v + i * C);
        cd pw[C];                                           // w_N^{m q}, q = 0..C-1
        pw[0] = make_double2(1.0, 0.0);
        pw[1] = wk[i];
#pragma unroll
        for (int q = 2; q < C; ++q) pw[q] = cmul(pw[q >> 1], pw[q - (q >> 1)]);
#pragma unroll
        for (int p = 0; p < C; ++p) {
            const int q = fftc::outidx<C>(p);
            const cd val = (q == 0) ? v[i * C + p] : cmul(v[i * C + p], pw[q]);
            if constexpr (ASYNC) {
                if (q == c) smem[TL::slot(wq, m)] = val;
                else st_async_cd(cluster_map_u32(smem_u32(smem + TL::slot(wq, m)), q), val, cluster_map_u32(bar, q));
            } else {
                cd* dst = cluster.map_shared_rank(smem, q);
                dst[TL::slot(wq, m)] = val;
            }
        }
    }
#undef NIWQG_BF
    if constexpr (ASYNC) {
        mbar_wait(bar, 0);       // the other C-1 blocks of my sub-sequence have landed
        __syncthreads();         // my own block
    } else {
        cluster_arrive_release();
        cluster_wait_acquire();
    }
#pragma unroll
    for (int e = 0; e < fftc::E; ++e) v[e] = smem[TL::slot(w, j + e * TL::TPF)];
    __syncthreads();
    fft_stages<M, W, C, COL, NAT, true, 1>(v, j, w, c, smem, smtw, a, line, mbase);
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled tma_encoder() {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    return (PFN_tmapEncodeTiled)f;
}

// ---------------------------------------------------------------------------------------------------
// Three-pass column transform for the longest lines (natural layout): the radix-R butterfly across the R row blocks of
// M = N/R rows is a pure streaming pass (no shared memory, 1 KB-coalesced, out of place into a scratch array), and the
// R x (N/W) remaining M-point transforms fit one tile each (full 128 B rows, no cluster, no DSMEM).  One more trip
// through HBM than the cluster kernel, but both passes run at the one-tile rate.
template <int N, int R>
__global__ void __launch_bounds__(256) k_col_radix(FftArgs a) {
    constexpr int M = N / R;
    const int col = blockIdx.x * 256 + threadIdx.x, m = blockIdx.y;
    const size_t mbase = (size_t)blockIdx.z * a.mstride;
    const cd* in = (const cd*)a.in + mbase;
    cd* out = (cd*)a.out + mbase;
    cd v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = in[(size_t)(m + M * r) * N + col];
    const cd w1 = a.twc[m];                                  // w_N^m
#define NIWQG_PRO_CASE(P)                                                                     \
    case P:                                                                                   \
        _Pragma("unroll") for (int r = 0; r < R; ++r) v[r] = fft_prologue_one<N, P>(a, m + M * r, col, v[r]); \
        break;
    if (a.pro > PRO_REAL_IN) {
        switch (a.pro) {
            NIWQG_PRO_CASE(PRO_IK)
            NIWQG_PRO_CASE(PRO_IL)
            NIWQG_PRO_CASE(PRO_NEG_WV2)
            NIWQG_PRO_CASE(PRO_WV4)
            NIWQG_PRO_CASE(PRO_UV)
            NIWQG_PRO_CASE(PRO_IL_CONJ)
            default: break;
        }
    }
#undef NIWQG_PRO_CASE
    if (a.conj_in) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[r].y = -v[r].y;
    }
    fftc::dft<R, 1>(v);
    cd u[R];
#pragma unroll
    for (int p = 0; p < R; ++p) u[fftc::outidx<R>(p)] = v[p];
    fftc::apply_twiddles<R, 1>(u, w1);                       // Y_q[m] *= w_N^{m q}
#pragma unroll
    for (int q = 0; q < R; ++q) out[(size_t)(m + M * q) * N + col] = u[q];
}

// second half: CTA (group, q) transforms the M rows [q M, (q+1) M) of its W columns and stores X[R k + q]
template <int M, int W, int R>
__global__ void __launch_bounds__(W * M / 16, 2) k_fft_colsub(FftArgs a) {
    using TL = Tile<M, W, R, true>;
    constexpr int N = TL::N;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd* smem = reinterpret_cast<cd*>(smem_raw);
    cd* smtw = smem + (size_t)W * TL::LINE;
    const int tid = threadIdx.x, w = tid % W, j = tid / W;
    const int q = (int)(blockIdx.x % R), group = (int)(blockIdx.x / R);
    const int line = group * W + w;
    const size_t mbase = (size_t)blockIdx.y * a.mstride;
    const cd* in = (const cd*)a.in + mbase;
    cd v[fftc::E];
#pragma unroll
    for (int e = 0; e < fftc::E; ++e) v[e] = in[(size_t)(q * M + j + e * TL::TPF) * N + line];
    for (int t = tid; t < TL::TWLEN; t += TL::T) smtw[t] = a.tw[t];
    fft_stages<M, W, R, true, true, true, 1>(v, j, w, q, smem, smtw, a, line, mbase);
}

template <int N, int R, int W>
static cudaError_t launch_col3(const FftArgs& a, void* scratch, int batch, cudaStream_t st) {
    constexpr int M = N / R;
    using TL = Tile<M, W, R, true>;
    // the opt-in to > 48 KB of dynamic shared memory is per DEVICE: one flag per device ordinal (a handle may be
    // created on any device of the process)
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_fft_colsub<M, W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL::SMEM);
        if (e != cudaSuccess) return e;
        if (dev < 64) attr_set[dev] = true;
    }
    FftArgs b = a;               // pass A: prologue + conj on load, no epilogue
    b.out = scratch; b.epi = EPI_NONE; b.scale = 1.0; b.scale_im = 1.0;
    k_col_radix<N, R><<<dim3(N / 256, M, batch), 256, 0, st>>>(b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    FftArgs c = a;               // pass B: plain transform + the pass's epilogue
    c.in = scratch; c.pro = PRO_NONE; c.conj_in = 0;
    k_fft_colsub<M, W, R><<<dim3((N / W) * R, batch), TL::T, TL::SMEM, st>>>(c);
    return cudaGetLastError();
}

// ---- pass geometry: (M, W, C) per grid size
#ifndef NIWQG_COL_M
#define NIWQG_COL_M 1024      // local transform length of a column pass
#endif
#ifndef NIWQG_ROW_MAXM
#define NIWQG_ROW_MAXM 4096   // longest line one CTA transforms alone in a row pass (longer lines: cluster)
#endif
#ifndef NIWQG_COL_TILE
#define NIWQG_COL_TILE 4096   // points per CTA of a column pass
#endif
template <int N, bool COL> struct PassCfg {
    // column pass: W = 4 adjacent columns (64 B per row access), 1024-point local transforms for N >= 1024
    // row pass: whole line per CTA up to 4096 points, two CTAs per line at 8192
    static constexpr int M = COL ? (N >= NIWQG_COL_M ? NIWQG_COL_M : N) : (N > NIWQG_ROW_MAXM ? NIWQG_ROW_MAXM : N);
    static constexpr int C = N / M;
    static constexpr int W = (COL && N >= NIWQG_COL_M) ? NIWQG_COL_TILE / M : ((M >= 4096) ? 1 : (M >= 512) ? 4096 / M : 8);
};

template <int N, bool COL, bool NAT>
static cudaError_t launch_pass_g(const FftArgs& a, int batch, cudaStream_t st) {
    constexpr int M = PassCfg<N, COL>::M, W = PassCfg<N, COL>::W, C = PassCfg<N, COL>::C;
    using TL = Tile<M, W, C, COL>;
    constexpr size_t HALF_SM = 116 * 1024;      // more than half of the 227 KB an SM can give to CTAs
    constexpr size_t SMEM_MAX = TL::SMEM > HALF_SM ? TL::SMEM : HALF_SM;
    static bool attr_set[64] = {};     // per device ordinal (see launch_col3)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_fft_pass<M, W, C, COL, NAT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)SMEM_MAX);
        if (e != cudaSuccess) return e;
        if constexpr (C > 1) {
            e = cudaFuncSetAttribute(k_fft_pass_dif<M, W, C, COL, NAT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(k_fft_pass_dif<M, W, C, COL, NAT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
            if (e != cudaSuccess) return e;
        }
        if (dev < 64) attr_set[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((a.nlines / W) * C, batch, 1);
    cfg.blockDim = dim3(TL::T, 1, 1);
    cfg.dynamicSmemBytes = (a.one_cta_per_sm && TL::SMEM < HALF_SM) ? HALF_SM : TL::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = (C > 1) ? 1 : 0;
    FftArgs b = a;
    if (a.nlines % W) return cudaErrorInvalidValue;
    b.pf_groups = (a.pf_groups > 0 && (a.nlines / W) * C > 2 * a.pf_groups) ? (a.pf_groups + C - 1) / C : 0;   // CTAs -> groups
    if constexpr (C > 1) {
        if (COL ? !(a.variant & 1) : (a.deint_out != 0)) {
            if (a.variant & 4) return cudaLaunchKernelEx(&cfg, k_fft_pass_dif<M, W, C, COL, NAT, false>, b);
            return cudaLaunchKernelEx(&cfg, k_fft_pass_dif<M, W, C, COL, NAT, true>, b);
        }
    }
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    b.tma_in = (!COL && NAT) ? (a.tma_in & 2) : 0;      // bit 1: row tiles by one bulk copy (k_fft_pass)
    if constexpr (COL && NAT && C == 1 && M >= 256 && W * sizeof(cd) < 128) {   // full 128 B rows (W = 8) are as fast with LDG
        if ((a.tma_in & 1) && a.pro != PRO_REAL_IN) {
            // 2-D view of the input: inner dimension = one grid row as doubles, outer = all rows of all members
            static PFN_tmapEncodeTiled enc = tma_encoder();
            if (enc) {
                const cuuint64_t dims[2] = {(cuuint64_t)2 * N, (cuuint64_t)N * batch};
                const cuuint64_t strides[1] = {(cuuint64_t)N * sizeof(cd)};
                const cuuint32_t box[2] = {2 * W, 256};
                const cuuint32_t estr[2] = {1, 1};
                if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(a.in), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                    b.tma_in = 1;
            }
        }
    }
    return cudaLaunchKernelEx(&cfg, k_fft_pass<M, W, C, COL, NAT>, b, tmap);
}

template <int N, bool COL>
static cudaError_t launch_pass_n(const FftArgs& a, int batch, cudaStream_t st) {
    // natural single-GPU geometry gets the kernels with compile-time strides
    const bool nat = a.pitch == N && !a.push && !a.xmap_in && !a.xmap_out && !a.g.sym &&
                     a.mstride == (size_t)N * N;
    return nat ? launch_pass_g<N, COL, true>(a, batch, st) : launch_pass_g<N, COL, false>(a, batch, st);
}

// local transform length of a pass (the stage twiddle table to bind)
static inline int pass_local_len(int N, bool col) { return col ? (N >= NIWQG_COL_M ? NIWQG_COL_M : N) : (N > NIWQG_ROW_MAXM ? NIWQG_ROW_MAXM : N); }

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
