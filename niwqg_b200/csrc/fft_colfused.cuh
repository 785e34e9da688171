// fft_colfused.cuh -- column pass of the longest lines (8192 rows, natural layout) as ONE persistent data-flow kernel
// whose intermediate never leaves L2.
//
// The 8192-point column transform is split 16 x 512 (decimation in frequency):
//   P items ("producers"): a streaming radix-16 butterfly across the 16 row blocks of 512 rows + the w_N^{m q} twiddles,
//       16 rows x 1 column per thread, read from HBM, written to a small SCRATCH RING in tile order;
//   C items ("consumers"): one-tile 512-point transforms of 8 adjacent columns (full 128 B rows), read from the ring,
//       written to the output array X[16 k + q].
// Launched as two kernels (k_col_radix + k_fft_colsub, fft2d.cuh) the intermediate makes a round trip through HBM
// (4.2 GB of DRAM traffic per 8192^2 pass against 2.1 GB algorithmic).  Here both item kinds are handed out by ONE
// ticket counter to the CTAs of a persistent grid, super-block by super-block of CW columns, consumers trailing the
// producers by D super-blocks: an intermediate value is re-read a few microseconds after it was written, while it is
// still in the 126 MB L2 (measured on B200: L2-resident copy 12.5 TB/s vs 6.1 TB/s from HBM), and the ring (NSLOT
// super-blocks of 8 MB) is overwritten in place, so it is never written back.  DRAM traffic = one read of the input
// + one write of the output.
//
// Dependencies are per super-block counters in global memory (release: __threadfence + atomicAdd by one thread after
// a CTA barrier; acquire: ld.acquire spin by one thread, then a CTA barrier).  Tickets are taken in dependency order
// and an item only ever waits for items with SMALLER tickets, which were taken by CTAs that are running: no deadlock,
// whatever number of CTAs is resident (another lane's kernel may share the GPU).  The last CTA to leave resets the
// counters, so every launch (and every CUDA-graph replay) starts from zero.
#pragma once
#include "fft2d.cuh"

struct ColFusedArgs {
    cd* ring;          // NSLOT * N * CW elements, tile order [slot][q][group][row][w]
    unsigned* ctr;     // [0] ticket, [1] exit count, [2 .. 2+NSB) producers done, [2+NSB .. 2+2 NSB) consumers done
    int nslot;
    int delay;         // consumers trail the producers by this many super-blocks (>= 1, < nslot)
    int hints;         // bit 0: streaming (evict-first) loads of the input / stores of the output
    unsigned* stats;   // optional debug counters: [0] items whose dependency was not met when first looked at, [1] of those, still not met at the item boundary
    int tma;           // consumers fetch their tile with one bulk copy (cp.async.bulk -> UBLKCP) instead of 16 x LDG.128
};

template <int N, int CW, int REP = 1> struct ColFused {
    static constexpr int R = 16, M = N / R, W = 8;
    static constexpr int NSB = N / CW;            // super-blocks
    static constexpr int G = CW / W;              // column groups (tiles per row block) of a super-block
    static constexpr int RM = 256 / CW;           // rows m per producer item (256 threads = CW columns x RM rows)
    static constexpr int PI = M / RM / REP;       // producer items per super-block (REP row groups each)
    static constexpr int CI = G * R / REP;        // consumer items per super-block (REP tiles each)
    static constexpr int TOTAL = NSB * (PI + CI);
    static constexpr size_t SLOT = (size_t)N * CW;   // elements per ring slot
    static_assert(PI == CI, "both item kinds move REP x 64 KB in and out");
    static_assert(CW <= 256 && CW % W == 0 && 256 % CW == 0, "super-block width");
    using TL = Tile<M, W, R, true>;
    static constexpr size_t SMEM = TL::SMEM;
    static constexpr int NCTR = 2 + 2 * NSB;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Non-blocking look at a counter (the acquire form waits for the value and then invalidates L1, which stalls the issuing
// warp for an L2 round trip).  Everything that is read after a counter was seen complete is read from L2 (ld.global.cg /
// bulk copy), never from L1, and the producer's __threadfence put the data into L2 before it bumped the counter.
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ cd ld_stream(const cd* p) {     // read once: do not keep the line in L2 longer than needed
    cd v;
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
// ticket -> (kind, super-block, index): step t holds the producers of super-block t interleaved with the consumers of
// super-block t - D
template <int N, int CW, int REP>
__device__ __forceinline__ void col_fused_decode(int ticket, int D, bool& prod, int& sb, int& idx) {
    using CF = ColFused<N, CW, REP>;
    constexpr int PI = CF::PI, CI = CF::CI, NSB = CF::NSB;
    if (ticket < D * PI) { prod = true; sb = ticket / PI; idx = ticket % PI; return; }
    const int t2 = ticket - D * PI, step = t2 / (PI + CI) + D, r = t2 % (PI + CI);
    if (step < NSB) { prod = !(r & 1); sb = prod ? step : step - D; idx = r >> 1; }
    else { const int t3 = t2 - (NSB - D) * (PI + CI); prod = false; sb = NSB - D + t3 / CI; idx = t3 % CI; }
}

template <int N, int CW, int REP>
__global__ void __launch_bounds__(256, 2) k_col_fused(FftArgs a, ColFusedArgs f) {
    using CF = ColFused<N, CW, REP>;
    using TL = typename CF::TL;
    constexpr int R = CF::R, M = CF::M, W = CF::W, PI = CF::PI, CI = CF::CI, NSB = CF::NSB, G = CF::G;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd* smem = reinterpret_cast<cd*>(smem_raw);
    cd* smtw = smem + (size_t)W * TL::LINE;
    __shared__ int s_ticket[2], s_wait[2];
    const int tid = threadIdx.x;
    unsigned* done_p = f.ctr + 2;
    unsigned* done_c = f.ctr + 2 + NSB;
    const unsigned bar = smem_u32(smtw + TL::TWLEN);
    for (int t = tid; t < TL::TWLEN; t += 256) smtw[t] = a.tw[t];
    // what thread 0 knows about the dependency of a ticket: the counter it waits for and the value it must reach
    auto dependency = [&](int ticket, const unsigned*& dep, unsigned& target) {
        dep = nullptr; target = 0;
        if (ticket >= CF::TOTAL) return;
        bool np; int nsb, nidx;
        col_fused_decode<N, CW, REP>(ticket, f.delay, np, nsb, nidx);
        if (np) { if (nsb >= f.nslot) { dep = &done_c[nsb - f.nslot]; target = CI; } }   // ring slot free again
        else { dep = &done_p[nsb]; target = PI; }                                          // super-block produced
    };
    if (tid == 0) {
        if (f.tma) {
            mbar_init(bar, 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        const int t0 = (int)atomicAdd(&f.ctr[0], 1u);
        const unsigned* dep; unsigned target;
        dependency(t0, dep, target);
        if (dep) while (ld_acquire_u32(dep) < target) __nanosleep(32);
        s_ticket[0] = t0;
    }
    __syncthreads();
    unsigned tma_phase = 0;
    int pending = -1;          // thread 0: super-block whose "produced" signal is still owed (sent once the stores have landed)
    for (int it = 0;; ++it) {
        // here: the ticket is in the mailbox, its dependency is met, everybody is done with the previous item
        const int ticket = s_ticket[it & 1];
        if (ticket >= CF::TOTAL) break;
        // thread 0 takes the NEXT ticket now and looks at its dependency while this item is being worked on
        int nxt = 0;
        if (tid == 0) nxt = (int)atomicAdd(&f.ctr[0], 1u);
        bool prod;
        int sb, idx;
        col_fused_decode<N, CW, REP>(ticket, f.delay, prod, sb, idx);
        cd* slot = f.ring + (size_t)(sb % f.nslot) * CF::SLOT;
        const unsigned* dep = nullptr;   // thread 0: counter the next item waits for, its target and a first look at it
        unsigned dep_target = 0, dep_seen = 0;
        auto midpoint = [&]() {          // thread 0, once this item's loads have been consumed
            if (pending >= 0) { __threadfence(); atomicAdd(&done_p[pending], 1u); pending = -1; }
            dependency(nxt, dep, dep_target);
            if (dep) dep_seen = ld_relaxed_u32(dep);
        };
        if (prod) {
            const int cl = tid % CW, col = sb * CW + cl;
            const cd* in = (const cd*)a.in;
#pragma unroll 1
            for (int sub = 0; sub < REP; ++sub) {
            const int m = (idx * REP + sub) * CF::RM + tid / CW;
            cd v[R];
            if (f.hints & 1) {
#pragma unroll
                for (int r = 0; r < R; ++r) v[r] = ld_stream(&in[(size_t)(m + M * r) * N + col]);
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) v[r] = in[(size_t)(m + M * r) * N + col];
            }
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
            if (tid == 0 && sub == 0) midpoint();
            cd u[R];
#pragma unroll
            for (int p = 0; p < R; ++p) u[fftc::outidx<R>(p)] = v[p];
            fftc::apply_twiddles<R, 1>(u, w1);                       // Y_q[m] *= w_N^{m q}
            cd* dst = slot + ((size_t)(cl / W) * M + m) * W + (cl % W);
#pragma unroll
            for (int q = 0; q < R; ++q) dst[(size_t)q * G * M * W] = u[q];
            }
            if (tid == 0) pending = sb;
        } else {
            const int w = tid % W, j = tid / W;
#pragma unroll 1
            for (int sub = 0; sub < REP; ++sub) {
            const int q = (idx * REP + sub) % R, gl = (idx * REP + sub) / R;
            const cd* tile = slot + (size_t)(q * G + gl) * M * W;
            const int line = (sb * G + gl) * W + w;
            cd v[fftc::E];
            if (f.tma) {
                if (tid == 0) {
                    asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes (other SMs' stores, my STS) before the bulk copy
                    mbar_expect_tx(bar, (unsigned)(M * W * sizeof(cd)));
                    bulk_load(smem_u32(smem), tile, (unsigned)(M * W * sizeof(cd)), bar);
                }
                mbar_wait(bar, tma_phase);
                tma_phase ^= 1u;
#pragma unroll
                for (int e = 0; e < fftc::E; ++e) v[e] = smem[(size_t)(j + e * TL::TPF) * W + w];
                __syncthreads();        // the landing area is the exchange buffer of the stages
            } else {
#pragma unroll
                for (int e = 0; e < fftc::E; ++e) v[e] = __ldcg(&tile[(size_t)(j + e * TL::TPF) * W + w]);
            }
            if (tid == 0 && sub == 0) midpoint();
            fft_stages<M, W, R, true, true, true, 1>(v, j, w, q, smem, smtw, a, line, 0);
            }
            // every thread has passed a CTA barrier after consuming its loads: the tile has been read
            if (tid == 0) atomicAdd(&done_c[sb], 1u);
        }
        if (tid == 0) {
            s_wait[it & 1] = (dep && dep_seen < dep_target) ? 1 : 0;
            s_ticket[(it + 1) & 1] = nxt;
        }
        __syncthreads();
        if (s_wait[it & 1]) {
            // the next item's dependency is not met yet.  Whoever we wait for may (transitively) wait for our own signal,
            // so it is sent first (every thread's stores of this item are behind the barrier above)
            if (tid == 0) {
                if (f.stats) atomicAdd(&f.stats[0], 1u);
                if (ld_acquire_u32(dep) < dep_target) {
                    if (f.stats) atomicAdd(&f.stats[1], 1u);
                    if (pending >= 0) { __threadfence(); atomicAdd(&done_p[pending], 1u); pending = -1; }
                    while (ld_acquire_u32(dep) < dep_target) __nanosleep(32);
                }
            }
            __syncthreads();
        }
    }
    // the last CTA out resets the counters for the next launch
    if (tid == 0) {
        if (pending >= 0) { __threadfence(); atomicAdd(&done_p[pending], 1u); }
        __threadfence();
        const unsigned prev = atomicAdd(&f.ctr[1], 1u);
        if (prev == gridDim.x - 1) {
            for (int i = 0; i < CF::NCTR; ++i) f.ctr[i] = 0u;
            __threadfence();
        }
    }
}

// ring + counters of one lane
template <int N, int CW> static size_t col_fused_ring_bytes(int nslot) { return (size_t)nslot * ColFused<N, CW>::SLOT * sizeof(cd); }

template <int N, int CW, int REP = 1>
static cudaError_t launch_col_fused(const FftArgs& a, const ColFusedArgs& f, int nctas, cudaStream_t st) {
    using CF = ColFused<N, CW, REP>;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_col_fused<N, CW, REP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::SMEM);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    k_col_fused<N, CW, REP><<<nctas, 256, CF::SMEM, st>>>(a, f);
    return cudaGetLastError();
}
