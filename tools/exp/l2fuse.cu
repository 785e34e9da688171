// Experiment (development aid, not product code): does an intermediate that stays in L2 make the three-pass column
// transform of 8192^2 cheaper?  Runs the existing k_col_radix / k_fft_colsub kernels over column super-blocks, in place,
// and a few bandwidth probes.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo
//   -I<nccl include> -o tools/exp/l2fuse tools/exp/l2fuse.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include "../../niwqg_b200/csrc/fft2d.cuh"

#define CKE(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static void build_twiddles(int N, std::vector<cd>& tw) {
    tw.assign(fftc::tw_table_len(N) + 1, make_double2(1.0, 0.0));
    const long double PI = 3.141592653589793238462643383279502884L;
    for (int NS = 16; NS < N; NS *= 16) {
        const int R = (N / NS >= 16) ? 16 : N / NS;
        for (int kk = 0; kk < NS; ++kk) {
            const long double a = -2.0L * PI * (long double)kk / ((long double)NS * R);
            tw[fftc::tw_offset(NS) + kk] = make_double2((double)cosl(a), (double)sinl(a));
        }
    }
}

// ---- bandwidth probes: copy / read over a working set of `n` cd elements, `iters` sweeps inside one launch
__global__ void k_copy(const cd* __restrict__ in, cd* __restrict__ out, size_t n, int iters) {
    for (int it = 0; it < iters; ++it)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            cd v = __ldcg(in + i);
            v.x += 1.0;
            __stcg(out + i, v);
        }
}
__global__ void k_read(const cd* __restrict__ in, size_t n, int iters, double* sink) {
    double s = 0;
    for (int it = 0; it < iters; ++it)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            cd v = __ldcg(in + i);
            s += v.x + v.y;
        }
    if (s == 1.2345e300) *sink = s;
}

// ---- super-block variants of the three-pass column transform (natural layout, N = 8192, R = 16, M = 512)
template <int N, int R>
__global__ void __launch_bounds__(256) k_col_radix_sb(FftArgs a, int col0, int TC) {
    constexpr int M = N / R;
    const int RM = 256 / TC;
    const int col = col0 + blockIdx.x * TC + (threadIdx.x % TC), m = blockIdx.y * RM + threadIdx.x / TC;
    const cd* in = (const cd*)a.in;
    cd* out = (cd*)a.out;
    cd v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = __ldcg(&in[(size_t)(m + M * r) * N + col]);
    const cd w1 = a.twc[m];
    if (a.conj_in) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[r].y = -v[r].y;
    }
    fftc::dft<R, 1>(v);
    cd u[R];
#pragma unroll
    for (int p = 0; p < R; ++p) u[fftc::outidx<R>(p)] = v[p];
    fftc::apply_twiddles<R, 1>(u, w1);
#pragma unroll
    for (int q = 0; q < R; ++q) out[(size_t)(m + M * q) * N + col] = u[q];
}

template <int M, int W, int R>
__global__ void __launch_bounds__(W * M / 16, 2) k_fft_colsub_sb(FftArgs a, int g0) {
    using TL = Tile<M, W, R, true>;
    constexpr int N = TL::N;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd* smem = reinterpret_cast<cd*>(smem_raw);
    cd* smtw = smem + (size_t)W * TL::LINE;
    const int tid = threadIdx.x, w = tid % W, j = tid / W;
    const int q = (int)(blockIdx.x % R), group = g0 + (int)(blockIdx.x / R);
    const int line = group * W + w;
    const cd* in = (const cd*)a.in;
    cd v[fftc::E];
#pragma unroll
    for (int e = 0; e < fftc::E; ++e) v[e] = __ldcg(&in[(size_t)(q * M + j + e * TL::TPF) * N + line]);
    for (int t = tid; t < TL::TWLEN; t += TL::T) smtw[t] = a.tw[t];
    fft_stages<M, W, R, true, true, true, 1>(v, j, w, q, smem, smtw, a, line, 0);
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void start(cudaStream_t s) { cudaEventRecord(a, s); }
    float stop(cudaStream_t s) { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
};

int main(int argc, char** argv) {
    constexpr int N = 8192, R = 16, M = N / R, W = 8;
    using TL = Tile<M, W, R, true>;
    const size_t npts = (size_t)N * N;
    cd *A, *B, *S;
    CKE(cudaMalloc(&A, npts * sizeof(cd)));
    CKE(cudaMalloc(&B, npts * sizeof(cd)));
    CKE(cudaMalloc(&S, npts * sizeof(cd)));
    {   // pseudo-random input
        std::vector<cd> h(npts);
        unsigned long long s = 88172645463325252ULL;
        for (size_t i = 0; i < npts; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            h[i] = make_double2((double)(s & 0xffffff) / 16777216.0 - 0.5, (double)((s >> 24) & 0xffffff) / 16777216.0 - 0.5);
        }
        CKE(cudaMemcpy(A, h.data(), npts * sizeof(cd), cudaMemcpyHostToDevice));
    }
    std::vector<cd> tw, twc(N);
    build_twiddles(512, tw);
    const long double PI = 3.141592653589793238462643383279502884L;
    for (int t = 0; t < N; ++t) { long double a = -2.0L * PI * t / N; twc[t] = make_double2((double)cosl(a), (double)sinl(a)); }
    cd *d_tw, *d_twc, *d_tw_row;
    CKE(cudaMalloc(&d_tw, tw.size() * sizeof(cd)));
    CKE(cudaMemcpy(d_tw, tw.data(), tw.size() * sizeof(cd), cudaMemcpyHostToDevice));
    CKE(cudaMalloc(&d_twc, N * sizeof(cd)));
    CKE(cudaMemcpy(d_twc, twc.data(), N * sizeof(cd), cudaMemcpyHostToDevice));
    std::vector<cd> tw4;
    build_twiddles(4096, tw4);
    CKE(cudaMalloc(&d_tw_row, tw4.size() * sizeof(cd)));
    CKE(cudaMemcpy(d_tw_row, tw4.data(), tw4.size() * sizeof(cd), cudaMemcpyHostToDevice));
    cudaStream_t st, st2;
    CKE(cudaStreamCreate(&st));
    CKE(cudaStreamCreate(&st2));
    Timer T;
    const double GB = 2.0 * npts * sizeof(cd) / 1e9;   // algorithmic bytes of one pass

    // ---------------- E1: bandwidth probes
    printf("== E1 bandwidth probes (copy = ld.cg + st.cg, bytes counted read+write; read = ld.cg only)\n");
    for (size_t mb : {8, 16, 32, 48, 64, 96, 128, 1024}) {
        const size_t n = mb * 1024 * 1024 / sizeof(cd) / 2;    // half in, half out -> working set = mb
        const int iters = (int)std::max<size_t>(1, 4096 / mb);
        k_copy<<<148 * 8, 256, 0, st>>>(A, A + n, n, 2);       // warm
        T.start(st);
        k_copy<<<148 * 8, 256, 0, st>>>(A, A + n, n, iters);
        float ms = T.stop(st);
        const double gbs = 2.0 * n * sizeof(cd) * iters / (ms * 1e-3) / 1e9;
        const size_t nr = mb * 1024 * 1024 / sizeof(cd);
        k_read<<<148 * 8, 256, 0, st>>>(B, nr, 2, (double*)S);
        T.start(st);
        k_read<<<148 * 8, 256, 0, st>>>(B, nr, iters, (double*)S);
        float ms2 = T.stop(st);
        const double gbr = (double)nr * sizeof(cd) * iters / (ms2 * 1e-3) / 1e9;
        printf("   working set %5zu MB: copy %8.0f GB/s   read %8.0f GB/s\n", mb, gbs, gbr);
    }
    CKE(cudaGetLastError());
    // re-create the input (the copy probe modified A)
    {
        std::vector<cd> h(npts);
        unsigned long long s = 88172645463325252ULL;
        for (size_t i = 0; i < npts; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            h[i] = make_double2((double)(s & 0xffffff) / 16777216.0 - 0.5, (double)((s >> 24) & 0xffffff) / 16777216.0 - 0.5);
        }
        CKE(cudaMemcpy(A, h.data(), npts * sizeof(cd), cudaMemcpyHostToDevice));
    }

    FftArgs a{};
    a.twc = d_twc; a.tw = d_tw; a.dk = 1.0; a.pf_groups = 0; a.variant = 6;
    a.g = Grid{N, 1.0, N, N / 2, 0, 0};
    a.nlines = N; a.pitch = N; a.mstride = npts; a.scale = 1.0; a.scale_im = 1.0; a.pro = PRO_NONE; a.epi = EPI_NONE;

    // ---------------- E2: reference three-pass (out of place through scratch), as the library runs it
    printf("== E2 three-pass column transform, whole array (library form: A -> S -> B)\n");
    CKE(cudaFuncSetAttribute(k_fft_colsub<M, W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL::SMEM));
    CKE(cudaFuncSetAttribute(k_fft_colsub_sb<M, W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL::SMEM));
    float best = 1e9;
    for (int rep = 0; rep < 6; ++rep) {
        FftArgs b = a; b.in = A; b.out = S;
        FftArgs c = a; c.in = S; c.out = B;
        T.start(st);
        k_col_radix<N, R><<<dim3(N / 256, M, 1), 256, 0, st>>>(b);
        k_fft_colsub<M, W, R><<<dim3((N / W) * R, 1), TL::T, TL::SMEM, st>>>(c);
        float ms = T.stop(st);
        if (rep) best = std::min(best, ms);
    }
    CKE(cudaGetLastError());
    printf("   A->S->B: %.4f ms  (%.0f GB/s algorithmic)\n", best, GB / (best * 1e-3));
    // each half alone
    for (int which = 0; which < 2; ++which) {
        best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            FftArgs b = a; b.in = A; b.out = S;
            FftArgs c = a; c.in = S; c.out = B;
            T.start(st);
            if (which == 0) k_col_radix<N, R><<<dim3(N / 256, M, 1), 256, 0, st>>>(b);
            else k_fft_colsub<M, W, R><<<dim3((N / W) * R, 1), TL::T, TL::SMEM, st>>>(c);
            float ms = T.stop(st);
            if (rep) best = std::min(best, ms);
        }
        printf("   %s alone: %.4f ms (%.0f GB/s)\n", which ? "k_fft_colsub" : "k_col_radix", best, GB / (best * 1e-3));
    }

    // ---------------- E3: super-blocks: radix A->B on CW columns, colsub B->B in place on the same columns
    printf("== E3 super-blocked (radix A->B, colsub in place on B; intermediate re-read while L2-resident)\n");
    std::vector<cd> ref(1 << 16), got(1 << 16);
    CKE(cudaMemcpy(ref.data(), B + 12345 * (size_t)N, ref.size() * sizeof(cd), cudaMemcpyDeviceToHost));
    for (int CW : {64, 128, 256, 512, 1024, 2048, 8192}) {
        for (int mode = 0; mode < 3; ++mode) {      // 0: one stream direct, 1: CUDA graph, 2: two streams alternating
            const int TC = std::min(CW, 256), RM = 256 / TC;
            auto issue = [&](cudaStream_t s0, cudaStream_t s1) {
                for (int c0 = 0, i = 0; c0 < N; c0 += CW, ++i) {
                    cudaStream_t s = (i & 1) ? s1 : s0;
                    FftArgs b = a; b.in = A; b.out = B;
                    k_col_radix_sb<N, R><<<dim3(CW / TC, M / RM, 1), 256, 0, s>>>(b, c0, TC);
                    FftArgs c = a; c.in = B; c.out = B;
                    k_fft_colsub_sb<M, W, R><<<dim3((CW / W) * R, 1), TL::T, TL::SMEM, s>>>(c, c0 / W);
                }
            };
            cudaGraphExec_t gexec = nullptr;
            if (mode == 1) {
                cudaGraph_t g;
                CKE(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                issue(st, st);
                CKE(cudaStreamEndCapture(st, &g));
                CKE(cudaGraphInstantiate(&gexec, g, 0));
                cudaGraphDestroy(g);
            }
            best = 1e9;
            cudaEvent_t fork, join;
            cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
            for (int rep = 0; rep < 5; ++rep) {
                T.start(st);
                if (mode == 0) issue(st, st);
                else if (mode == 1) CKE(cudaGraphLaunch(gexec, st));
                else {
                    cudaEventRecord(fork, st); cudaStreamWaitEvent(st2, fork, 0);
                    issue(st, st2);
                    cudaEventRecord(join, st2); cudaStreamWaitEvent(st, join, 0);
                }
                float ms = T.stop(st);
                if (rep) best = std::min(best, ms);
            }
            CKE(cudaGetLastError());
            CKE(cudaMemcpy(got.data(), B + 12345 * (size_t)N, got.size() * sizeof(cd), cudaMemcpyDeviceToHost));
            double md = 0;
            for (size_t i = 0; i < got.size(); ++i) md = std::max(md, std::max(fabs(got[i].x - ref[i].x), fabs(got[i].y - ref[i].y)));
            printf("   CW %5d (%4d MB/super-block) %-10s: %.4f ms  (%.0f GB/s algorithmic)  maxdiff vs E2 %.1e\n", CW,
                   (int)((size_t)CW * N * 16 >> 20), mode == 0 ? "direct" : mode == 1 ? "graph" : "2 streams", best, GB / (best * 1e-3), md);
            if (gexec) cudaGraphExecDestroy(gexec);
        }
    }

    // ---------------- E5: each half on an L2-resident super-block (same columns over and over): is it faster than from HBM?
    printf("== E5 halves on an L2-resident super-block (20 repeats of the same columns, per-repeat time scaled to the whole array)\n");
    for (int CW : {64, 256, 512}) {
        const int TC = std::min(CW, 256), RM = 256 / TC, reps = 20;
        for (int which = 0; which < 2; ++which) {
            best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                T.start(st);
                for (int i = 0; i < reps; ++i) {
                    FftArgs b = a; b.in = B; b.out = B;
                    if (which == 0) k_col_radix_sb<N, R><<<dim3(CW / TC, M / RM, 1), 256, 0, st>>>(b, 0, TC);
                    else k_fft_colsub_sb<M, W, R><<<dim3((CW / W) * R, 1), TL::T, TL::SMEM, st>>>(b, 0);
                }
                float ms = T.stop(st);
                if (rep) best = std::min(best, ms);
            }
            const double per = best / reps, full = per * (N / CW);
            printf("   CW %4d %-12s: %.2f us per super-block -> %.4f ms per whole array (%.0f GB/s)\n", CW, which ? "k_fft_colsub" : "k_col_radix",
                   per * 1e3, full, GB / (full * 1e-3));
        }
    }
    CKE(cudaGetLastError());

    // ---------------- E4: row passes
    printf("== E4 row passes on 1 GiB\n");
    {
        FftArgs r = a; r.tw = d_tw_row; r.in = A; r.out = B; r.nlines = N; r.pf_groups = 296;
        best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            T.start(st);
            CKE((launch_pass_g<8192, false, true>(r, 1, st)));
            float ms = T.stop(st);
            if (rep) best = std::min(best, ms);
        }
        printf("   8192-point lines, cluster C=2 pull kernel: %.4f ms (%.0f GB/s)\n", best, GB / (best * 1e-3));
        FftArgs r4 = r; r4.nlines = 4096; r4.g = Grid{4096, 1.0, 4096, 2048, 0, 0}; r4.pitch = 4096; r4.mstride = (size_t)4096 * 4096;
        best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            T.start(st);
            CKE((launch_pass_g<4096, false, true>(r4, 4, st)));
            float ms = T.stop(st);
            if (rep) best = std::min(best, ms);
        }
        printf("   4096-point lines (16384 of them), one tile per CTA: %.4f ms (%.0f GB/s)\n", best, GB / (best * 1e-3));
    }
    CKE(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
