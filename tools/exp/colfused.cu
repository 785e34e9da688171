// Development aid: k_col_fused (fft_colfused.cuh) against the two-kernel three-pass column transform: bit-exact check + timing.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include "../../niwqg_b200/csrc/fft_colfused.cuh"

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
struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void start(cudaStream_t s) { cudaEventRecord(a, s); }
    float stop(cudaStream_t s) { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
};
__global__ void k_maxdiff(const cd* x, const cd* y, size_t n, unsigned long long* out) {
    unsigned long long bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (x[i].x != y[i].x || x[i].y != y[i].y) ++bad;
    if (bad) atomicAdd(out, bad);
}

constexpr int N = 8192;
static cd *A, *B, *C, *S, *ring;
static unsigned* ctr;
static unsigned long long* d_bad;
static cudaStream_t st;
static Timer T;

template <int CW, int REP = 1>
static void run_variant(const FftArgs& a, int nslot, int hints, int tma, int nctas, const char* tag, int delay = 2) {
    ColFusedArgs f{};
    f.ring = ring; f.ctr = ctr; f.stats = ctr + 2048; f.nslot = nslot; f.delay = delay; f.hints = hints; f.tma = tma;
    FftArgs b = a; b.in = A; b.out = C;
    CKE(cudaMemsetAsync(C, 0, (size_t)N * N * sizeof(cd), st));
    CKE(cudaMemsetAsync(ctr + 2048, 0, 64, st));
    float best = 1e9;
    for (int rep = 0; rep < 6; ++rep) {
        T.start(st);
        CKE((launch_col_fused<N, CW, REP>(b, f, nctas, st)));
        float ms = T.stop(st);
        if (rep) best = std::min(best, ms);
    }
    CKE(cudaMemsetAsync(d_bad, 0, 8, st));
    k_maxdiff<<<1184, 256, 0, st>>>(B, C, (size_t)N * N, d_bad);
    unsigned long long bad = 0;
    CKE(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
    unsigned stats[2];
    CKE(cudaMemcpyAsync(stats, ctr + 2048, 8, cudaMemcpyDeviceToHost, st));
    CKE(cudaStreamSynchronize(st));
    const double GB = 2.0 * (double)N * N * sizeof(cd) / 1e9;
    printf("   %-14s REP %d CW %3d D %d nslot %2d hints %d tma %d ctas %3d: %.4f ms  (%.0f GB/s algorithmic)  mismatch %llu  early-miss %u late-miss %u per launch (of %d items)\n",
           tag, REP, CW, delay, nslot, hints, tma, nctas, best, GB / (best * 1e-3), bad, stats[0] / 6, stats[1] / 6, ColFused<N, CW, REP>::TOTAL);
    fflush(stdout);
}

int main(int argc, char** argv) {
    const bool quick = argc > 1 && !strcmp(argv[1], "ncu");
    const size_t npts = (size_t)N * N;
    CKE(cudaMalloc(&A, npts * sizeof(cd)));
    CKE(cudaMalloc(&B, npts * sizeof(cd)));
    CKE(cudaMalloc(&C, npts * sizeof(cd)));
    CKE(cudaMalloc(&S, npts * sizeof(cd)));
    CKE(cudaMalloc(&ring, (size_t)16 * N * 256 * sizeof(cd)));   // up to 16 slots of 256 columns (512 MB)
    CKE(cudaMalloc(&ctr, 4096 * sizeof(unsigned)));
    CKE(cudaMemset(ctr, 0, 4096 * sizeof(unsigned)));
    CKE(cudaMalloc(&d_bad, 8));
    {
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
    for (int t = 0; t < N; ++t) { long double x = -2.0L * PI * t / N; twc[t] = make_double2((double)cosl(x), (double)sinl(x)); }
    cd *d_tw, *d_twc;
    CKE(cudaMalloc(&d_tw, tw.size() * sizeof(cd)));
    CKE(cudaMemcpy(d_tw, tw.data(), tw.size() * sizeof(cd), cudaMemcpyHostToDevice));
    CKE(cudaMalloc(&d_twc, N * sizeof(cd)));
    CKE(cudaMemcpy(d_twc, twc.data(), N * sizeof(cd), cudaMemcpyHostToDevice));
    CKE(cudaStreamCreate(&st));

    FftArgs a{};
    a.twc = d_twc; a.tw = d_tw; a.dk = 1e-5; a.pf_groups = 0; a.variant = 6;
    a.g = Grid{N, 1e-5, N, N / 2, 0, 0};
    a.nlines = N; a.pitch = N; a.mstride = npts; a.pro = PRO_NONE; a.epi = EPI_NONE;
    for (int cfg = 0; cfg < 2; ++cfg) {
        if (cfg == 0) { a.conj_in = 0; a.scale = 1.0; a.scale_im = 1.0; a.pro = PRO_NONE; }
        else { a.conj_in = 1; a.scale = 1.0 / 64.0; a.scale_im = -1.0 / 64.0; a.pro = PRO_IL_CONJ; }   // second pass of an inverse transform with a prologue
        printf("== config %d (pro %d conj_in %d)\n", cfg, a.pro, a.conj_in);
        FftArgs b = a; b.in = A; b.out = B;
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            T.start(st);
            CKE((launch_col3<N, 16, 8>(b, S, 1, st)));
            float ms = T.stop(st);
            if (rep) best = std::min(best, ms);
        }
        printf("   two kernels through HBM scratch: %.4f ms\n", best);
        run_variant<64>(a, 4, 0, 0, 296, "base");
        if (quick) { run_variant<64>(a, 4, 1, 0, 296, "hints"); run_variant<128>(a, 4, 1, 0, 296, "cw128"); break; }
        if (cfg == 0) {
            run_variant<64>(a, 4, 1, 0, 296, "hints");
            run_variant<64>(a, 5, 1, 0, 296, "D3", 3);
            run_variant<64>(a, 6, 1, 0, 296, "D3", 3);
            run_variant<64>(a, 6, 1, 0, 296, "D4", 4);
            run_variant<64>(a, 7, 1, 0, 296, "D4", 4);
            run_variant<64>(a, 8, 1, 0, 296, "D5", 5);
            run_variant<64>(a, 8, 1, 0, 296, "D6", 6);
            run_variant<32>(a, 8, 1, 0, 296, "cw32 D6", 6);
            run_variant<32>(a, 12, 1, 0, 296, "cw32 D8", 8);
            run_variant<32>(a, 16, 1, 0, 296, "cw32 D12", 12);
            run_variant<128>(a, 4, 1, 0, 296, "cw128 D2", 2);
            run_variant<128>(a, 5, 1, 0, 296, "cw128 D3", 3);
            run_variant<64, 2>(a, 8, 1, 0, 296, "rep2 D6", 6);
            run_variant<64, 2>(a, 10, 1, 0, 296, "rep2 D8", 8);
            run_variant<64, 4>(a, 12, 1, 0, 296, "rep4 D10", 10);
            run_variant<64>(a, 6, 1, 1, 296, "D4 tma", 4);
            run_variant<64>(a, 6, 0, 0, 296, "D4 nohint", 4);
        }
    }
    // in place (in == out)
    {
        a.conj_in = 0; a.scale = 1.0; a.scale_im = 1.0; a.pro = PRO_NONE;
        FftArgs b = a; b.in = A; b.out = B;
        CKE((launch_col3<N, 16, 8>(b, S, 1, st)));
        CKE(cudaMemcpyAsync(C, A, npts * sizeof(cd), cudaMemcpyDeviceToDevice, st));
        ColFusedArgs f{};
        f.ring = ring; f.ctr = ctr; f.stats = nullptr; f.nslot = 4; f.delay = 2; f.hints = 1; f.tma = 0;
        FftArgs c = a; c.in = C; c.out = C;
        CKE((launch_col_fused<N, 64>(c, f, 296, st)));
        CKE(cudaMemsetAsync(d_bad, 0, 8, st));
        k_maxdiff<<<1184, 256, 0, st>>>(B, C, npts, d_bad);
        unsigned long long bad = 0;
        CKE(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
        CKE(cudaStreamSynchronize(st));
        printf("== in place: mismatching elements %llu\n", bad);
    }
    CKE(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
