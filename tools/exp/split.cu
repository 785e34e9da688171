// Development aid: the split 2-D transform (fft_split.cuh) against the cluster / three-pass path: accuracy + per-kernel timing.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include "../../niwqg_b200/csrc/fft_split.cuh"
#include "../../niwqg_b200/csrc/kernels_family.cuh"

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
static cd* upload_tw(int n) {
    std::vector<cd> tw;
    build_twiddles(n, tw);
    cd* d;
    CKE(cudaMalloc(&d, tw.size() * sizeof(cd)));
    CKE(cudaMemcpy(d, tw.data(), tw.size() * sizeof(cd), cudaMemcpyHostToDevice));
    return d;
}
struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void start(cudaStream_t s) { cudaEventRecord(a, s); }
    float stop(cudaStream_t s) { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
};
__global__ void k_err(const cd* x, const cd* y, size_t n, double* out) {   // out[0] += |x-y|^2, out[1] += |y|^2
    double e = 0, r = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double dx = x[i].x - y[i].x, dy = x[i].y - y[i].y;
        e += dx * dx + dy * dy; r += y[i].x * y[i].x + y[i].y * y[i].y;
    }
    atomicAdd(out, e); atomicAdd(out + 1, r);
}

template <int N> static void run() {
    constexpr int Nh = N / 2, M = N / 16;
    const size_t npts = (size_t)N * N;
    cd *A, *Ad, *REF, *T1, *T2, *OUT[3], *S;
    CKE(cudaMalloc(&A, npts * sizeof(cd))); CKE(cudaMalloc(&Ad, npts * sizeof(cd))); CKE(cudaMalloc(&REF, npts * sizeof(cd)));
    CKE(cudaMalloc(&T1, npts * sizeof(cd))); CKE(cudaMalloc(&T2, npts * sizeof(cd))); CKE(cudaMalloc(&S, npts * sizeof(cd)));
    for (int o = 0; o < 3; ++o) CKE(cudaMalloc(&OUT[o], npts * sizeof(cd)));
    double* d_err;
    CKE(cudaMalloc(&d_err, 16));
    {
        std::vector<cd> h(npts);
        unsigned long long s = 88172645463325252ULL;
        for (size_t i = 0; i < npts; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            h[i] = make_double2((double)(s & 0xffffff) / 16777216.0 - 0.5, (double)((s >> 24) & 0xffffff) / 16777216.0 - 0.5);
        }
        CKE(cudaMemcpy(A, h.data(), npts * sizeof(cd), cudaMemcpyHostToDevice));
    }
    std::vector<cd> twc(N);
    const long double PI = 3.141592653589793238462643383279502884L;
    for (int t = 0; t < N; ++t) { long double x = -2.0L * PI * t / N; twc[t] = make_double2((double)cosl(x), (double)sinl(x)); }
    cd* d_twc;
    CKE(cudaMalloc(&d_twc, N * sizeof(cd)));
    CKE(cudaMemcpy(d_twc, twc.data(), N * sizeof(cd), cudaMemcpyHostToDevice));
    cd* tw_row_old = upload_tw(pass_local_len(N, false));
    cd* tw_col_old = upload_tw(pass_local_len(N, true));
    cd* tw_c3 = upload_tw(512);
    cd* tw_half = upload_tw(Nh);
    cd* tw_m = upload_tw(M);
    cudaStream_t st;
    CKE(cudaStreamCreate(&st));
    Timer T;
    const double dk = 2.0 * M_PI / 5e5;
    const double GB = 2.0 * npts * sizeof(cd) / 1e9;
    auto err = [&](const cd* x, const cd* y) {
        CKE(cudaMemsetAsync(d_err, 0, 16, st));
        k_err<<<1184, 256, 0, st>>>(x, y, npts, d_err);
        double h[2];
        CKE(cudaMemcpyAsync(h, d_err, 16, cudaMemcpyDeviceToHost, st));
        CKE(cudaStreamSynchronize(st));
        return sqrt(h[0] / h[1]);
    };
    auto timeit = [&](const char* what, auto&& f) {
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) { T.start(st); f(); float ms = T.stop(st); if (rep) best = std::min(best, ms); }
        CKE(cudaGetLastError());
        printf("      %-44s %.4f ms  (%.0f GB/s at 32 B/pt)\n", what, best, GB / (best * 1e-3));
        return best;
    };
    FftArgs base{};
    base.twc = d_twc; base.dk = dk; base.pf_groups = 296; base.variant = 6; base.g = Grid{N, dk, N, N / 2, 0, 0};
    base.nlines = N; base.pitch = N; base.mstride = npts; base.scale = 1.0; base.scale_im = 1.0; base.pro = PRO_NONE; base.epi = EPI_NONE;
    base.tma_in = 1;
    // old path pieces
    auto old_row = [&](const cd* in, cd* out, int pro, int conj_in) {
        FftArgs a = base; a.in = in; a.out = out; a.pro = pro; a.conj_in = conj_in; a.tw = tw_row_old;
        CKE((launch_pass_g<N, false, true>(a, 1, st)));
    };
    auto old_col = [&](const cd* in, cd* out, int pro, double sc, int conj_out) {
        FftArgs a = base; a.in = in; a.out = out; a.pro = pro; a.tw = tw_col_old; a.scale = sc; a.scale_im = conj_out ? -sc : sc;
        if (N == 8192) { a.tw = tw_c3; CKE((launch_col3<8192, 16, 8>(a, T2, 1, st))); }
        else CKE((launch_pass_g<N, true, true>(a, 1, st)));
    };
    // new path pieces
    int row_bulk = 0;
    auto new_rows = [&](const cd* in, cd* out, double sc, int conj_out) {
        FftArgs a = base; a.in = in; a.out = out; a.tw = tw_half; a.nlines = 2 * N; a.pitch = Nh; a.mstride = (size_t)Nh * Nh;
        a.g = Grid{Nh, dk, Nh, Nh / 2, 0, 0}; a.scale = sc; a.scale_im = conj_out ? -sc : sc; a.tma_in = row_bulk ? 2 : 0;
        CKE((launch_pass_g<Nh, false, true>(a, 1, st)));
    };
    auto new_colsub = [&](const cd* in, cd* out, bool dit) {
        FftArgs a = base; a.in = in; a.out = out; a.tw = tw_m;
        if (dit) CKE((launch_split_colsub<N, true>(a, st))); else CKE((launch_split_colsub<N, false>(a, st)));
    };
    auto new_p = [&](const cd* in, cd** out, const int* pro, int nout, bool dit, int conj_in) {
        SplitPArgs p{};
        p.in = in; p.nout = nout; p.conj_in = conj_in; p.scale = 1.0; p.dk = dk; p.twc = d_twc;
        for (int o = 0; o < nout; ++o) { p.out[o] = out[o]; p.pro[o] = pro[o]; }
        if (dit) CKE((launch_split_p<N, true>(p, st))); else CKE((launch_split_p<N, false>(p, st)));
    };
    printf("=== N = %d\n", N);
    // ---------------- forward
    old_row(A, T1, PRO_NONE, 0);
    old_col(T1, REF, PRO_NONE, 1.0, 0);
    k_deint<cd><<<1184, 256, 0, st>>>(A, Ad, npts, N, Nh, 2, 1);
    new_rows(Ad, T1, 1.0, 0);
    new_colsub(T1, OUT[0], true);
    { const int pro[1] = {PRO_NONE}; new_p(OUT[0], OUT, pro, 1, true, 0); }
    printf("   forward: rel-L2 split vs cluster path %.3e\n", err(OUT[0], REF));
    CKE(cudaMemcpyAsync(S, REF, npts * sizeof(cd), cudaMemcpyDeviceToDevice, st));   // a spectrum for the inverse tests
    // ---------------- inverse with the three prologues
    const int pros[3] = {PRO_NONE, PRO_IK, PRO_IL};
    const double sc = 1.0 / ((double)N * N);
    { cd* outs[3] = {T1, OUT[1], OUT[2]}; new_p(S, outs, pros, 3, false, 1); }
    for (int o = 0; o < 3; ++o) {
        old_row(S, Ad, pros[o], 1);
        old_col(Ad, REF, PRO_NONE, sc, 1);
        cd* src = o == 0 ? T1 : OUT[o];
        new_colsub(src, Ad, false);
        new_rows(Ad, Ad, sc, 1);
        k_deint<cd><<<1184, 256, 0, st>>>(Ad, src, npts, N, Nh, 2, 0);      // back to natural x order
        printf("   inverse (prologue %d): rel-L2 split vs cluster path %.3e\n", pros[o], err(src, REF));
    }
    // round trip
    {
        const int pro[1] = {PRO_NONE};
        cd* outs[1] = {T1};
        new_p(S, outs, pro, 1, false, 1);
        new_colsub(T1, T2, false);
        new_rows(T2, T2, sc, 1);
        printf("   round trip ifft2(fft2(x)): rel-L2 %.3e\n", err(T2, Ad == nullptr ? A : (k_deint<cd><<<1184, 256, 0, st>>>(A, Ad, npts, N, Nh, 2, 1), Ad)));
    }
    // ---------------- timing
    printf("   timing, old path:\n");
    float o1 = timeit("row pass (cluster for N = 8192)", [&] { old_row(A, T1, PRO_NONE, 0); });
    float o2 = timeit("column pass (three-pass / cluster)", [&] { old_col(T1, REF, PRO_NONE, 1.0, 0); });
    printf("      2-D transform: %.4f ms\n", o1 + o2);
    printf("   timing, split path:\n");
    float n1 = timeit("rows: 2N one-tile lines of N/2", [&] { new_rows(Ad, T1, 1.0, 0); });
    row_bulk = 1;
    timeit("rows, tiles fetched by cp.async.bulk (UBLKCP)", [&] { new_rows(Ad, T1, 1.0, 0); });
    new_rows(Ad, T2, 1.0, 0);
    row_bulk = 0;
    new_rows(Ad, T1, 1.0, 0);
    printf("      bulk-copy rows vs register-load rows: rel-L2 %.1e\n", err(T2, T1));
    timeit("rows, bulk copy + no L2 prefetch", [&] { FftArgs a = base; a.in = Ad; a.out = T1; a.tw = tw_half; a.nlines = 2 * N; a.pitch = Nh;
        a.mstride = (size_t)Nh * Nh; a.g = Grid{Nh, dk, Nh, Nh / 2, 0, 0}; a.pf_groups = 0; a.tma_in = 2; CKE((launch_pass_g<Nh, false, true>(a, 1, st))); });
    timeit("rows, pf_groups = 0 (no L2 prefetch)", [&] { FftArgs a = base; a.in = Ad; a.out = T1; a.tw = tw_half; a.nlines = 2 * N; a.pitch = Nh;
        a.mstride = (size_t)Nh * Nh; a.g = Grid{Nh, dk, Nh, Nh / 2, 0, 0}; a.pf_groups = 0; CKE((launch_pass_g<Nh, false, true>(a, 1, st))); });
    float n2 = timeit("k_fft_colsub2<DIT>", [&] { new_colsub(T1, OUT[0], true); });
    float n3 = timeit("k_split_p<DIT> in place", [&] { const int pro[1] = {PRO_NONE}; new_p(OUT[0], OUT, pro, 1, true, 0); });
    printf("      forward 2-D transform: %.4f ms\n", n1 + n2 + n3);
    float i1 = timeit("k_split_p<DIF> 1 output", [&] { const int pro[1] = {PRO_NONE}; cd* outs[1] = {T1}; new_p(S, outs, pro, 1, false, 1); });
    float i3 = timeit("k_split_p<DIF> 3 outputs (phi, phix, phiy)", [&] { cd* outs[3] = {T1, OUT[1], OUT[2]}; new_p(S, outs, pros, 3, false, 1); });
    float i2 = timeit("k_fft_colsub2<DIF>", [&] { new_colsub(T1, Ad, false); });
    float i4 = timeit("rows with conj + scale", [&] { new_rows(Ad, Ad, sc, 1); });
    printf("      inverse 2-D transform: %.4f ms;  three inverse transforms of one spectrum: %.4f ms (old: %.4f)\n", i1 + i2 + i4,
           i3 + 3 * (i2 + i4), 3 * (o1 + o2));
    cudaFree(A); cudaFree(Ad); cudaFree(REF); cudaFree(T1); cudaFree(T2); cudaFree(S);
    for (int o = 0; o < 3; ++o) cudaFree(OUT[o]);
}

int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 0 || which == 2048) run<2048>();
    if (which == 0 || which == 4096) run<4096>();
    if (which == 0 || which == 8192) run<8192>();
    CKE(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
