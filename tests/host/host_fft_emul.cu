// CPU replay of the shared-memory Stockham index maps in niwqg_b200/csrc/fft_core.cuh.
// Build+run (no GPU needed): nvcc -O1 -o /tmp/host_fft_emul tests/host/host_fft_emul.cu && /tmp/host_fft_emul
#include <cstdio>
#include <cmath>
#include <vector>
#include <complex>
#include "../../niwqg_b200/csrc/fft_core.cuh"
using namespace fftc;

template <int N> void build_tw(std::vector<cd>& tw) {
    tw.assign(tw_table_len(N) + 1, make_double2(0, 0));
    for (int NS = 16; NS < N; NS *= 16) {
        int R = (N / NS >= 16) ? 16 : N / NS;
        for (int kk = 0; kk < NS; ++kk) {
            long double a = -2.0L * 3.141592653589793238462643383279502884L * kk / ((long double)NS * R);
            tw[tw_offset(NS) + kk] = make_double2((double)cosl(a), (double)sinl(a));
        }
    }
}

template <int N, int NS> struct Run {
    static void go(std::vector<cd>& regs, std::vector<cd>& smem, std::vector<cd>& out, const cd* tw) {
        constexpr int R = StageRadix<N, NS>::R;
        constexpr int S = E / R;
        constexpr bool LAST = (NS * R == N);
        const int T = N / E;
        for (int j = 0; j < T; ++j) {
            cd* v = &regs[j * E];
            stage_compute<N, NS>(v, j, tw);
            for (int u = 0; u < S; ++u)
                for (int p = 0; p < R; ++p) {
                    int o = stage_out_index<N, NS>(j, u, p);
                    if (LAST) out[o] = v[u + p * S]; else smem[phys(o)] = v[u + p * S];
                }
        }
        if (!LAST) {
            for (int j = 0; j < T; ++j)
                for (int e = 0; e < E; ++e) regs[j * E + e] = smem[phys(j + e * (N / E))];
            Run<N, (LAST ? NS : NS * R)>::go(regs, smem, out, tw);
        }
    }
};

template <int N> double test() {
    std::vector<cd> x(N), regs(N), smem(phys_len(N) + 16), out(N), tw;
    build_tw<N>(tw);
    for (int i = 0; i < N; ++i) x[i] = make_double2(sin(0.37 * i * i + 1.0), cos(1.3 * i) + 0.01 * i);
    for (int j = 0; j < N / E; ++j) for (int e = 0; e < E; ++e) regs[j * E + e] = x[j + e * (N / E)];
    Run<N, 1>::go(regs, smem, out, tw.data());
    // naive DFT in long double for a subset of outputs
    double err = 0, nrm = 0;
    int step = N > 512 ? N / 97 : 1;
    for (int k = 0; k < N; k += step) {
        long double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            long double a = -2.0L * 3.141592653589793238462643383279502884L * ((long long)k * n % N) / N;
            long double c = cosl(a), s = sinl(a);
            sr += x[n].x * c - x[n].y * s; si += x[n].x * s + x[n].y * c;
        }
        err += (out[k].x - sr) * (out[k].x - sr) + (out[k].y - si) * (out[k].y - si);
        nrm += sr * sr + si * si;
    }
    return sqrt(err / nrm);
}

// Cluster decimation-in-time split of fft2d.cuh: C CTAs run local M-point transforms of x[C m + c], multiply by
// w_N^{c k}, and CTA c' finishes k in [c' M/C, (c'+1) M/C) with a radix-C butterfly over the C partial results.
template <int M, int C> double test_cluster() {
    constexpr int N = M * C;
    std::vector<cd> x(N), out(N), tw, twc(N);
    build_tw<M>(tw);
    for (int t = 0; t < N; ++t) {
        long double a = -2.0L * 3.141592653589793238462643383279502884L * t / N;
        twc[t] = make_double2((double)cosl(a), (double)sinl(a));
    }
    for (int i = 0; i < N; ++i) x[i] = make_double2(sin(0.37 * i * i + 1.0), cos(1.3 * i) + 0.01 * i);
    std::vector<std::vector<cd>> park(C, std::vector<cd>(phys_len(M) + 16));
    for (int c = 0; c < C; ++c) {
        std::vector<cd> regs(M), smem(phys_len(M) + 16), loc(M);
        for (int j = 0; j < M / E; ++j) for (int e = 0; e < E; ++e) regs[j * E + e] = x[C * (j + e * (M / E)) + c];
        Run<M, 1>::go(regs, smem, loc, tw.data());
        for (int k = 0; k < M; ++k) park[c][phys(k)] = cmul(loc[k], twc[c * k]);
    }
    for (int c = 0; c < C; ++c)
        for (int g = 0; g < M / C; ++g) {
            const int k = c * (M / C) + g;
            cd v[16];
            for (int r = 0; r < C; ++r) v[r] = park[r][phys(k)];
            dft<C, 1>(v);
            for (int p = 0; p < C; ++p) out[k + M * outidx<C>(p)] = v[p];
        }
    double err = 0, nrm = 0;
    int step = N > 512 ? N / 97 : 1;
    for (int k = 0; k < N; k += step) {
        long double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            long double a = -2.0L * 3.141592653589793238462643383279502884L * ((long long)k * n % N) / N;
            long double c = cosl(a), s = sinl(a);
            sr += x[n].x * c - x[n].y * s; si += x[n].x * s + x[n].y * c;
        }
        err += (out[k].x - sr) * (out[k].x - sr) + (out[k].y - si) * (out[k].y - si);
        nrm += sr * sr + si * si;
    }
    return sqrt(err / nrm);
}

// Cluster decimation-in-frequency split (k_fft_pass_dif): radix-C butterfly over x[m + M r], twiddle w_N^{m q},
// Y_q pushed to CTA q, local M-point transform, X[C k + q].
template <int M, int C> double test_cluster_dif() {
    constexpr int N = M * C;
    std::vector<cd> x(N), out(N), tw, twc(N);
    build_tw<M>(tw);
    for (int t = 0; t < N; ++t) {
        long double a = -2.0L * 3.141592653589793238462643383279502884L * t / N;
        twc[t] = make_double2((double)cosl(a), (double)sinl(a));
    }
    for (int i = 0; i < N; ++i) x[i] = make_double2(sin(0.37 * i * i + 1.0), cos(1.3 * i) + 0.01 * i);
    std::vector<std::vector<cd>> Y(C, std::vector<cd>(M));
    for (int m = 0; m < M; ++m) {
        cd v[16], pw[C > 1 ? C : 2];
        for (int r = 0; r < C; ++r) v[r] = x[m + M * r];
        dft<C, 1>(v);
        pw[0] = make_double2(1.0, 0.0); pw[1] = twc[m];
        for (int q = 2; q < C; ++q) pw[q] = cmul(pw[q >> 1], pw[q - (q >> 1)]);
        for (int p = 0; p < C; ++p) { int q = outidx<C>(p); Y[q][m] = q ? cmul(v[p], pw[q]) : v[p]; }
    }
    for (int q = 0; q < C; ++q) {
        std::vector<cd> regs(M), smem(phys_len(M) + 16), loc(M);
        for (int j = 0; j < M / E; ++j) for (int e = 0; e < E; ++e) regs[j * E + e] = Y[q][j + e * (M / E)];
        Run<M, 1>::go(regs, smem, loc, tw.data());
        for (int k = 0; k < M; ++k) out[C * k + q] = loc[k];
    }
    double err = 0, nrm = 0;
    int step = N > 512 ? N / 97 : 1;
    for (int k = 0; k < N; k += step) {
        long double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            long double a = -2.0L * 3.141592653589793238462643383279502884L * ((long long)k * n % N) / N;
            long double c = cosl(a), s = sinl(a);
            sr += x[n].x * c - x[n].y * s; si += x[n].x * s + x[n].y * c;
        }
        err += (out[k].x - sr) * (out[k].x - sr) + (out[k].y - si) * (out[k].y - si);
        nrm += sr * sr + si * si;
    }
    return sqrt(err / nrm);
}

int main() {
    int bad = 0;
#define TD(M, C) { double e = test_cluster_dif<M, C>(); printf("DIF M=%5d C=%d rel err %.3e\n", M, C, e); if (!(e < 1e-14)) bad = 1; }
    TD(1024, 2) TD(1024, 4) TD(1024, 8) TD(4096, 2)
#define TC(M, C) { double e = test_cluster<M, C>(); printf("M=%5d C=%d rel err %.3e\n", M, C, e); if (!(e < 1e-14)) bad = 1; }
    TC(1024, 2) TC(1024, 4) TC(1024, 8) TC(4096, 2)
#define T(N) { double e = test<N>(); printf("N=%5d rel err %.3e\n", N, e); if (!(e < 1e-14)) bad = 1; }
    T(32) T(64) T(128) T(256) T(512) T(1024) T(2048) T(4096) T(8192)
    return bad;
}
