// fft_core.cuh -- fp64 complex radix-16 Stockham building blocks (host+device).
//
// Replaces the reference's FFT backend seam (niwqg/Kernel.py:553-566 binds
// numpy.fft.fft2/ifft2; niwqg/QGModel.py:551-552 binds rfft2/irfft2).  Each
// thread keeps E=16 complex points of one length-N transform in registers;
// a stage is: twiddle -> radix-R butterfly in registers -> scatter through
// shared memory (Stockham autosort index map) -> contiguous gather.  The first
// stage reads straight from global memory and the last writes straight back,
// so a length-N pass costs log16(N)-1 shared-memory exchanges and exactly one
// HBM read + one HBM write per point.
//
// Everything here is __host__ __device__ so tests/host_fft_emul.cu can replay
// the index maps on the CPU (there is no GPU in the build container).
#pragma once
#include <cuda_runtime.h>

#define HD __host__ __device__ __forceinline__

typedef double2 cd;

HD cd cmul(cd a, cd b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
HD cd cadd(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
HD cd csub(cd a, cd b) { return make_double2(a.x - b.x, a.y - b.y); }
HD cd cconj(cd a) { return make_double2(a.x, -a.y); }
HD cd cscale(cd a, double s) { return make_double2(a.x * s, a.y * s); }
HD cd mul_mi(cd a) { return make_double2(a.y, -a.x); }   // a * (-i)
HD cd mul_pi(cd a) { return make_double2(-a.y, a.x); }   // a * (+i)

namespace fftc {

constexpr int E = 16;                 // points per thread
constexpr double C8 = 0.70710678118654752440;   // cos(pi/4)
constexpr double C16 = 0.92387953251128675613;  // cos(pi/8)
constexpr double S16 = 0.38268343236508977173;  // sin(pi/8)

// ---- leaf DFTs, forward sign (w = exp(-2 pi i / R)), in place, natural order, stride S
template <int S> HD void dft2(cd* v) {
    cd a = v[0], b = v[S];
    v[0] = cadd(a, b);
    v[S] = csub(a, b);
}
template <int S> HD void dft4(cd* v) {
    cd a0 = cadd(v[0], v[2 * S]), a1 = csub(v[0], v[2 * S]);
    cd a2 = cadd(v[S], v[3 * S]), a3 = csub(v[S], v[3 * S]);
    v[0] = cadd(a0, a2);
    v[2 * S] = csub(a0, a2);
    cd t = mul_mi(a3);
    v[S] = cadd(a1, t);
    v[3 * S] = csub(a1, t);
}

// position p of an in-place radix-R result holds frequency index outidx<R>(p)
template <int R> HD constexpr int outidx(int p) {
    return R == 8 ? ((p >> 2) + 2 * (p & 3)) : R == 16 ? ((p >> 2) + 4 * (p & 3)) : p;
}

// radix-8 = 2 x 4 (t = 4a+b): DFT2 over a, twiddle w8^{b k1}, DFT4 over b
template <int S> HD void dft8(cd* v) {
#pragma unroll
    for (int b = 0; b < 4; ++b) dft2<4 * S>(v + b * S);
    // k1 = 1 row: positions 4+b
    {
        cd x = v[5 * S];
        v[5 * S] = make_double2(C8 * (x.x + x.y), C8 * (x.y - x.x));       // * (C8 - i C8)
        v[6 * S] = mul_mi(v[6 * S]);
        x = v[7 * S];
        v[7 * S] = make_double2(C8 * (x.y - x.x), -C8 * (x.x + x.y));      // * (-C8 - i C8)
    }
    dft4<S>(v);
    dft4<S>(v + 4 * S);
}

// radix-16 = 4 x 4 (t = 4a+b): DFT4 over a (stride 4S), twiddle w16^{b k1}, DFT4 over b
template <int S> HD void dft16(cd* v) {
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4<4 * S>(v + b * S);
    const cd w1 = make_double2(C16, -S16), w2 = make_double2(C8, -C8), w3 = make_double2(S16, -C16);
    // k1=1: w^b ; k1=2: w^{2b} ; k1=3: w^{3b}
    v[5 * S] = cmul(v[5 * S], w1);
    {
        cd x = v[6 * S];
        v[6 * S] = make_double2(C8 * (x.x + x.y), C8 * (x.y - x.x));
    }
    v[7 * S] = cmul(v[7 * S], w3);
    {
        cd x = v[9 * S];
        v[9 * S] = make_double2(C8 * (x.x + x.y), C8 * (x.y - x.x));
    }
    v[10 * S] = mul_mi(v[10 * S]);
    {
        cd x = v[11 * S];
        v[11 * S] = make_double2(C8 * (x.y - x.x), -C8 * (x.x + x.y));      // w^6 = -C8 - i C8
    }
    v[13 * S] = cmul(v[13 * S], w3);
    {
        cd x = v[14 * S];
        v[14 * S] = make_double2(C8 * (x.y - x.x), -C8 * (x.x + x.y));      // w^6
    }
    v[15 * S] = cmul(v[15 * S], make_double2(-C16, S16));                  // w^9 = -w^1
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4<S>(v + 4 * k1 * S);
    (void)w2;
}

template <int R, int S> HD void dft(cd* v) {
    if (R == 2) dft2<S>(v);
    else if (R == 4) dft4<S>(v);
    else if (R == 8) dft8<S>(v);
    else dft16<S>(v);
}

// multiply v[t*S] by w^t, t = 1..R-1 (log-depth power chain)
template <int R, int S> HD void apply_twiddles(cd* v, cd w1) {
    if (R == 2) { v[S] = cmul(v[S], w1); return; }
    cd w2 = cmul(w1, w1), w3 = cmul(w2, w1);
    v[S] = cmul(v[S], w1);
    v[2 * S] = cmul(v[2 * S], w2);
    v[3 * S] = cmul(v[3 * S], w3);
    if (R == 4) return;
    cd w4 = cmul(w2, w2);
    v[4 * S] = cmul(v[4 * S], w4);
    v[5 * S] = cmul(v[5 * S], cmul(w4, w1));
    v[6 * S] = cmul(v[6 * S], cmul(w4, w2));
    v[7 * S] = cmul(v[7 * S], cmul(w4, w3));
    if (R == 8) return;
    cd w8 = cmul(w4, w4);
    v[8 * S] = cmul(v[8 * S], w8);
    v[9 * S] = cmul(v[9 * S], cmul(w8, w1));
    v[10 * S] = cmul(v[10 * S], cmul(w8, w2));
    v[11 * S] = cmul(v[11 * S], cmul(w8, w3));
    cd w12 = cmul(w8, w4);
    v[12 * S] = cmul(v[12 * S], w12);
    v[13 * S] = cmul(v[13 * S], cmul(w12, w1));
    v[14 * S] = cmul(v[14 * S], cmul(w12, w2));
    v[15 * S] = cmul(v[15 * S], cmul(w12, w3));
}

// radix of the stage whose previous-stage product is NS
template <int N, int NS> struct StageRadix { static constexpr int R = (N / NS >= 16) ? 16 : (N / NS); };

// offset of stage NS's twiddle block inside the per-N table: blocks for NS=16,256,4096 hold NS entries each
HD constexpr int tw_offset(int NS) { return NS == 16 ? 0 : NS == 256 ? 16 : NS == 4096 ? 272 : 0; }
HD constexpr int tw_table_len(int N) { return N > 4096 ? 4368 : N > 256 ? 272 : N > 16 ? 16 : 0; }

// padded shared-memory slot of logical element o (one pad slot per 16: conflict-free scatter)
HD constexpr int phys(int o) { return o + (o >> 4); }
HD constexpr int phys_len(int N) { return N + (N >> 4); }

// One stage's register work for thread j (of N/16) : twiddle + butterflies.
// Register slot u + p*S (S = 16/R) then holds frequency p' = outidx<R>(p) of sub-butterfly u.
template <int N, int NS> HD void stage_compute(cd* v, int j, const cd* __restrict__ tw) {
    constexpr int R = StageRadix<N, NS>::R;
    constexpr int S = E / R;
#pragma unroll
    for (int u = 0; u < S; ++u) {
        if (NS > 1) {
            int b = j + u * (N / E);
            int kk = b & (NS - 1);
            cd w1 = tw[tw_offset(NS) + kk];
            apply_twiddles<R, S>(v + u, w1);
        }
        dft<R, S>(v + u);
    }
}

// logical output index (along the transform axis) of register slot (u, p) after stage NS
template <int N, int NS> HD int stage_out_index(int j, int u, int p) {
    constexpr int R = StageRadix<N, NS>::R;
    int b = j + u * (N / E);
    int kk = b & (NS - 1);
    return (b - kk) * R + kk + outidx<R>(p) * NS;
}

}  // namespace fftc
