// kernels_ic.cuh -- initial conditions generated ON the device (niwqg/InitialConditions.py): at 8192^2 the reference
// generators need several 1 GiB host arrays, an O(N^2) Python loop (LambDipole, :102-107) and four host<->device
// round trips of the FFT seam (McWilliams1984, :36-41); here a seeded model costs a few kernel launches.
// Arrays are written in the DEVICE layout of the physical fields (x de-interleaved mod deintC, rows of this rank).
#pragma once
#include "common.cuh"

struct IcGeom {
    int N;          // global grid edge
    int nyl;        // local rows
    int row0;       // first global row of this rank
    int deintC, deintM;
    double L;
};
// natural x index of position p of a device row
__device__ __forceinline__ int ic_x(const IcGeom& g, int p) { return g.deintC > 1 ? g.deintC * (p % g.deintM) + p / g.deintM : p; }

// Lamb dipole vorticity (InitialConditions.py:77-114): q = C J1(lam r) sin(theta) inside r <= R, 0 outside,
// centred on the grid point (N/2, N/2); x = (i + 0.5) / N * L as in Kernel.py:232-233
__global__ void k_ic_lamb(IcGeom g, double U, double R, double* __restrict__ q) {
    const size_t total = (size_t)g.nyl * g.N;
    const double x0 = ((double)(g.N / 2) + 0.5) / g.N * g.L, y0 = x0;
    const double lam = 3.8317 / R;
    const double C = -(2. * U * lam) / j0(lam * R);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / g.N), p = (int)(i % g.N);
        const double x = ((double)ic_x(g, p) + 0.5) / g.N * g.L, y = ((double)(g.row0 + row) + 0.5) / g.N * g.L;
        const double r = sqrt((x - x0) * (x - x0) + (y - y0) * (y - y0));
        const double s = (r != 0.0) ? (y - y0) / r : 0.0;
        q[(size_t)blockIdx.y * total + i] = (r <= R) ? C * j1(lam * r) * s : 0.0;
    }
}

// phi generators: 0 wave packet exp(i(k(x-x0) + l(y-y0))) exp(-(r/R)^2) (:117-145); 1 plane wave exp(i(kx+ly) + phase)
// - the phase sits outside 1j* in the reference and so scales the amplitude (:167); 2 uniform re + i im
__global__ void k_ic_phi(IcGeom g, int kind, double k, double l, double R, double x0, double y0, double phase, cd* __restrict__ phi) {
    const size_t total = (size_t)g.nyl * g.N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / g.N), p = (int)(i % g.N);
        const double x = ((double)ic_x(g, p) + 0.5) / g.N * g.L, y = ((double)(g.row0 + row) + 0.5) / g.N * g.L;
        cd v;
        if (kind == 0) {
            const double r = sqrt((x - x0) * (x - x0) + (y - y0) * (y - y0));
            double s, c;
            sincos(k * (x - x0) + l * (y - y0), &s, &c);
            const double a = exp(-((r / R) * (r / R)));
            v = make_double2(c * a, s * a);
        } else if (kind == 1) {
            double s, c;
            sincos(k * x + l * y, &s, &c);
            const double a = exp(phase);
            v = make_double2(a * c, a * s);
        } else {
            v = make_double2(k, l);
        }
        phi[(size_t)blockIdx.y * total + i] = v;
    }
}

// Philox4x32-10 counter-based generator (Salmon et al. 2011): uniform double in [0, 1) for counter i
__device__ __forceinline__ double ic_philox_uniform(unsigned long long i, unsigned long long seed) {
    unsigned c0 = (unsigned)i, c1 = (unsigned)(i >> 32), c2 = 0u, c3 = 0u;
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return ((double)(c0 >> 5) * 67108864.0 + (double)(c1 >> 6)) * (1.0 / 9007199254740992.0);
}

// random red spectrum, step 1 (InitialConditions.py:29-35 / :63-69): ph = ckappa (cos(phase) + i sin(phase)) on the natural
// c2c spectral grid; phase = 2 pi * rand, either given (device array, [N][N]) or drawn from Philox
__global__ void k_ic_spectrum(int N, double dk, int kind, double k0, const double* __restrict__ rand01, unsigned long long seed,
                              cd* __restrict__ ph) {
    const size_t total = (size_t)N * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ky = (int)(i / N), kx = (int)(i % N);
        const double k = dk * (double)sidx(kx, N), l = dk * (double)sidx(ky, N);
        const double wv2 = k * k + l * l, wv = sqrt(wv2);
        double ck = 0.0;
        if (wv != 0.0) {
            if (kind == 0) { const double t = wv2 / (k0 * k0); ck = 1.0 / sqrt(wv2 * (1. + t * t)); }     // McWilliams (1984)
            else ck = sqrt(wv * exp(-(wv2 / (k0 * k0))));                                                 // Danioux et al. (2015)
        }
        const double u = rand01 ? rand01[i] : ic_philox_uniform(i, seed);
        double s, c;
        sincos(u * 2 * 3.14159265358979323846, &s, &c);
        ph[i] = make_double2(ck * c, ck * s);
    }
}
// drop the imaginary part of a physical field in place (".real" of the reference)
__global__ void k_ic_real(cd* __restrict__ a, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i].y = 0.0;
}
// sum over the spectrum (mode (0,0) removed) of wv2 |ph|^2:  Eaux = 0.5 * that / M^2  (InitialConditions.py:38, spec_var)
__global__ void k_ic_energy(int N, double dk, const cd* __restrict__ ph, double* partials) {
    const size_t total = (size_t)N * N;
    double s[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        if (i == 0) continue;
        const int ky = (int)(i / N), kx = (int)(i % N);
        const double k = dk * (double)sidx(kx, N), l = dk * (double)sidx(ky, N);
        const cd p = ph[i];
        s[0] += (k * k + l * l) * (p.x * p.x + p.y * p.y);
    }
    block_reduce_store<1>(s, partials);
}
// qh-like spectrum -wv2 * sqrt(E / Eaux) * ph (InitialConditions.py:39-41), Eaux from the device sum
__global__ void k_ic_scale(int N, double dk, double E, double M2, const double* __restrict__ sum, cd* __restrict__ ph) {
    const size_t total = (size_t)N * N;
    const double f = sqrt(E / (0.5 * sum[0] / M2));
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ky = (int)(i / N), kx = (int)(i % N);
        const double k = dk * (double)sidx(kx, N), l = dk * (double)sidx(ky, N);
        const double w = -(k * k + l * l) * f;
        ph[i] = make_double2(w * ph[i].x, w * ph[i].y);
    }
}
