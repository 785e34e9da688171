// CPU replay of the work-unit geometry of the fused spectral kernels (niwqg_b200/csrc/kernels_fused.cuh): every spectral
// element (ky, kx) must be owned by exactly one thread of one unit, and the thread a unit names as the holder of -K must
// hold exactly (-ky, -kx) at the register index fused_qp() says.  Build + run (no GPU needed):
//   nvcc -O1 -std=c++17 -I<nccl include> -o /tmp/host_fused_map tests/host/host_fused_map.cu && /tmp/host_fused_map
#include <cstdio>
#include <vector>
#include "../../niwqg_b200/csrc/kernels_fused.cuh"

template <int N> int check() {
    using G = FusedGeom<N>;
    constexpr int M = G::M;
    std::vector<unsigned char> seen((size_t)N * N, 0);
    long bad = 0;
    for (int unit = 0; unit < G::UNITS; ++unit) {
        FusedThread t[256];
        for (int tid = 0; tid < 256; ++tid) fused_thread<N>(unit, tid, nullptr, t[tid]);
        for (int tid = 0; tid < 256; ++tid) {
            const FusedThread& a = t[tid];
            if (a.ptid < 0 || a.ptid >= 256) { ++bad; continue; }
            const FusedThread& b = t[a.ptid];
            // lanes l and l + 16 of a warp must hold the columns n and n + N/2 of the same family (x butterfly by shuffle)
            const FusedThread& o = t[tid ^ 16];
            if (o.fam != a.fam || o.n != a.n || o.side == a.side || a.col != a.n + a.side * (N / 2)) ++bad;
            for (int q = 0; q < 16; ++q) {
                const int ky = a.fam + M * q, kx = a.col;
                if (seen[(size_t)ky * N + kx]++) ++bad;                          // owned twice
                const int qp = fused_qp(a, q);
                const int kyp = b.fam + M * qp, kxp = b.col;
                if (kyp != ((N - ky) & (N - 1)) || kxp != ((N - kx) & (N - 1))) ++bad;   // partner is not -K
            }
        }
    }
    long missing = 0;
    for (size_t i = 0; i < seen.size(); ++i) if (seen[i] != 1) ++missing;
    printf("N=%5d units=%6d: %ld mapping errors, %ld elements not owned exactly once\n", N, G::UNITS, bad, missing);
    return (bad || missing) ? 1 : 0;
}

int main() { return check<2048>() | check<4096>() | check<8192>(); }
