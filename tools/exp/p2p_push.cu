// Development aid: how fast can a kernel PUSH into peer memory over NVLink, as a function of the store pattern?
// One process, all visible GPUs, peer access enabled; every GPU pushes (P-1)/P of a buffer to its peers at the same time
// (the all-to-all of a slab transform).  Patterns: 16 B per thread in 512 B warp-contiguous runs (forward row pass today),
// 64 B and 128 B segments (inverse column pass with W = 4 / 8 columns), bulk copies shared -> peer global (cp.async.bulk),
// and cudaMemcpyPeerAsync by the copy engines for reference.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CKE(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
typedef double2 cd;
struct Peers { cd* p[8]; };

// every thread stores 16 B; a warp covers `seg` contiguous bytes per row segment, consecutive segments of a run of
// `run` bytes go to the same peer, then the next peer
__global__ void k_push_st(Peers pe, int P, int rank, size_t nelem, int segElems, int runElems) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nelem; i += stride) {
        // element i -> run index -> peer; within the buffer the address is scattered by segment to mimic strided rows
        const size_t run = i / runElems;
        const int peer = (int)((run + rank + 1) % P);
        if (peer == rank) continue;
        const size_t seg = i / segElems, off = i % segElems;
        const size_t nseg = nelem / segElems;
        const size_t dseg = (seg * 2654435761ull) % nseg;        // scatter the segments (different DRAM pages / lines)
        pe.p[peer][dseg * segElems + off] = make_double2((double)i, 1.0);
    }
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// bulk stores: every CTA fills an 8 KB (or chunk) staging area in shared memory and ships it with cp.async.bulk
__global__ void k_push_bulk(Peers pe, int P, int rank, size_t nelem, int chunkElems) {
    extern __shared__ __align__(128) unsigned char sm[];
    cd* buf = reinterpret_cast<cd*>(sm);
    const size_t nchunk = nelem / chunkElems;
    for (size_t c = blockIdx.x; c < nchunk; c += gridDim.x) {
        const int peer = (int)((c + rank + 1) % P);
        if (peer == rank) continue;
        for (int t = threadIdx.x; t < chunkElems; t += blockDim.x) buf[t] = make_double2((double)(c + t), 2.0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(pe.p[peer] + c * chunkElems), "r"(smem_u32(buf)), "r"((unsigned)(chunkElems * sizeof(cd))) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    int P = 0;
    CKE(cudaGetDeviceCount(&P));
    if (P < 2) { printf("needs >= 2 GPUs\n"); return 0; }
    if (P > 8) P = 8;
    const size_t nelem = (size_t)1 << 24;            // 256 MiB per GPU
    std::vector<cd*> buf(P);
    std::vector<cudaStream_t> st(P);
    std::vector<cudaEvent_t> e0(P), e1(P);
    for (int d = 0; d < P; ++d) {
        CKE(cudaSetDevice(d));
        for (int o = 0; o < P; ++o) if (o != d) { cudaError_t e = cudaDeviceEnablePeerAccess(o, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CKE(e); }
        CKE(cudaMalloc(&buf[d], nelem * sizeof(cd)));
        CKE(cudaMemset(buf[d], 0, nelem * sizeof(cd)));
        CKE(cudaStreamCreate(&st[d])); CKE(cudaEventCreate(&e0[d])); CKE(cudaEventCreate(&e1[d]));
        CKE(cudaFuncSetAttribute(k_push_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    }
    Peers pe{};
    for (int d = 0; d < P; ++d) pe.p[d] = buf[d];
    const double pushed = (double)nelem * sizeof(cd) * (P - 1) / P;      // bytes every GPU sends
    auto run = [&](const char* what, auto&& launch) {
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            for (int d = 0; d < P; ++d) { CKE(cudaSetDevice(d)); CKE(cudaDeviceSynchronize()); }
            for (int d = 0; d < P; ++d) { CKE(cudaSetDevice(d)); CKE(cudaEventRecord(e0[d], st[d])); launch(d); CKE(cudaEventRecord(e1[d], st[d])); }
            float worst = 0;
            for (int d = 0; d < P; ++d) { CKE(cudaSetDevice(d)); CKE(cudaEventSynchronize(e1[d])); float ms; CKE(cudaEventElapsedTime(&ms, e0[d], e1[d])); worst = std::max(worst, ms); }
            if (rep) best = std::min(best, worst);
        }
        printf("   %-62s %.3f ms  -> %.0f GB/s sent per GPU\n", what, best, pushed / (best * 1e-3) / 1e9);
        fflush(stdout);
    };
    printf("== %d GPUs, every GPU pushes %.0f MB to its %d peers at the same time\n", P, pushed / 1e6, P - 1);
    for (int ctas : {148, 296, 592}) {
        char w[128];
        snprintf(w, sizeof w, "16 B/thread, 512 B warp runs, 8 KB per peer run, %d CTAs", ctas);
        run(w, [&](int d) { k_push_st<<<ctas, 256, 0, st[d]>>>(pe, P, d, nelem, 512, 512); });
    }
    run("16 B/thread, 64 B segments (W = 4 columns), 296 CTAs", [&](int d) { k_push_st<<<296, 256, 0, st[d]>>>(pe, P, d, nelem, 4, 512); });
    run("16 B/thread, 128 B segments (W = 8 columns), 296 CTAs", [&](int d) { k_push_st<<<296, 256, 0, st[d]>>>(pe, P, d, nelem, 8, 512); });
    run("16 B/thread, 256 B segments, 296 CTAs", [&](int d) { k_push_st<<<296, 256, 0, st[d]>>>(pe, P, d, nelem, 16, 512); });
    for (int chunk : {128, 512, 2048}) {       // elements: 2 KB, 8 KB, 32 KB
        for (int ctas : {148, 296}) {
            char w[128];
            snprintf(w, sizeof w, "cp.async.bulk shared -> peer, %d KB chunks, %d CTAs", chunk * 16 / 1024, ctas);
            run(w, [&](int d) { k_push_bulk<<<ctas, 256, chunk * sizeof(cd), st[d]>>>(pe, P, d, nelem, chunk); });
        }
    }
    run("cudaMemcpyPeerAsync, one chunk per peer (copy engines)", [&](int d) {
        const size_t per = nelem / P;
        for (int k = 1; k < P; ++k) { const int o = (d + k) % P; CKE(cudaMemcpyPeerAsync(buf[o] + (size_t)d * per, o, buf[d] + (size_t)o * per, d, per * sizeof(cd), st[d])); }
    });
    for (int d = 0; d < P; ++d) { CKE(cudaSetDevice(d)); CKE(cudaDeviceSynchronize()); }
    printf("done\n");
    return 0;
}
