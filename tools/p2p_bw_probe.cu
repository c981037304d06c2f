// Standalone NVLink probe (one process, two GPUs, cudaDeviceEnablePeerAccess): how fast can SM-issued traffic
// move data to / from a peer, by access flavour?   nvcc -arch=sm_100a -O3 -o /tmp/p2p_bw_probe tools/p2p_bw_probe.cu
//   push_stg128   : peer stores, 16 B per thread                  pull_ldg128 : peer loads, 16 B per thread
//   push_tma      : smem tile -> cp.async.bulk.global.shared::cta to the peer (TILE bytes per bulk store)
//   pull_tma      : cp.async.bulk.shared.global from the peer into an smem ring (TILE bytes per bulk load)
// Each test runs on both GPUs at once ("bidir") or on GPU 0 only ("unidir").
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void push_stg128(const uint4* __restrict__ src, uint4* __restrict__ dst, long long nvec) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = (long long)blockIdx.x * blockDim.x * 4 + threadIdx.x; i < nvec; i += stride) {
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) if (i + j * blockDim.x < nvec) v[j] = src[i + j * blockDim.x];
#pragma unroll
        for (int j = 0; j < 4; ++j) if (i + j * blockDim.x < nvec) dst[i + j * blockDim.x] = v[j];
    }
}
// pull: same kernel with src = peer, dst = local

template <int TILE>
__global__ void push_tma(const uint4* __restrict__ src, char* __restrict__ dst, long long bytes) {
    extern __shared__ __align__(128) unsigned char sm[];   // 2 tiles
    const long long ntiles = bytes / TILE;
    int buf = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        unsigned char* tile = sm + buf * TILE;
        // make sure the bulk store that last read this buffer has finished reading
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        const uint4* s = src + t * (TILE / 16);
        for (int i = threadIdx.x; i < TILE / 16; i += blockDim.x) reinterpret_cast<uint4*>(tile)[i] = s[i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst + t * TILE), "r"(s32(tile)), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        buf ^= 1;
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int TILE, int STAGES>
__global__ void pull_tma(const char* __restrict__ src, uint4* __restrict__ dst, long long bytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + STAGES * TILE);
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(full + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long ntiles = bytes / TILE;
    long long my = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) ++my;
    // prologue: issue up to STAGES loads
    long long issued = 0, t_issue = blockIdx.x;
    if (threadIdx.x == 0)
        for (; issued < STAGES && issued < my; ++issued, t_issue += gridDim.x) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(full + issued)), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(s32(sm + issued * TILE)), "l"(src + t_issue * TILE), "r"(TILE), "r"(s32(full + issued)) : "memory");
        }
    long long k = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++k) {
        const int st = (int)(k % STAGES);
        const uint32_t ph = (uint32_t)((k / STAGES) & 1);
        uint32_t done;
        do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(s32(full + st)), "r"(ph) : "memory"); } while (!done);
        const uint4* tile = reinterpret_cast<const uint4*>(sm + st * TILE);
        uint4* d = dst + t * (TILE / 16);
        for (int i = threadIdx.x; i < TILE / 16; i += blockDim.x) d[i] = tile[i];
        __syncthreads();
        if (threadIdx.x == 0 && k + STAGES < my) {
            const long long tn = t + (long long)STAGES * gridDim.x;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(full + st)), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(s32(sm + st * TILE)), "l"(src + tn * TILE), "r"(TILE), "r"(s32(full + st)) : "memory");
        }
    }
}

int main() {
    int ndev = 0; CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("need 2 GPUs\n"); return 0; }
    const long long bytes = 512LL << 20;
    char* buf[2][2];
    cudaStream_t st[2]; cudaEvent_t e0[2], e1[2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d)); CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&buf[d][0], bytes)); CK(cudaMalloc(&buf[d][1], bytes));
        CK(cudaMemset(buf[d][0], d + 1, bytes)); CK(cudaStreamCreate(&st[d])); CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d]));
    }
    const int grid = 148 * 4, thr = 256;
    constexpr int TILE = 16384;
    auto run = [&](const char* name, int mode, bool bidir) -> int {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            for (int d = 0; d < (bidir ? 2 : 1); ++d) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], st[d])); 
                char* local = buf[d][0]; char* peer = buf[1 - d][1]; char* peer_src = buf[1 - d][0]; char* local_dst = buf[d][1];
                if (mode == 0) push_stg128<<<grid, thr, 0, st[d]>>>((const uint4*)local, (uint4*)peer, bytes / 16);
                if (mode == 1) push_stg128<<<grid, thr, 0, st[d]>>>((const uint4*)peer_src, (uint4*)local_dst, bytes / 16);
                if (mode == 2) { cudaFuncSetAttribute(push_tma<TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE);
                                 push_tma<TILE><<<grid, thr, 2 * TILE, st[d]>>>((const uint4*)local, peer, bytes); }
                if (mode == 3) { cudaFuncSetAttribute(pull_tma<TILE, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * TILE + 64);
                                 pull_tma<TILE, 4><<<148 * 3, thr, 4 * TILE + 64, st[d]>>>(peer_src, (uint4*)local_dst, bytes); }
                if (mode == 4) CK(cudaMemcpyPeerAsync(peer, 1 - d, local, d, bytes, st[d]));
                CK(cudaEventRecord(e1[d], st[d])); }
            float worst = 0;
            for (int d = 0; d < (bidir ? 2 : 1); ++d) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); float ms; CK(cudaEventElapsedTime(&ms, e0[d], e1[d])); if (ms > worst) worst = ms; }
            CK(cudaGetLastError());
            if (worst < best) best = worst;
        }
        printf("%-14s %-7s %8.3f ms  %7.1f GB/s per direction\n", name, bidir ? "bidir" : "unidir", best, bytes / (best * 1e-3) / 1e9);
        return 0;
    };
    const char* names[5] = {"push_stg128", "pull_ldg128", "push_tma16k", "pull_tma16k", "memcpyPeer"};
    for (int m = 0; m < 5; ++m) { if (run(names[m], m, false)) return 1; if (run(names[m], m, true)) return 1; }
    // verify one TMA result
    CK(cudaSetDevice(0)); std::vector<char> h(64); CK(cudaMemcpy(h.data(), buf[0][1], 64, cudaMemcpyDeviceToHost));
    printf("check byte on GPU0 dst buffer: %d (expect 2)\n", (int)h[0]);
    return 0;
}
