// K4: the two-term gradient combine on flat fp32 gradient buffers.
//   K4a siss_norm3   : { sum g_x^2, sum g_a^2, sum g_x g_a } in fp64 (8 B/param)
//   K4b siss_combine : out = clip * (g_x - s * g_a)            (12 B/param)
// Reference: delete_celeb.py:714-753 (+ delete_tshirt.py:688-690 inf guard) and the
// clip_grad_norm_(1.0) at delete_celeb.py:767. See include/siss_b200.h.
//
// Both are persistent grid-stride streaming kernels (grid = SMs x resident CTAs). K4b walks the
// buffers in the opposite direction to K4a so that the tail K4a left in the 126 MB L2 is the
// first thing K4b reads. The three scalars travel from K4a to K4b through device memory: no host
// synchronisation between them (an NCCL all-reduce of the three doubles can sit in between).

#include "common.cuh"
#include "combine_scalars.cuh"

namespace siss {

int env_int(const char* name, int dflt);

int cached_sm_count();

constexpr int kK4Occ = 4;     // resident CTAs per SM the kernels are compiled for
constexpr int kK4Unroll = 4;  // float4 per stream per thread per iteration
constexpr long long kK4Chunk = (long long)kThreads * kK4Unroll;  // float4 per CTA iteration

struct Norm3Workspace {
    unsigned int* counter;  // 1 ticket counter, zero between launches
    double* partials;       // [grid][3]
};

constexpr int kMaxNormGrid = 148 * 8;

inline Norm3Workspace carve_norm3(void* ws) {
    Norm3Workspace w;
    w.counter = reinterpret_cast<unsigned int*>(ws);
    w.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 256);
    return w;
}

__global__ void __launch_bounds__(kThreads, kK4Occ)
norm3_kernel(const float* __restrict__ gx, const float* __restrict__ ga, long long n, long long nvec,
             double* __restrict__ sums3, Norm3Workspace ws, long long keep_from_vec) {
    __shared__ double red[3 * kWarps];
    __shared__ int flag;
    double acc[3] = {0.0, 0.0, 0.0};
    // The combine that follows walks the buffers in REVERSE, so the tail [keep_from_vec, nvec) of both buffers is what
    // it reads first: load it with L2 evict_last priority and everything before it as read-once (evict_first), so the
    // tail is still L2-resident when the combine starts (instead of whatever the default replacement leaves).
    const unsigned long long pol_keep = l2_policy_evict_last(), pol_once = l2_policy_evict_first();

    const long long nchunks = (nvec + kK4Chunk - 1) / kK4Chunk;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const long long base = c * kK4Chunk + threadIdx.x;
        const unsigned long long pol = (c * kK4Chunk >= keep_from_vec) ? pol_keep : pol_once;   // uniform per chunk
        uint4 rx[kK4Unroll], ra[kK4Unroll];
        bool ok[kK4Unroll];
#pragma unroll
        for (int j = 0; j < kK4Unroll; ++j) {
            const long long i = base + (long long)j * kThreads;
            ok[j] = i < nvec;
            if (ok[j]) {
                rx[j] = ldg_stream_hint(gx + 4 * i, pol);
                ra[j] = ldg_stream_hint(ga + 4 * i, pol);
            }
        }
        // Products and sums in fp64 (exact products of fp32 values): the clip needs
        // ||g_x - s g_a||^2 = sxx - 2 s sxa + s^2 saa, which cancels when g_x ~ s g_a, so the three
        // sums must be good to fp64 rounding, not fp32. 3 DFMA + 2 cvt per parameter is ~15% of the
        // B200 fp64 pipe at HBM speed; the kernel stays memory-bound.
#pragma unroll
        for (int j = 0; j < kK4Unroll; ++j) {
            if (!ok[j]) continue;
            float x[4], a[4];
            VecTraits<float>::unpack(rx[j], x);
            VecTraits<float>::unpack(ra[j], a);
            // per-float4 partial sums first: the dependent DFMA chain on each accumulator is one DADD per
            // float4 instead of four DFMAs (the 16 elements of an iteration are otherwise one serial chain)
            double pxx = 0.0, paa = 0.0, pxa = 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double xd = (double)x[q], ad = (double)a[q];
                pxx = fma(xd, xd, pxx);
                paa = fma(ad, ad, paa);
                pxa = fma(xd, ad, pxa);
            }
            acc[0] += pxx; acc[1] += paa; acc[2] += pxa;
        }
    }
    // scalar remainder (n % 4 elements, or everything when the buffers are not 16B aligned)
    {
        const long long start = nvec * 4;
        const long long stride = (long long)gridDim.x * kThreads;
        for (long long i = start + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
            const double x = (double)gx[i], a = (double)ga[i];
            acc[0] = fma(x, x, acc[0]); acc[1] = fma(a, a, acc[1]); acc[2] = fma(x, a, acc[2]);
        }
    }

    block_sum<3>(acc, red);
    if (threadIdx.x == 0) {
        ws.partials[3 * blockIdx.x + 0] = acc[0];
        ws.partials[3 * blockIdx.x + 1] = acc[1];
        ws.partials[3 * blockIdx.x + 2] = acc[2];
    }
    if (last_cta_ticket(ws.counter, gridDim.x, &flag)) {
        // fixed-order final reduce by warp 0: lane l sums partials l, l+32, ... then butterfly
        if (threadIdx.x < 32) {
            double t[3] = {0.0, 0.0, 0.0};
            const volatile double* p = ws.partials;
            for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) {
                t[0] += p[3 * b + 0]; t[1] += p[3 * b + 1]; t[2] += p[3 * b + 2];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) t[k] = warp_sum(t[k]);
            if (threadIdx.x == 0) { sums3[0] = t[0]; sums3[1] = t[1]; sums3[2] = t[2]; }
        }
    }
}

// scalar prologue of K4b: combine_scalars.cuh (shared with the peer-memory variant in p2p.cu)
__device__ __forceinline__ CombineScalars combine_scalars(const double* __restrict__ sums3, int mode, float value,
                                                          float max_norm, int inf_guard, float* stats5,
                                                          bool write_stats) {
    return combine_scalars_from(sums3[0], sums3[1], sums3[2], mode, value, max_norm, inf_guard, stats5, write_stats);
}

__global__ void __launch_bounds__(kThreads, kK4Occ)
combine_kernel(const float* gx, const float* ga, float* out, long long n, long long nvec,
               const double* __restrict__ sums3, int mode, float value, float max_norm, int inf_guard,
               float* __restrict__ stats5) {
    const CombineScalars cs =
        combine_scalars(sums3, mode, value, max_norm, inf_guard, stats5, blockIdx.x == 0 && threadIdx.x == 0);
    const float s = cs.s, clip = cs.clip;

    const long long nchunks = (nvec + kK4Chunk - 1) / kK4Chunk;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const long long base = (nchunks - 1 - c) * kK4Chunk + threadIdx.x;  // reverse of K4a's order
        uint4 rx[kK4Unroll], ra[kK4Unroll];
        bool ok[kK4Unroll];
#pragma unroll
        for (int j = 0; j < kK4Unroll; ++j) {
            const long long i = base + (long long)j * kThreads;
            ok[j] = i < nvec;
            if (ok[j]) {
                rx[j] = ldg_v4(gx + 4 * i);   // coherent loads: `out` may alias g_x / g_a
                ra[j] = ldg_v4(ga + 4 * i);
            }
        }
#pragma unroll
        for (int j = 0; j < kK4Unroll; ++j) {
            if (!ok[j]) continue;
            const long long i = base + (long long)j * kThreads;
            float x[4], a[4], o[4];
            VecTraits<float>::unpack(rx[j], x);
            VecTraits<float>::unpack(ra[j], a);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                o[q] = __fmul_rn(__fsub_rn(x[q], __fmul_rn(s, a[q])), clip);  // mul, sub, then clip's mul_
            stg_stream(out + 4 * i, VecTraits<float>::pack(o));
        }
    }
    {
        const long long start = nvec * 4;
        const long long stride = (long long)gridDim.x * kThreads;
        for (long long i = start + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
            out[i] = __fmul_rn(__fsub_rn(gx[i], __fmul_rn(s, ga[i])), clip);
    }
}

static int k4_grid(long long nvec, long long n) {
    long long work = (nvec + kK4Chunk - 1) / kK4Chunk;
    if (nvec * 4 < n) {
        const long long tail_blocks = (n - nvec * 4 + kThreads - 1) / kThreads;
        if (tail_blocks > work) work = tail_blocks;
    }
    long long grid = (long long)cached_sm_count() * kK4Occ;
    if (grid > kMaxNormGrid) grid = kMaxNormGrid;
    if (work < grid) grid = work;
    if (grid < 1) grid = 1;
    return (int)grid;
}

}  // namespace siss

using namespace siss;

extern "C" {

int64_t siss_norm3_workspace_bytes(void) { return 256 + (int64_t)kMaxNormGrid * 3 * (int64_t)sizeof(double); }

int siss_norm3(const float* g_x, const float* g_a, int64_t n, double* sums3, void* workspace,
               siss_stream_t stream) {
    if (!g_x || !g_a || !sums3 || !workspace || n < 0) return SISS_EINVAL;
    const long long nvec = (aligned16(g_x) && aligned16(g_a)) ? n / 4 : 0;
    Norm3Workspace ws = carve_norm3(workspace);
    // bytes of (both) buffers' tails to keep in the 126 MB L2 for the combine; 0 disables the hints' effect
    static const long long keep_mb = env_int("SISS_L2_KEEP_MB", 80);
    long long keep_vec = (keep_mb << 20) / 2 / 16;
    if (g_a == g_x) keep_vec *= 2;                                   // single-term: one buffer, all of the budget
    const long long keep_from = (keep_mb <= 0) ? nvec : (nvec > keep_vec ? nvec - keep_vec : 0);
    norm3_kernel<<<k4_grid(nvec, n), kThreads, 0, (cudaStream_t)stream>>>(g_x, g_a, n, nvec, sums3, ws, keep_from);
    return (int)cudaGetLastError();
}

int siss_combine(const float* g_x, const float* g_a, float* out, int64_t n, const double* sums3,
                 int mode, float value, float max_norm, int inf_guard, float* stats5, siss_stream_t stream) {
    if (!g_x || !g_a || !out || !sums3 || n < 0) return SISS_EINVAL;
    if (mode < SISS_COMBINE_SCALING_NORM || mode > SISS_COMBINE_NONE) return SISS_EINVAL;
    const long long nvec = (aligned16(g_x) && aligned16(g_a) && aligned16(out)) ? n / 4 : 0;
    combine_kernel<<<k4_grid(nvec, n), kThreads, 0, (cudaStream_t)stream>>>(
        g_x, g_a, out, n, nvec, sums3, mode, value, max_norm, inf_guard, stats5);
    return (int)cudaGetLastError();
}

int siss_abi_version(void) { return SISS_B200_ABI_VERSION; }

const char* siss_error_string(int code) {
    switch (code) {
        case SISS_OK: return "ok";
        case SISS_EINVAL: return "siss: invalid argument (null pointer, negative size or bad enum)";
        case SISS_EUNSUPPORTED: return "siss: dtype combination not compiled into libsiss_b200";
        case SISS_EARCH: return "siss: device is not sm_100 (B200); this library has no other target";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "siss: unknown error";
    }
}

int siss_check_device(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0, major = 0, minor = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    return major == 10 ? SISS_OK : SISS_EARCH;
}

}  // extern "C"
