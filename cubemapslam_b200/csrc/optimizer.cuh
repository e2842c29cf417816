// Shared host/device pieces of the optimizer translation units (ba.cu: LocalBundleAdjustment, pose_opt.cu: PoseOptimization).
#pragma once
#include <algorithm>
#include <vector>
#include "ba_math.cuh"
#include "common.cuh"

typedef void* nccl_comm_t;

struct cslam_optimizer {
    int device = 0;
    cudaStream_t stream = nullptr;
    nccl_comm_t comm = nullptr; int rank = 0, nranks = 1;
    int64_t launches = 0;
    // device arena: chunks are kept across calls (cudaMalloc/cudaFree per BA call cost more than the solve itself)
    struct Chunk { char* p; size_t size, used; };
    std::vector<Chunk> chunks;
    double* h_scal = nullptr;   // pinned, 16 doubles
    int clusterSize = 8;        // CTAs of the reduced-camera-system solver cluster
    bool useGraphs = true;      // single-GPU LM trials are replayed as one CUDA graph (CSLAM_BA_GRAPH=0 disables)
    // host scratch of cslam_local_ba, kept across calls (fresh multi-megabyte vectors cost page faults on every call)
    struct HostScratch { std::vector<int> lmStart, perm, eMP, eKF, fill, peStart, peList; std::vector<double> X; } hs;
    // optional per-kernel timing (cslam_optimizer_set_timing)
    bool timing = false; std::vector<cudaEvent_t> ev; std::vector<int> evKind; int evUsed = 0; double kindMs[16] = {0}; int64_t kindCount[16] = {0};
    // one-shot NVLink all-reduce of the reduced camera system (multi-GPU LocalBA): peer views of every rank's exchange buffer
    struct Peer { void* base = nullptr; bool mine = false; };
    std::vector<Peer> peers; size_t xchgBytes = 0; uint32_t epoch = 0, epoch2 = 0; bool oneShot = false;
};

namespace cslam {

template <class T>
static int dalloc(cslam_optimizer* o, T** p, size_t count, bool zero = false) {
    const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~(size_t)255;
    char* q = nullptr;
    for (auto& c : o->chunks) if (c.size - c.used >= bytes) { q = c.p + c.used; c.used += bytes; break; }
    if (!q) {
        cslam_optimizer::Chunk c; c.size = std::max<size_t>(bytes, (size_t)64 << 20); c.used = bytes;
        CSLAM_CUDA(cudaMalloc((void**)&c.p, c.size));
        o->chunks.push_back(c); q = c.p;
    }
    if (zero) CSLAM_CUDA(cudaMemsetAsync(q, 0, bytes, o->stream));
    *p = (T*)q;
    return 0;
}
template <class T>
static int dupload(cslam_optimizer* o, const T** p, const std::vector<T>& v) {
    T* q = nullptr; int rc = dalloc(o, &q, v.size());
    if (rc) return rc;
    if (!v.empty()) CSLAM_CUDA(cudaMemcpyAsync(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, o->stream));
    *p = q;
    return 0;
}
static inline void free_pool(cslam_optimizer* o) { for (auto& c : o->chunks) c.used = 0; }

#ifdef __CUDACC__
// fixed-shape (deterministic) block sum; result valid in thread 0. sh: >= 32 doubles
__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0;
    if (w == 0) {
        r = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;
}
#endif

}  // namespace cslam
