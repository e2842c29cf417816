// 256-bit Hamming matching of libcubemap_b200.so, sm_100a.
//
// Reference (CPU): ORBMatcher::DescriptorDistance src/ORBMatcher.cpp:951-967 (SWAR popcount of 8 x 32-bit XOR),
//   SearchByBoW(KeyFrame*,Frame&,...) :409-539, ComputeThreeMaxima :905-946, thresholds :42-45.
// The all-pairs ("brute force") matcher of BASELINE config 3 has no reference function; DESIGN.md defines it as the
// SearchByBoW acceptance rule + rotation histogram applied to the best/second-best column of every row of A.
//
// k_match_bruteforce  one CTA per pair; B is staged through shared memory in 128-descriptor tiles, every thread
//                     owns one row of A (8 x u32 in registers), __popc per 32-bit lane; per-pair rotation histogram and
//                     3-maxima filter in the same CTA. Bound by the integer/POPC issue rate, not by HBM (192 KB per pair).
// k_search_by_bow     one CTA per pair; both FeatureVectors are rebuilt on the fly by an in-CTA bitonic sort of
//                     (node, index); one warp walks the KF features of a node in order (the reference's "already
//                     matched" skip makes that order matter), lanes scan the F features of the node.
#include <cstring>
#include <vector>
#include "common.cuh"

namespace cslam {

static const int HISTO_BINS = 30;      // ceil(360 / HISTO_LENGTH), HISTO_LENGTH = 12
static const int BF_THREADS = 256;
static const int BF_TILE = 128;

__device__ __forceinline__ int rot_bin(float a, float b) {
    const float factor = 1.0f / 12.0f;
    float rot = __fsub_rn(a, b);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, factor));   // C round(): half away from zero
    if (bin == HISTO_BINS) bin = 0;
    return bin;
}

// ComputeThreeMaxima on bin counts (thread 0 only); keep[] = 1 for the bins that survive
__device__ void three_maxima(const int* hist, int* keep) {
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
    for (int i = 0; i < HISTO_BINS; i++) {
        const int s = hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
    for (int i = 0; i < HISTO_BINS; i++) keep[i] = (i == ind1 || i == ind2 || i == ind3);
}

__device__ __forceinline__ int hamming256(const uint32_t (&a)[8], const uint4 b0, const uint4 b1) {
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w) +
           __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}

// Layout: pair p reads rows [p*strideA, p*strideA + nA) of descA (and B likewise); angles are read with a stride of `angStride` floats
// (1 for a plain float array, 7 for cv::KeyPoint records, whose `angle` field the pointer then addresses); when nAarr / nBarr are given they
// hold per-pair row counts (the front end's n_out). Outputs are strided like A.
struct BfLayout { int strideA, strideB, angStride; const int32_t* nAarr; const int32_t* nBarr; };
__global__ void __launch_bounds__(BF_THREADS) k_match_bruteforce(const uint8_t* __restrict__ descA, const float* __restrict__ angA, int nA_,
                                                                 const uint8_t* __restrict__ descB, const float* __restrict__ angB, int nB_, float nnratio,
                                                                 int thLow, int checkOri, int32_t* __restrict__ match12, int32_t* __restrict__ dist12,
                                                                 int32_t* __restrict__ second12, int32_t* __restrict__ nmatches, BfLayout L, int* __restrict__ gHist) {
    // gridDim.y > 1: the rows of A are dealt to several CTAs per pair (a batch of ~128 pairs would otherwise leave 148 SMs one CTA each, or none);
    // the rotation histogram then goes through gHist[pair][32] (30 bins, total, arrival counter; zeroed by the launcher) and the CTA that arrives
    // last applies the three-maxima filter to the whole pair.
    __shared__ __align__(16) uint4 sB[2 * BF_TILE];
    __shared__ int hist[HISTO_BINS], keep[HISTO_BINS], total;
    __shared__ bool lastCta;
    const int pair = blockIdx.x, tid = threadIdx.x, part = blockIdx.y, parts = gridDim.y;
    const int nA = L.nAarr ? min(L.nAarr[pair], L.strideA) : nA_, nB = L.nBarr ? min(L.nBarr[pair], L.strideB) : nB_;
    const uint4* A4 = reinterpret_cast<const uint4*>(descA + (size_t)pair * L.strideA * 32);
    const uint4* B4 = reinterpret_cast<const uint4*>(descB + (size_t)pair * L.strideB * 32);
    const float* aA = angA + (size_t)pair * L.strideA * L.angStride;
    const float* aB = angB + (size_t)pair * L.strideB * L.angStride;
    int32_t* m12 = match12 + (size_t)pair * L.strideA;
    if (tid < HISTO_BINS) hist[tid] = 0;
    if (tid == 0) total = 0;
    __syncthreads();
    for (int rbase = part * BF_THREADS; rbase < nA; rbase += parts * BF_THREADS) {
        const int i = rbase + tid;
        uint32_t a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (i < nA) {
            const uint4 a0 = __ldg(A4 + 2 * i), a1 = __ldg(A4 + 2 * i + 1);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        }
        int best1 = 256, best2 = 256, bestIdx = -1;
        for (int cb = 0; cb < nB; cb += BF_TILE) {
            const int nt = min(BF_TILE, nB - cb);
            __syncthreads();
            for (int k = tid; k < 2 * nt; k += BF_THREADS) sB[k] = __ldg(B4 + 2 * cb + k);
            __syncthreads();
            if (i < nA) {
#pragma unroll 4
                for (int j = 0; j < nt; j++) {
                    const int d = hamming256(a, sB[2 * j], sB[2 * j + 1]);
                    if (d < best1) { best2 = best1; best1 = d; bestIdx = cb + j; }
                    else if (d < best2) best2 = d;
                }
            }
        }
        if (i < nA) {
            int m = -1;
            if (bestIdx >= 0 && best1 <= thLow && (float)best1 < __fmul_rn(nnratio, (float)best2)) {
                m = bestIdx;
                if (checkOri) atomicAdd(&hist[rot_bin(aA[(size_t)i * L.angStride], aB[(size_t)bestIdx * L.angStride])], 1);
                atomicAdd(&total, 1);
            }
            m12[i] = m;
            if (dist12) dist12[(size_t)pair * L.strideA + i] = best1;
            if (second12) second12[(size_t)pair * L.strideA + i] = best2;
        }
    }
    __syncthreads();
    if (parts > 1) {
        int* g = gHist + (size_t)pair * 32;
        if (tid < HISTO_BINS && hist[tid]) atomicAdd(g + tid, hist[tid]);
        if (tid == 0 && total) atomicAdd(g + 30, total);
        __threadfence();
        __syncthreads();
        if (tid == 0) lastCta = atomicAdd(g + 31, 1) == parts - 1;
        __syncthreads();
        if (!lastCta) return;
        __threadfence();
        if (tid < HISTO_BINS) hist[tid] = __ldcg(g + tid);
        if (tid == 0) total = __ldcg(g + 30);
        __syncthreads();
    }
    if (checkOri) {
        if (tid == 0) three_maxima(hist, keep);
        __syncthreads();
        for (int i = tid; i < nA; i += BF_THREADS) {
            const int m = parts > 1 ? __ldcg(m12 + i) : m12[i];
            if (m >= 0 && !keep[rot_bin(aA[(size_t)i * L.angStride], aB[(size_t)m * L.angStride])]) { m12[i] = -1; atomicSub(&total, 1); }
        }
        __syncthreads();
    }
    if (tid == 0) nmatches[pair] = total;
}

// ------------------------------------------------------------------------------------------------- SearchByBoW
static const int BOW_THREADS = 256;

__device__ void bitonic_sort(uint32_t* key, int n2) {   // ascending, n2 power of two
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n2; t += blockDim.x) {
                const int ixj = t ^ j;
                if (ixj > t) {
                    const uint32_t a = key[t], b = key[ixj];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) { key[t] = b; key[ixj] = a; }
                }
            }
            __syncthreads();
        }
}

// keys: node << 12 | index (index < 4096, node < 2^20). Shared memory: keyKF[n2K], keyF[n2F], takenF[nF] bytes.
__global__ void __launch_bounds__(BOW_THREADS) k_search_by_bow(const uint8_t* __restrict__ descKF, const float* __restrict__ angKF,
                                                               const uint8_t* __restrict__ kfValid, const int32_t* __restrict__ nodeKF, int nKF,
                                                               const uint8_t* __restrict__ descF, const float* __restrict__ angF,
                                                               const int32_t* __restrict__ nodeF, int nF, int n2K, int n2F, float nnratio, int checkOri,
                                                               int32_t* __restrict__ matchF, int32_t* __restrict__ nmatches,
                                                               const uint8_t* __restrict__ fValid, int kfkfMode) {
    // kfkfMode = 0: SearchByBoW(KeyFrame*, Frame&)  (src/ORBMatcher.cpp:409-539), output per F feature.
    // kfkfMode = 1: SearchByBoW(KeyFrame*, KeyFrame*) (:541-674): F-side features need fValid, acceptance is strict (< TH_LOW),
    //               output per KF feature (index of the matched feature of the second key frame).
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* keyKF = reinterpret_cast<uint32_t*>(smem);
    uint32_t* keyF = keyKF + n2K;
    int* mF = reinterpret_cast<int*>(keyF + n2F);   // match per F feature (KF index or -1), staged in shared memory
    __shared__ int hist[HISTO_BINS], keep[HISTO_BINS], total;
    const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = BOW_THREADS / 32;
    const uint4* K4 = reinterpret_cast<const uint4*>(descKF + (size_t)pair * nKF * 32);
    const uint4* F4 = reinterpret_cast<const uint4*>(descF + (size_t)pair * nF * 32);
    const float* aK = angKF + (size_t)pair * nKF;
    const float* aF = angF + (size_t)pair * nF;
    const uint8_t* val = kfValid + (size_t)pair * nKF;
    const int32_t* nK = nodeKF + (size_t)pair * nKF;
    const int32_t* nFd = nodeF + (size_t)pair * nF;
    for (int i = tid; i < n2K; i += BOW_THREADS) keyKF[i] = i < nKF ? ((uint32_t)nK[i] << 12) | (uint32_t)i : 0xffffffffu;
    for (int i = tid; i < n2F; i += BOW_THREADS) keyF[i] = i < nF ? ((uint32_t)nFd[i] << 12) | (uint32_t)i : 0xffffffffu;
    for (int i = tid; i < nF; i += BOW_THREADS) mF[i] = -1;
    if (tid < HISTO_BINS) hist[tid] = 0;
    if (tid == 0) total = 0;
    __syncthreads();
    bitonic_sort(keyKF, n2K);
    bitonic_sort(keyF, n2F);
    // one warp per KF node group: a group starts where the node id changes
    for (int s = warp; s < nKF; s += nwarps) {
        // warps take candidate group starts round-robin: position s is a start iff s==0 or node differs from s-1
        // (cheap: every warp scans its own positions; groups are independent of each other)
        const uint32_t node = keyKF[s] >> 12;
        if (s > 0 && (keyKF[s - 1] >> 12) == node) continue;
        // KF group [s, e)
        int e = s + 1;
        while (e < nKF && (keyKF[e] >> 12) == node) e++;
        // F group by binary search of node << 12
        int lo = 0, hi = nF;
        const uint32_t target = node << 12;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (keyF[mid] < target) lo = mid + 1; else hi = mid; }
        const int fs = lo;
        int fe = fs;
        while (fe < nF && (keyF[fe] >> 12) == node) fe++;
        if (fe == fs) continue;
        for (int q = s; q < e; q++) {
            const int iK = keyKF[q] & 0xfff;
            if (!val[iK]) continue;
            const uint4 k0 = __ldg(K4 + 2 * iK), k1 = __ldg(K4 + 2 * iK + 1);
            const uint32_t a[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
            // lanes scan the F features of the node in list order; (best1, pos, best2) with first-position ties
            int best1 = 256, best2 = 256, bpos = 0x7fffffff;
            for (int p = fs + lane; p < fe; p += 32) {
                const int iF = keyF[p] & 0xfff;
                if (mF[iF] >= 0) continue;
                if (kfkfMode && !fValid[(size_t)pair * nF + iF]) continue;
                const int d = hamming256(a, __ldg(F4 + 2 * iF), __ldg(F4 + 2 * iF + 1));
                if (d < best1) { best2 = best1; best1 = d; bpos = p; }
                else if (d < best2) best2 = d;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const int ob1 = __shfl_xor_sync(0xffffffffu, best1, o), ob2 = __shfl_xor_sync(0xffffffffu, best2, o), op = __shfl_xor_sync(0xffffffffu, bpos, o);
                // merge two partial scans: the winner is the smaller distance, ties go to the earlier list position
                const bool otherWins = ob1 < best1 || (ob1 == best1 && op < bpos);
                if (otherWins) { best2 = min(best1, ob2); best1 = ob1; bpos = op; }
                else best2 = min(best2, ob1);
            }
            if ((kfkfMode ? best1 < 50 : best1 <= 50) && (float)best1 < __fmul_rn(nnratio, (float)best2)) {
                const int iF = keyF[bpos] & 0xfff;
                if (lane == 0) {
                    mF[iF] = iK;
                    if (checkOri) atomicAdd(&hist[rot_bin(aK[iK], aF[iF])], 1);
                    atomicAdd(&total, 1);
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (checkOri) {
        if (tid == 0) three_maxima(hist, keep);
        __syncthreads();
        for (int j = tid; j < nF; j += BOW_THREADS) {
            const int m = mF[j];
            if (m >= 0 && !keep[rot_bin(aK[m], aF[j])]) { mF[j] = -1; atomicSub(&total, 1); }
        }
        __syncthreads();
    }
    if (!kfkfMode) {
        int32_t* out = matchF + (size_t)pair * nF;
        for (int j = tid; j < nF; j += BOW_THREADS) out[j] = mF[j];
    } else {
        int32_t* out = matchF + (size_t)pair * nKF;
        for (int i = tid; i < nKF; i += BOW_THREADS) out[i] = -1;
        __syncthreads();
        for (int j = tid; j < nF; j += BOW_THREADS) if (mF[j] >= 0) out[mF[j]] = j;
    }
    if (tid == 0) nmatches[pair] = total;
}

__global__ void k_hamming(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n, int32_t* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* A4 = reinterpret_cast<const uint4*>(a) + 2 * (size_t)i;
    const uint4* B4 = reinterpret_cast<const uint4*>(b) + 2 * (size_t)i;
    const uint4 a0 = __ldg(A4), a1 = __ldg(A4 + 1);
    const uint32_t av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    dist[i] = hamming256(av, __ldg(B4), __ldg(B4 + 1));
}

}  // namespace cslam

using namespace cslam;

struct cslam_matcher {
    int device = 0, maxPairs = 0, maxFeat = 0;
    cudaStream_t stream = nullptr;
    uint8_t *dA = nullptr, *dB = nullptr, *dValid = nullptr;
    float *aA = nullptr, *aB = nullptr;
    int32_t *nodeA = nullptr, *nodeB = nullptr, *dMatch = nullptr, *dDist = nullptr, *dSecond = nullptr, *dN = nullptr;
    cslam_keypoint* dKps = nullptr;   // staging of cslam_match_frames (lazily allocated)
    int* dHist = nullptr;             // per-pair rotation histogram of the row-split brute-force launches (lazily allocated)
    int64_t launches = 0;
};

extern "C" int cslam_matcher_create(cslam_matcher** out, int device, int max_pairs, int max_features) {
    if (!out || max_pairs <= 0 || max_features <= 0 || max_features > 4096) { set_error("cslam_matcher_create: bad argument (max_features <= 4096)"); return CSLAM_E_BADARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no CUDA device (this library has no CPU fallback)"); return CSLAM_E_NODEVICE; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return CSLAM_E_BADARG; }
    CSLAM_CUDA(cudaSetDevice(device));
    cslam_matcher* m = new cslam_matcher;
    m->device = device; m->maxPairs = max_pairs; m->maxFeat = max_features;
    const size_t nf = (size_t)max_pairs * max_features;
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess || cudaMalloc(&m->dA, nf * 32) != cudaSuccess || cudaMalloc(&m->dB, nf * 32) != cudaSuccess ||
        cudaMalloc(&m->dValid, nf) != cudaSuccess || cudaMalloc(&m->aA, nf * 4) != cudaSuccess || cudaMalloc(&m->aB, nf * 4) != cudaSuccess ||
        cudaMalloc(&m->nodeA, nf * 4) != cudaSuccess || cudaMalloc(&m->nodeB, nf * 4) != cudaSuccess || cudaMalloc(&m->dMatch, nf * 4) != cudaSuccess ||
        cudaMalloc(&m->dDist, nf * 4) != cudaSuccess || cudaMalloc(&m->dSecond, nf * 4) != cudaSuccess || cudaMalloc(&m->dN, (size_t)max_pairs * 4) != cudaSuccess) {
        set_error("matcher: device allocation failed");
        cslam_matcher_destroy(m);
        return CSLAM_E_CUDA;
    }
    cudaFuncSetAttribute(k_search_by_bow, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    *out = m;
    return CSLAM_OK;
}

extern "C" void cslam_matcher_destroy(cslam_matcher* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) { cudaStreamSynchronize(m->stream); cudaStreamDestroy(m->stream); }
    void* ptrs[] = {m->dA, m->dB, m->dValid, m->aA, m->aB, m->nodeA, m->nodeB, m->dMatch, m->dDist, m->dSecond, m->dN, m->dKps, m->dHist};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete m;
}
extern "C" void* cslam_matcher_stream(const cslam_matcher* m) { return m ? (void*)m->stream : nullptr; }
extern "C" int64_t cslam_matcher_launches(const cslam_matcher* m) { return m ? m->launches : 0; }
extern "C" int cslam_matcher_sync(cslam_matcher* m) {
    if (!m) return CSLAM_E_BADARG;
    CSLAM_CUDA(cudaSetDevice(m->device));
    CSLAM_CUDA(cudaStreamSynchronize(m->stream));
    return CSLAM_OK;
}

// CTAs per pair of a brute-force launch: enough CTAs for ~2 per SM when the batch of pairs is small; the per-pair histogram buffer is zeroed here
static int bf_split(cslam_matcher* m, int npairs, int** gHist) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
    int parts = 1;
    if (npairs < 2 * sms) parts = std::min(4, (2 * sms + npairs - 1) / npairs);
    *gHist = nullptr;
    if (parts > 1) {
        if (!m->dHist && cudaMalloc(&m->dHist, (size_t)2 * sms * 32 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return 1; }   // npairs < 2 * sms whenever parts > 1
        if (cudaMemsetAsync(m->dHist, 0, (size_t)npairs * 32 * sizeof(int), m->stream) != cudaSuccess) return 1;
        *gHist = m->dHist;
    }
    return parts;
}

static int check_sizes(cslam_matcher* m, int nA, int nB, int npairs) {
    if (!m || nA < 0 || nB < 0 || npairs <= 0 || npairs > m->maxPairs || nA > m->maxFeat || nB > m->maxFeat) {
        set_error("matcher: sizes out of range (pairs %d/%d, features %d,%d/%d)", npairs, m ? m->maxPairs : 0, nA, nB, m ? m->maxFeat : 0);
        return CSLAM_E_BADARG;
    }
    CSLAM_CUDA(cudaSetDevice(m->device));
    return 0;
}

extern "C" int cslam_match_bruteforce_dev(cslam_matcher* m, const uint8_t* descA, const float* angA, int nA, const uint8_t* descB, const float* angB, int nB,
                                          int npairs, float nnratio, int th_low, int check_ori, int32_t* match12, int32_t* dist12, int32_t* second12,
                                          int32_t* nmatches) {
    int rc = check_sizes(m, nA, nB, npairs);
    if (rc) return rc;
    if (nA == 0) { CSLAM_CUDA(cudaMemsetAsync(nmatches, 0, (size_t)npairs * 4, m->stream)); return CSLAM_OK; }
    const BfLayout L = {nA, nB, 1, nullptr, nullptr};
    int* gHist; const int parts = bf_split(m, npairs, &gHist);
    k_match_bruteforce<<<dim3(npairs, parts), BF_THREADS, 0, m->stream>>>(descA, angA, nA, descB, angB, nB, nnratio, th_low, check_ori, match12, dist12, second12, nmatches, L, gHist);
    m->launches++;
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

extern "C" int cslam_match_bruteforce(cslam_matcher* m, const uint8_t* descA, const float* angA, int nA, const uint8_t* descB, const float* angB, int nB, int npairs,
                                      float nnratio, int th_low, int check_ori, int32_t* match12, int32_t* dist12, int32_t* second12, int32_t* nmatches) {
    int rc = check_sizes(m, nA, nB, npairs);
    if (rc) return rc;
    const size_t na = (size_t)npairs * nA, nb = (size_t)npairs * nB;
    if (na) { CSLAM_CUDA(cudaMemcpyAsync(m->dA, descA, na * 32, cudaMemcpyHostToDevice, m->stream)); CSLAM_CUDA(cudaMemcpyAsync(m->aA, angA, na * 4, cudaMemcpyHostToDevice, m->stream)); }
    if (nb) { CSLAM_CUDA(cudaMemcpyAsync(m->dB, descB, nb * 32, cudaMemcpyHostToDevice, m->stream)); CSLAM_CUDA(cudaMemcpyAsync(m->aB, angB, nb * 4, cudaMemcpyHostToDevice, m->stream)); }
    rc = cslam_match_bruteforce_dev(m, m->dA, m->aA, nA, m->dB, m->aB, nB, npairs, nnratio, th_low, check_ori, m->dMatch, m->dDist, m->dSecond, m->dN);
    if (rc) return rc;
    if (na) {
        CSLAM_CUDA(cudaMemcpyAsync(match12, m->dMatch, na * 4, cudaMemcpyDeviceToHost, m->stream));
        if (dist12) CSLAM_CUDA(cudaMemcpyAsync(dist12, m->dDist, na * 4, cudaMemcpyDeviceToHost, m->stream));
        if (second12) CSLAM_CUDA(cudaMemcpyAsync(second12, m->dSecond, na * 4, cudaMemcpyDeviceToHost, m->stream));
    }
    CSLAM_CUDA(cudaMemcpyAsync(nmatches, m->dN, (size_t)npairs * 4, cudaMemcpyDeviceToHost, m->stream));
    CSLAM_CUDA(cudaStreamSynchronize(m->stream));
    return CSLAM_OK;
}

// All-pairs matcher between consecutive frames of a front-end batch, straight from cslam_frontend_run_dev's output layout: frame f (rows
// [f*kp_stride, f*kp_stride + n[f]) of kps / desc) is matched against frame f+1 for f = 0 .. nframes-2. match12: (nframes-1) x kp_stride.
extern "C" int cslam_match_frames_dev(cslam_matcher* m, const cslam_keypoint* kps, const uint8_t* desc, const int32_t* n, int kp_stride, int nframes, float nnratio, int th_low,
                                      int check_ori, int32_t* match12, int32_t* nmatches) {
    if (!m || !kps || !desc || !n || !match12 || !nmatches || nframes < 2 || kp_stride <= 0) { set_error("cslam_match_frames_dev: bad argument"); return CSLAM_E_BADARG; }
    CSLAM_CUDA(cudaSetDevice(m->device));
    const float* ang = &kps[0].angle;
    const BfLayout L = {kp_stride, kp_stride, (int)(sizeof(cslam_keypoint) / sizeof(float)), n, n + 1};
    int* gHist; const int parts = bf_split(m, nframes - 1, &gHist);
    k_match_bruteforce<<<dim3(nframes - 1, parts), BF_THREADS, 0, m->stream>>>(desc, ang, 0, desc + (size_t)kp_stride * 32, ang + (size_t)kp_stride * L.angStride, 0, nnratio, th_low,
                                                                                check_ori, match12, nullptr, nullptr, nmatches, L, gHist);
    m->launches++;
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

// host buffers in / out (copies inside, synchronous): kps / desc / n as cslam_frontend_run returns them
extern "C" int cslam_match_frames(cslam_matcher* m, const cslam_keypoint* kps, const uint8_t* desc, const int32_t* n, int kp_stride, int nframes, float nnratio, int th_low,
                                  int check_ori, int32_t* match12, int32_t* nmatches) {
    if (!m || nframes < 2 || nframes > m->maxPairs || kp_stride <= 0 || kp_stride > m->maxFeat) { set_error("cslam_match_frames: sizes out of range (frames %d/%d, stride %d/%d)", nframes, m ? m->maxPairs : 0, kp_stride, m ? m->maxFeat : 0); return CSLAM_E_BADARG; }
    CSLAM_CUDA(cudaSetDevice(m->device));
    const size_t nf = (size_t)nframes * kp_stride;
    if (!m->dKps) CSLAM_CUDA(cudaMalloc(&m->dKps, (size_t)m->maxPairs * m->maxFeat * sizeof(cslam_keypoint)));
    CSLAM_CUDA(cudaMemcpyAsync(m->dKps, kps, nf * sizeof(cslam_keypoint), cudaMemcpyHostToDevice, m->stream));
    CSLAM_CUDA(cudaMemcpyAsync(m->dA, desc, nf * 32, cudaMemcpyHostToDevice, m->stream));
    CSLAM_CUDA(cudaMemcpyAsync(m->nodeA, n, (size_t)nframes * 4, cudaMemcpyHostToDevice, m->stream));
    int rc = cslam_match_frames_dev(m, m->dKps, m->dA, m->nodeA, kp_stride, nframes, nnratio, th_low, check_ori, m->dMatch, m->dN);
    if (rc) return rc;
    CSLAM_CUDA(cudaMemcpyAsync(match12, m->dMatch, (size_t)(nframes - 1) * kp_stride * 4, cudaMemcpyDeviceToHost, m->stream));
    CSLAM_CUDA(cudaMemcpyAsync(nmatches, m->dN, (size_t)(nframes - 1) * 4, cudaMemcpyDeviceToHost, m->stream));
    CSLAM_CUDA(cudaStreamSynchronize(m->stream));
    return CSLAM_OK;
}

// POPC issue-rate micro-benchmark ("measure the real rate with a micro-benchmark on the box before quoting a fraction"): every thread
// runs `iters` rounds of 8 independent XOR+POPC+ADD chains on registers - the exact instruction mix of hamming256 without any memory traffic.
__global__ void __launch_bounds__(256) k_ubench_popc(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = seed * (threadIdx.x + 1) + k * 0x9e3779b9u + blockIdx.x; acc[k] = 0; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) { acc[k] += __popc(a[k] ^ acc[(k + 1) & 7]); }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k];
    if (s == 0xdeadbeefu) out[0] = s;   // keeps the chains alive
}
extern "C" int cslam_ubench_popc(cslam_matcher* m, double* popc32_per_s) {
    if (!m || !popc32_per_s) return CSLAM_E_BADARG;
    CSLAM_CUDA(cudaSetDevice(m->device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
    const int blocks = sms * 8, iters = 20000;
    cudaEvent_t e0, e1;
    CSLAM_CUDA(cudaEventCreate(&e0)); CSLAM_CUDA(cudaEventCreate(&e1));
    k_ubench_popc<<<blocks, 256, 0, m->stream>>>(reinterpret_cast<uint32_t*>(m->dN), 100, 1u);   // warm-up
    CSLAM_CUDA(cudaEventRecord(e0, m->stream));
    k_ubench_popc<<<blocks, 256, 0, m->stream>>>(reinterpret_cast<uint32_t*>(m->dN), iters, 1u);
    CSLAM_CUDA(cudaEventRecord(e1, m->stream));
    CSLAM_CUDA(cudaStreamSynchronize(m->stream));
    float ms = 0;
    CSLAM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    m->launches += 2;
    *popc32_per_s = (double)blocks * 256.0 * iters * 8.0 / (ms * 1e-3);
    return CSLAM_OK;
}

// The front end's k_fast is bound by the issue rate of 3-input u16x2 min / max (VIMNMX3.U16x2, 80 of them per pixel pair): the same kind of
// register-only chain measurement gives the denominator of its roofline.
__global__ void __launch_bounds__(256) k_ubench_minmax3(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = (seed * (threadIdx.x + 1) + k * 0x9e3779b9u + blockIdx.x) & 0x00ff00ffu; acc[k] = a[k] ^ 0x00550055u; }
    for (int i = 0; i < iters; i += 2) {
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = __vimin3_u16x2(acc[k] + 0x00010001u * (i & 1), a[k], acc[(k + 1) & 7]);
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = __vimax3_u16x2(acc[k], a[(k + 3) & 7], acc[(k + 5) & 7]);
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k];
    if (s == 0xdeadbeefu) out[0] = s;
}
extern "C" int cslam_ubench_minmax3(cslam_matcher* m, double* ops_per_s) {
    if (!m || !ops_per_s) return CSLAM_E_BADARG;
    CSLAM_CUDA(cudaSetDevice(m->device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
    const int blocks = sms * 8, iters = 20000;
    cudaEvent_t e0, e1;
    CSLAM_CUDA(cudaEventCreate(&e0)); CSLAM_CUDA(cudaEventCreate(&e1));
    k_ubench_minmax3<<<blocks, 256, 0, m->stream>>>(reinterpret_cast<uint32_t*>(m->dN), 100, 1u);   // warm-up
    CSLAM_CUDA(cudaEventRecord(e0, m->stream));
    k_ubench_minmax3<<<blocks, 256, 0, m->stream>>>(reinterpret_cast<uint32_t*>(m->dN), iters, 1u);
    CSLAM_CUDA(cudaEventRecord(e1, m->stream));
    CSLAM_CUDA(cudaStreamSynchronize(m->stream));
    float ms = 0;
    CSLAM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    m->launches += 2;
    *ops_per_s = (double)blocks * 256.0 * iters * 8.0 / (ms * 1e-3);   // warp-lane u16x2 3-input min/max per second
    return CSLAM_OK;
}

static int pow2_at_least(int n) { int p = 1; while (p < n) p <<= 1; return p; }

extern "C" int cslam_search_by_bow_dev(cslam_matcher* m, const uint8_t* descKF, const float* angKF, const uint8_t* kf_valid, const int32_t* node_kf, int nKF,
                                       const uint8_t* descF, const float* angF, const int32_t* node_f, int nF, int npairs, float nnratio, int check_ori,
                                       int32_t* match_f, int32_t* nmatches) {
    int rc = check_sizes(m, nKF, nF, npairs);
    if (rc) return rc;
    if (nF == 0 || nKF == 0) {
        CSLAM_CUDA(cudaMemsetAsync(nmatches, 0, (size_t)npairs * 4, m->stream));
        if (nF) CSLAM_CUDA(cudaMemsetAsync(match_f, 0xff, (size_t)npairs * nF * 4, m->stream));
        return CSLAM_OK;
    }
    const int n2K = pow2_at_least(nKF), n2F = pow2_at_least(nF);
    const size_t smem = (size_t)(n2K + n2F + nF) * 4;
    k_search_by_bow<<<npairs, BOW_THREADS, smem, m->stream>>>(descKF, angKF, kf_valid, node_kf, nKF, descF, angF, node_f, nF, n2K, n2F, nnratio, check_ori, match_f, nmatches, nullptr, 0);
    m->launches++;
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

extern "C" int cslam_search_by_bow(cslam_matcher* m, const uint8_t* descKF, const float* angKF, const uint8_t* kf_valid, const int32_t* node_kf, int nKF,
                                   const uint8_t* descF, const float* angF, const int32_t* node_f, int nF, int npairs, float nnratio, int check_ori, int32_t* match_f,
                                   int32_t* nmatches) {
    int rc = check_sizes(m, nKF, nF, npairs);
    if (rc) return rc;
    const size_t nk = (size_t)npairs * nKF, nf = (size_t)npairs * nF;
    if (nk) {
        CSLAM_CUDA(cudaMemcpyAsync(m->dA, descKF, nk * 32, cudaMemcpyHostToDevice, m->stream)); CSLAM_CUDA(cudaMemcpyAsync(m->aA, angKF, nk * 4, cudaMemcpyHostToDevice, m->stream));
        CSLAM_CUDA(cudaMemcpyAsync(m->dValid, kf_valid, nk, cudaMemcpyHostToDevice, m->stream)); CSLAM_CUDA(cudaMemcpyAsync(m->nodeA, node_kf, nk * 4, cudaMemcpyHostToDevice, m->stream));
    }
    if (nf) {
        CSLAM_CUDA(cudaMemcpyAsync(m->dB, descF, nf * 32, cudaMemcpyHostToDevice, m->stream)); CSLAM_CUDA(cudaMemcpyAsync(m->aB, angF, nf * 4, cudaMemcpyHostToDevice, m->stream));
        CSLAM_CUDA(cudaMemcpyAsync(m->nodeB, node_f, nf * 4, cudaMemcpyHostToDevice, m->stream));
    }
    rc = cslam_search_by_bow_dev(m, m->dA, m->aA, m->dValid, m->nodeA, nKF, m->dB, m->aB, m->nodeB, nF, npairs, nnratio, check_ori, m->dMatch, m->dN);
    if (rc) return rc;
    if (nf) CSLAM_CUDA(cudaMemcpyAsync(match_f, m->dMatch, nf * 4, cudaMemcpyDeviceToHost, m->stream));
    CSLAM_CUDA(cudaMemcpyAsync(nmatches, m->dN, (size_t)npairs * 4, cudaMemcpyDeviceToHost, m->stream));
    CSLAM_CUDA(cudaStreamSynchronize(m->stream));
    return CSLAM_OK;
}

extern "C" int cslam_hamming(cslam_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* dist) {
    if (!m || n < 0 || (size_t)n > (size_t)m->maxPairs * m->maxFeat) { set_error("cslam_hamming: n out of range"); return CSLAM_E_BADARG; }
    if (n == 0) return CSLAM_OK;
    CSLAM_CUDA(cudaSetDevice(m->device));
    CSLAM_CUDA(cudaMemcpyAsync(m->dA, a, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream));
    CSLAM_CUDA(cudaMemcpyAsync(m->dB, b, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream));
    k_hamming<<<cdiv(n, 256), 256, 0, m->stream>>>(m->dA, m->dB, n, m->dDist);
    m->launches++;
    CSLAM_CUDA(cudaMemcpyAsync(dist, m->dDist, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
    CSLAM_CUDA(cudaStreamSynchronize(m->stream));
    return CSLAM_OK;
}

// ORBMatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)  (reference src/ORBMatcher.cpp:541-674), host buffers.
extern "C" int cslam_search_by_bow_kf(cslam_matcher* m, const uint8_t* desc1, const float* ang1, const uint8_t* valid1, const int32_t* node1, int n1,
                                      const uint8_t* desc2, const float* ang2, const uint8_t* valid2, const int32_t* node2, int n2, int npairs, float nnratio,
                                      int check_ori, int32_t* match12, int32_t* nmatches) {
    int rc = check_sizes(m, n1, n2, npairs);
    if (rc) return rc;
    if (n1 == 0 || n2 == 0) { for (size_t i = 0; i < (size_t)npairs * n1; i++) match12[i] = -1; for (int p = 0; p < npairs; p++) nmatches[p] = 0; return CSLAM_OK; }
    const size_t k1 = (size_t)npairs * n1, k2 = (size_t)npairs * n2;
    CSLAM_CUDA(cudaMemcpyAsync(m->dA, desc1, k1 * 32, cudaMemcpyHostToDevice, m->stream)); CSLAM_CUDA(cudaMemcpyAsync(m->aA, ang1, k1 * 4, cudaMemcpyHostToDevice, m->stream));
    CSLAM_CUDA(cudaMemcpyAsync(m->dValid, valid1, k1, cudaMemcpyHostToDevice, m->stream)); CSLAM_CUDA(cudaMemcpyAsync(m->nodeA, node1, k1 * 4, cudaMemcpyHostToDevice, m->stream));
    CSLAM_CUDA(cudaMemcpyAsync(m->dB, desc2, k2 * 32, cudaMemcpyHostToDevice, m->stream)); CSLAM_CUDA(cudaMemcpyAsync(m->aB, ang2, k2 * 4, cudaMemcpyHostToDevice, m->stream));
    CSLAM_CUDA(cudaMemcpyAsync(m->nodeB, node2, k2 * 4, cudaMemcpyHostToDevice, m->stream));
    uint8_t* dValid2 = reinterpret_cast<uint8_t*>(m->dSecond);   // scratch: second12 is unused by this entry point
    CSLAM_CUDA(cudaMemcpyAsync(dValid2, valid2, k2, cudaMemcpyHostToDevice, m->stream));
    const int n2K = pow2_at_least(n1), n2F = pow2_at_least(n2);
    const size_t smem = (size_t)(n2K + n2F + n2) * 4;
    k_search_by_bow<<<npairs, BOW_THREADS, smem, m->stream>>>(m->dA, m->aA, m->dValid, m->nodeA, n1, m->dB, m->aB, m->nodeB, n2, n2K, n2F, nnratio, check_ori, m->dMatch, m->dN,
                                                              dValid2, 1);
    m->launches++;
    CSLAM_CUDA(cudaGetLastError());
    CSLAM_CUDA(cudaMemcpyAsync(match12, m->dMatch, k1 * 4, cudaMemcpyDeviceToHost, m->stream));
    CSLAM_CUDA(cudaMemcpyAsync(nmatches, m->dN, (size_t)npairs * 4, cudaMemcpyDeviceToHost, m->stream));
    CSLAM_CUDA(cudaStreamSynchronize(m->stream));
    return CSLAM_OK;
}
