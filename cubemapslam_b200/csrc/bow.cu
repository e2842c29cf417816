// DBoW2 vocabulary-tree transform of libcubemap_b200.so, sm_100a.
//
// Reference (CPU): Frame::ComputeBoW / KeyFrame::ComputeBoW (src/Frame.cpp:719-726) -> TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, 4)
//   ThirdParty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1193 (per feature :1218-1262: greedy descent, first minimum wins; FORB::distance FORB.cpp:81-101),
//   BowVector::addWeight / normalize(L1) BowVector.cpp:31-45,62-85. ORBvoc.txt is k = 10, L = 6, L1_NORM scoring, TF_IDF weighting.
//
// k_bow_descend   thread per feature: L levels x k children, 256-bit Hamming against node descriptors (35 MB for the full vocabulary: L2-resident);
//                 output word id, leaf node, FeatureVector node (level L - levelsup)
// k_bow_vector    CTA per frame: the BowVector std::map restated as sort by (word, feature index) + in-order segment sums (double, the order
//                 addWeight sees) + the L1 normalisation summed in word order like BowVector::normalize -> bit-identical doubles
#include <cstring>
#include <vector>
#include "common.cuh"

namespace cslam {

struct VocDev { int k, L, nNodes; const int* childStart; const int* childList; const uint4* desc; const double* weight; const int* wordId; };

__global__ void __launch_bounds__(128) k_bow_descend(VocDev V, const uint8_t* __restrict__ feats, const int32_t* __restrict__ n, int stride, int levelsup,
                                                      int32_t* __restrict__ word, int32_t* __restrict__ leaf, int32_t* __restrict__ node) {
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(n[f], stride)) return;
    const uint4* F4 = reinterpret_cast<const uint4*>(feats + ((size_t)f * stride + i) * 32);
    const uint4 a0 = __ldg(F4), a1 = __ldg(F4 + 1);
    const int nidLevel = V.L - levelsup;
    int fin = 0, level = 0, nid = 0;
    do {
        ++level;
        const int s = V.childStart[fin], e = V.childStart[fin + 1];
        int best = 0x7fffffff, bestId = V.childList[s];
        for (int c = s; c < e; c++) {
            const int id = V.childList[c];
            const uint4 b0 = __ldg(V.desc + 2 * (size_t)id), b1 = __ldg(V.desc + 2 * (size_t)id + 1);
            const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) +
                          __popc(a1.w ^ b1.w);
            if (d < best) { best = d; bestId = id; }
        }
        fin = bestId;
        if (level == nidLevel) nid = fin;
    } while (V.childStart[fin + 1] > V.childStart[fin]);
    const size_t o = (size_t)f * stride + i;
    word[o] = V.wordId[fin]; leaf[o] = fin; node[o] = V.weight[fin] > 0 ? nid : -1;   // stopped words (weight 0) join neither vector
}

__device__ void bitonic_sort_u32(uint32_t* key, int n2) {
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n2; t += blockDim.x) {
                const int ixj = t ^ j;
                if (ixj > t) { const uint32_t a = key[t], b = key[ixj]; const bool up = (t & k) == 0; if ((a > b) == up) { key[t] = b; key[ixj] = a; } }
            }
            __syncthreads();
        }
}

// keys: word << 12 | feature index (index < 4096, word < 2^20)
__global__ void __launch_bounds__(256) k_bow_vector(VocDev V, const int32_t* __restrict__ n, int stride, int n2, const int32_t* __restrict__ word, const int32_t* __restrict__ leaf,
                                                    const int32_t* __restrict__ node, int32_t* __restrict__ bowWord, double* __restrict__ bowVal, int32_t* __restrict__ bowCount) {
    extern __shared__ uint32_t key[];           // n2 keys, then n2 flags / ranks
    uint32_t* rank = key + n2;
    __shared__ uint32_t wsum[8], carry;
    __shared__ double norm;
    const int f = blockIdx.x, tid = threadIdx.x, nf = min(n[f], stride);
    const size_t o = (size_t)f * stride;
    for (int i = tid; i < n2; i += 256) key[i] = (i < nf && node[o + i] >= 0) ? ((uint32_t)word[o + i] << 12) | (uint32_t)i : 0xffffffffu;
    __syncthreads();
    bitonic_sort_u32(key, n2);
    // segment starts -> output rank (exclusive scan of the start flags)
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n2; base += 256) {
        const int p = base + tid;
        const uint32_t x = (p < n2 && key[p] != 0xffffffffu && (p == 0 || (key[p - 1] >> 12) != (key[p] >> 12))) ? 1u : 0u;
        uint32_t s = x;
#pragma unroll
        for (int sh = 1; sh < 32; sh <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, s, sh); if ((tid & 31) >= sh) s += t; }
        if ((tid & 31) == 31) wsum[tid >> 5] = s;
        __syncthreads();
        uint32_t pre = carry;
        for (int w = 0; w < (tid >> 5); w++) pre += wsum[w];
        if (p < n2) rank[p] = x ? pre + s - 1 : 0xffffffffu;
        __syncthreads();
        if (tid == 255) carry = pre + s;
        __syncthreads();
    }
    const int nWords = (int)carry;
    // in-order segment sums: v[word] += weight, feature by feature (BowVector::addWeight)
    for (int p = tid; p < n2; p += 256) {
        if (rank[p] == 0xffffffffu) continue;
        const uint32_t w = key[p] >> 12;
        double s = 0;
        for (int q = p; q < n2 && key[q] != 0xffffffffu && (key[q] >> 12) == w; q++) s += V.weight[leaf[o + (key[q] & 0xfff)]];
        bowWord[o + rank[p]] = (int32_t)w; bowVal[o + rank[p]] = s;
    }
    __syncthreads();
    if (tid == 0) {   // BowVector::normalize(L1): norm += fabs(value) in word order
        double s = 0;
        for (int i = 0; i < nWords; i++) s += fabs(bowVal[o + i]);
        norm = s; bowCount[f] = nWords;
    }
    __syncthreads();
    if (norm > 0.0) for (int i = tid; i < nWords; i += 256) bowVal[o + i] = bowVal[o + i] / norm;
}

}  // namespace cslam

using namespace cslam;

struct cslam_vocabulary {
    int device = 0; VocDev V; cudaStream_t stream = nullptr; std::vector<void*> owned; int64_t launches = 0;
    int maxFrames = 0, maxFeat = 0;
    uint8_t* dFeat = nullptr; int32_t *dN = nullptr, *dWord = nullptr, *dLeaf = nullptr, *dNode = nullptr, *dBowWord = nullptr, *dBowCount = nullptr; double* dBowVal = nullptr;
};

template <class T> static int valloc(cslam_vocabulary* v, T** p, size_t n) { void* q = nullptr; CSLAM_CUDA(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T))); v->owned.push_back(q); *p = (T*)q; return 0; }

extern "C" void cslam_vocabulary_destroy(cslam_vocabulary* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    if (v->stream) { cudaStreamSynchronize(v->stream); cudaStreamDestroy(v->stream); }
    for (void* p : v->owned) cudaFree(p);
    delete v;
}

// nodes 1..n_nodes in ORBvoc.txt's order (TemplatedVocabulary::loadFromTextFile :1337-1415): parent id, leaf flag, 32 descriptor bytes, weight
extern "C" int cslam_vocabulary_create(cslam_vocabulary** out, int device, int k, int L, int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc,
                                       const double* weight, int max_frames, int max_features) {
    if (!out || k < 2 || L < 1 || n_nodes <= 0 || !parent || !is_leaf || !desc || !weight || max_frames <= 0 || max_features <= 0 || max_features > 4096) { set_error("cslam_vocabulary_create: bad argument"); return CSLAM_E_BADARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no CUDA device (this library has no CPU fallback)"); return CSLAM_E_NODEVICE; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return CSLAM_E_BADARG; }
    CSLAM_CUDA(cudaSetDevice(device));
    const int N = n_nodes + 1;
    std::vector<int> cnt(N + 1, 0), wordId(N, -1);
    int words = 0;
    for (int i = 0; i < n_nodes; i++) {
        if (parent[i] < 0 || parent[i] > i) { set_error("vocabulary node %d: parent %d must precede it", i + 1, parent[i]); return CSLAM_E_BADARG; }
        cnt[parent[i] + 1]++;
        if (is_leaf[i]) wordId[i + 1] = words++;
    }
    if (words >= (1 << 20)) { set_error("vocabulary has %d words; the packed sort keys hold word ids < 2^20", words); return CSLAM_E_BADARG; }
    for (int i = 0; i < N; i++) cnt[i + 1] += cnt[i];
    std::vector<int> start(cnt), fill(cnt.begin(), cnt.end() - 1), list(n_nodes);
    for (int i = 0; i < n_nodes; i++) list[fill[parent[i]]++] = i + 1;       // children in file order (push_back order)
    std::vector<uint8_t> d((size_t)N * 32, 0); std::vector<double> w(N, 0.0);
    std::memcpy(d.data() + 32, desc, (size_t)n_nodes * 32);
    std::memcpy(w.data() + 1, weight, (size_t)n_nodes * 8);
    cslam_vocabulary* v = new cslam_vocabulary; v->device = device; v->maxFrames = max_frames; v->maxFeat = max_features;
    int *dStart, *dList, *dWid; uint4* dDesc; double* dW; int rc;
    auto fail = [&](int code) { cslam_vocabulary_destroy(v); return code; };
    if (cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); return fail(CSLAM_E_CUDA); }
    const size_t nf = (size_t)max_frames * max_features;
    if ((rc = valloc(v, &dStart, N + 1)) || (rc = valloc(v, &dList, n_nodes)) || (rc = valloc(v, &dWid, N)) || (rc = valloc(v, &dDesc, (size_t)N * 2)) || (rc = valloc(v, &dW, N)) ||
        (rc = valloc(v, &v->dFeat, nf * 32)) || (rc = valloc(v, &v->dN, max_frames)) || (rc = valloc(v, &v->dWord, nf)) || (rc = valloc(v, &v->dLeaf, nf)) || (rc = valloc(v, &v->dNode, nf)) ||
        (rc = valloc(v, &v->dBowWord, nf)) || (rc = valloc(v, &v->dBowVal, nf)) || (rc = valloc(v, &v->dBowCount, max_frames))) return fail(rc);
    if (cudaMemcpy(dStart, start.data(), (N + 1) * 4, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(dList, list.data(), (size_t)n_nodes * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(dWid, wordId.data(), N * 4, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(dDesc, d.data(), (size_t)N * 32, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(dW, w.data(), (size_t)N * 8, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("vocabulary upload failed"); return fail(CSLAM_E_CUDA); }
    v->V.k = k; v->V.L = L; v->V.nNodes = N; v->V.childStart = dStart; v->V.childList = dList; v->V.desc = dDesc; v->V.weight = dW; v->V.wordId = dWid;
    cudaFuncSetAttribute(k_bow_vector, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * 4);
    *out = v;
    return CSLAM_OK;
}
extern "C" void* cslam_vocabulary_stream(const cslam_vocabulary* v) { return v ? (void*)v->stream : nullptr; }
extern "C" int cslam_vocabulary_sync(cslam_vocabulary* v) { if (!v) return CSLAM_E_BADARG; CSLAM_CUDA(cudaSetDevice(v->device)); CSLAM_CUDA(cudaStreamSynchronize(v->stream)); return CSLAM_OK; }

static int pow2_ge(int n) { int p = 1; while (p < n) p <<= 1; return p; }

extern "C" int cslam_bow_transform_dev(cslam_vocabulary* v, const uint8_t* desc, const int32_t* n, int stride, int nframes, int levelsup, int32_t* word, int32_t* leaf, int32_t* node,
                                       int32_t* bow_word, double* bow_val, int32_t* bow_count) {
    if (!v || !desc || !n || !word || !leaf || !node || !bow_word || !bow_val || !bow_count || nframes <= 0 || stride <= 0 || stride > 4096) { set_error("cslam_bow_transform_dev: bad argument"); return CSLAM_E_BADARG; }
    CSLAM_CUDA(cudaSetDevice(v->device));
    k_bow_descend<<<dim3(cdiv(stride, 128), nframes), 128, 0, v->stream>>>(v->V, desc, n, stride, levelsup, word, leaf, node);
    const int n2 = pow2_ge(stride);
    k_bow_vector<<<nframes, 256, (size_t)2 * n2 * 4, v->stream>>>(v->V, n, stride, n2, word, leaf, node, bow_word, bow_val, bow_count);
    v->launches += 2;
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

extern "C" int cslam_bow_transform(cslam_vocabulary* v, const uint8_t* desc, const int32_t* n, int stride, int nframes, int levelsup, int32_t* word, int32_t* node, int32_t* bow_word,
                                   double* bow_val, int32_t* bow_count) {
    if (!v || nframes <= 0 || nframes > v->maxFrames || stride <= 0 || stride > v->maxFeat) { set_error("cslam_bow_transform: sizes out of range"); return CSLAM_E_BADARG; }
    CSLAM_CUDA(cudaSetDevice(v->device));
    const size_t nf = (size_t)nframes * stride;
    CSLAM_CUDA(cudaMemcpyAsync(v->dFeat, desc, nf * 32, cudaMemcpyHostToDevice, v->stream));
    CSLAM_CUDA(cudaMemcpyAsync(v->dN, n, (size_t)nframes * 4, cudaMemcpyHostToDevice, v->stream));
    int rc = cslam_bow_transform_dev(v, v->dFeat, v->dN, stride, nframes, levelsup, v->dWord, v->dLeaf, v->dNode, v->dBowWord, v->dBowVal, v->dBowCount);
    if (rc) return rc;
    if (word) CSLAM_CUDA(cudaMemcpyAsync(word, v->dWord, nf * 4, cudaMemcpyDeviceToHost, v->stream));
    CSLAM_CUDA(cudaMemcpyAsync(node, v->dNode, nf * 4, cudaMemcpyDeviceToHost, v->stream));
    CSLAM_CUDA(cudaMemcpyAsync(bow_word, v->dBowWord, nf * 4, cudaMemcpyDeviceToHost, v->stream));
    CSLAM_CUDA(cudaMemcpyAsync(bow_val, v->dBowVal, nf * 8, cudaMemcpyDeviceToHost, v->stream));
    CSLAM_CUDA(cudaMemcpyAsync(bow_count, v->dBowCount, (size_t)nframes * 4, cudaMemcpyDeviceToHost, v->stream));
    CSLAM_CUDA(cudaStreamSynchronize(v->stream));
    return CSLAM_OK;
}
