// ORACLE (test infrastructure, not product code).
// CPU restatement of the Hamming-matching part of ORBMatcher:
//   DescriptorDistance      src/ORBMatcher.cpp:951-967
//   SearchByBoW(KF,F)       src/ORBMatcher.cpp:409-539
//   ComputeThreeMaxima      src/ORBMatcher.cpp:905-946
//   constants               src/ORBMatcher.cpp:42-45
// plus the all-pairs "brute force" matcher BASELINE.json config 3 names. The reference has no all-pairs
// function; it is DEFINED as SearchByBoW's per-row best/second-best + acceptance rule +
// rotation histogram over ALL columns, without the order-dependent "already matched" skip, reported per row
// of A (like vnMatches12 of SearchForInitialization, src/ORBMatcher.cpp:676-794).
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

namespace orc {

static const int TH_HIGH = 100;
static const int TH_LOW = 50;
static const int HISTO_LENGTH = 12;

static inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    const uint32_t* pa = (const uint32_t*)a;
    const uint32_t* pb = (const uint32_t*)b;
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

static inline void compute_three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

static inline int rot_bin(float angA, float angB, int nBins) {
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = angA - angB;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * factor);
    if (bin == nBins) bin = 0;
    return bin;
}

// SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&).
//   kfValid[i] != 0  <=>  KF feature i has a MapPoint that is not bad.
//   nodeKF/nodeF: vocabulary node id of each feature (DBoW2 FeatureVector, levelsup=4); the FeatureVector
//   is the map node -> ascending feature indices (ThirdParty/DBoW2/DBoW2/FeatureVector.cpp:31-45).
//   matchF[j] = index of the KF feature whose MapPoint was assigned to F feature j, or -1.
static inline int search_by_bow(const uint8_t* descKF, const float* angKF, const uint8_t* kfValid, const int* nodeKF, int nKF,
                                const uint8_t* descF, const float* angF, const int* nodeF, int nF,
                                float nnratio, bool checkOri, int* matchF) {
    std::map<int, std::vector<unsigned>> fvKF, fvF;
    for (int i = 0; i < nKF; i++) fvKF[nodeKF[i]].push_back(i);
    for (int i = 0; i < nF; i++) fvF[nodeF[i]].push_back(i);
    for (int j = 0; j < nF; j++) matchF[j] = -1;
    int nmatches = 0;
    const int nBinsAngle = (int)std::ceil(360.0f / HISTO_LENGTH);
    std::vector<std::vector<int>> rotHist(nBinsAngle);
    auto KFit = fvKF.begin(), KFend = fvKF.end();
    auto Fit = fvF.begin(), Fend = fvF.end();
    while (KFit != KFend && Fit != Fend) {
        if (KFit->first == Fit->first) {
            const std::vector<unsigned>& vIndicesKF = KFit->second;
            const std::vector<unsigned>& vIndicesF = Fit->second;
            for (size_t iKF = 0; iKF < vIndicesKF.size(); iKF++) {
                const unsigned realIdxKF = vIndicesKF[iKF];
                if (!kfValid[realIdxKF]) continue;
                const uint8_t* dKF = descKF + 32 * (size_t)realIdxKF;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (size_t iF = 0; iF < vIndicesF.size(); iF++) {
                    const unsigned realIdxF = vIndicesF[iF];
                    if (matchF[realIdxF] >= 0) continue;
                    const int dist = descriptor_distance(dKF, descF + 32 * (size_t)realIdxF);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= TH_LOW) {
                    if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                        matchF[bestIdxF] = (int)realIdxKF;
                        if (checkOri) rotHist[rot_bin(angKF[realIdxKF], angF[bestIdxF], nBinsAngle)].push_back(bestIdxF);
                        nmatches++;
                    }
                }
            }
            KFit++; Fit++;
        } else if (KFit->first < Fit->first) KFit = fvKF.lower_bound(Fit->first);
        else Fit = fvF.lower_bound(KFit->first);
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotHist.data(), nBinsAngle, ind1, ind2, ind3);
        for (int i = 0; i < nBinsAngle; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { matchF[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)  (src/ORBMatcher.cpp:541-674): both sides need a good MapPoint, features of
// KF2 are consumed (vbMatched2), acceptance is the STRICT bestDist1 < TH_LOW. match12[i] = index of the KF2 feature or -1.
static inline int search_by_bow_kf(const uint8_t* desc1, const float* ang1, const uint8_t* valid1, const int* node1, int n1, const uint8_t* desc2,
                                   const float* ang2, const uint8_t* valid2, const int* node2, int n2, float nnratio, bool checkOri, int* match12) {
    std::map<int, std::vector<unsigned>> fv1, fv2;
    for (int i = 0; i < n1; i++) fv1[node1[i]].push_back(i);
    for (int i = 0; i < n2; i++) fv2[node2[i]].push_back(i);
    for (int i = 0; i < n1; i++) match12[i] = -1;
    std::vector<bool> vbMatched2(n2, false);
    int nmatches = 0;
    const int nBinsAngle = (int)std::ceil(360.0f / HISTO_LENGTH);
    std::vector<std::vector<int>> rotHist(nBinsAngle);
    auto f1it = fv1.begin(), f1end = fv1.end(); auto f2it = fv2.begin(), f2end = fv2.end();
    while (f1it != f1end && f2it != f2end) {
        if (f1it->first == f2it->first) {
            for (size_t i1 = 0; i1 < f1it->second.size(); i1++) {
                const size_t idx1 = f1it->second[i1];
                if (!valid1[idx1]) continue;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (size_t i2 = 0; i2 < f2it->second.size(); i2++) {
                    const size_t idx2 = f2it->second[i2];
                    if (vbMatched2[idx2] || !valid2[idx2]) continue;
                    const int dist = descriptor_distance(desc1 + 32 * idx1, desc2 + 32 * idx2);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = (int)idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW && static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                    match12[idx1] = bestIdx2; vbMatched2[bestIdx2] = true;
                    if (checkOri) rotHist[rot_bin(ang1[idx1], ang2[bestIdx2], nBinsAngle)].push_back((int)idx1);
                    nmatches++;
                }
            }
            f1it++; f2it++;
        } else if (f1it->first < f2it->first) f1it = fv1.lower_bound(f2it->first);
        else f2it = fv2.lower_bound(f1it->first);
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotHist.data(), nBinsAngle, ind1, ind2, ind3);
        for (int i = 0; i < nBinsAngle; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { match12[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// All-pairs matcher (defined above). match12[i] = column of B or -1; dist12[i] = best distance of row i
// (always reported, also for rejected rows); second12[i] = second-best distance.
static inline int match_bruteforce(const uint8_t* descA, const float* angA, int nA, const uint8_t* descB, const float* angB,
                                   int nB, float nnratio, int thLow, bool checkOri, int* match12, int* dist12, int* second12) {
    int nmatches = 0;
    const int nBinsAngle = (int)std::ceil(360.0f / HISTO_LENGTH);
    std::vector<std::vector<int>> rotHist(nBinsAngle);
    for (int i = 0; i < nA; i++) {
        int best1 = 256, best2 = 256, bestIdx = -1;
        for (int j = 0; j < nB; j++) {
            const int dist = descriptor_distance(descA + 32 * (size_t)i, descB + 32 * (size_t)j);
            if (dist < best1) { best2 = best1; best1 = dist; bestIdx = j; }
            else if (dist < best2) best2 = dist;
        }
        match12[i] = -1; dist12[i] = best1; second12[i] = best2;
        if (bestIdx >= 0 && best1 <= thLow && static_cast<float>(best1) < nnratio * static_cast<float>(best2)) {
            match12[i] = bestIdx;
            if (checkOri) rotHist[rot_bin(angA[i], angB[bestIdx], nBinsAngle)].push_back(i);
            nmatches++;
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotHist.data(), nBinsAngle, ind1, ind2, ind3);
        for (int i = 0; i < nBinsAngle; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { match12[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

}  // namespace orc
