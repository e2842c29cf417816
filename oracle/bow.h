// ORACLE (test infrastructure, not product code): CPU restatement of DBoW2's vocabulary-tree transform as CubemapSLAM uses it
// (Frame::ComputeBoW / KeyFrame::ComputeBoW -> mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4), src/Frame.cpp:719-726).
//   TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup)   ThirdParty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1193
//   TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)          :1218-1262 (greedy descent, first minimum wins)
//   FORB::distance   ThirdParty/DBoW2/DBoW2/FORB.cpp:81-101 ;  BowVector::addWeight / normalize   BowVector.cpp:31-45,62-85
// ORBvoc.txt is "10 6 0 0": k = 10, L = 6, L1_NORM scoring, TF_IDF weighting -> weights are summed per word, then L1-normalised.
// Pinned against the reference's own DBoW2 compiled in oracle/_ref/libref.so (tests/test_oracle_bow.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
#include "orb_matcher.h"

namespace orc {

struct Vocabulary {
    int k = 0, L = 0;
    std::vector<int> parent, wordId; std::vector<uint8_t> desc; std::vector<double> weight; std::vector<std::vector<int> > children;   // node 0 = root
    // nodes 1..n in file order: parent id, leaf flag, 32 descriptor bytes, weight (loadFromTextFile :1337-1415)
    void build(int k_, int L_, int n, const int* par, const uint8_t* isLeaf, const uint8_t* d, const double* w) {
        k = k_; L = L_;
        parent.assign(n + 1, 0); wordId.assign(n + 1, -1); desc.assign((size_t)(n + 1) * 32, 0); weight.assign(n + 1, 0.0); children.assign(n + 1, std::vector<int>());
        int words = 0;
        for (int i = 0; i < n; i++) {
            const int nid = i + 1;
            parent[nid] = par[i]; children[par[i]].push_back(nid);
            std::copy(d + 32 * (size_t)i, d + 32 * (size_t)i + 32, desc.begin() + 32 * (size_t)nid);
            weight[nid] = w[i];
            if (isLeaf[i]) wordId[nid] = words++;
        }
    }
    bool is_leaf(int id) const { return children[id].empty(); }
    void transform_feature(const uint8_t* f, int levelsup, int& word, double& w, int& nid) const {
        const int nid_level = L - levelsup;
        nid = 0;   // `if(nid_level <= 0 && nid != NULL) *nid = 0` - and NodeId is otherwise left as the caller initialised it; callers pass a fresh variable
        int final_id = 0, current_level = 0;
        do {
            ++current_level;
            const std::vector<int>& nodes = children[final_id];
            final_id = nodes[0];
            double best_d = descriptor_distance(f, &desc[32 * (size_t)final_id]);
            for (size_t c = 1; c < nodes.size(); c++) {
                const double dd = descriptor_distance(f, &desc[32 * (size_t)nodes[c]]);
                if (dd < best_d) { best_d = dd; final_id = nodes[c]; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (!is_leaf(final_id));
        word = wordId[final_id]; w = weight[final_id];
    }
    // BowVector as parallel sorted arrays; nodeOf[i] = FeatureVector node of feature i (-1: stopped word, weight 0)
    void transform(const uint8_t* feats, int n, int levelsup, std::vector<int>& bowWord, std::vector<double>& bowVal, std::vector<int>& nodeOf, std::vector<int>& wordOf) const {
        std::vector<std::pair<int, double> > v;   // sorted by word: std::map<WordId, WordValue>
        nodeOf.assign(n, -1); wordOf.assign(n, -1);
        for (int i = 0; i < n; i++) {
            int word, nid; double w;
            transform_feature(feats + 32 * (size_t)i, levelsup, word, w, nid);
            wordOf[i] = word;
            if (w > 0) {
                auto it = std::lower_bound(v.begin(), v.end(), std::make_pair(word, -1e300));
                if (it != v.end() && it->first == word) it->second += w; else v.insert(it, std::make_pair(word, w));
                nodeOf[i] = nid;
            }
        }
        double norm = 0.0;   // L1 scoring: mustNormalize -> BowVector::normalize(L1), summed in word order
        for (auto& e : v) norm += std::fabs(e.second);
        if (norm > 0.0) for (auto& e : v) e.second /= norm;
        bowWord.clear(); bowVal.clear();
        for (auto& e : v) { bowWord.push_back(e.first); bowVal.push_back(e.second); }
    }
};

}  // namespace orc
