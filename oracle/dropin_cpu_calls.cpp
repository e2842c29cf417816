// ORACLE / TEST HARNESS: compiled with -DSearchByBoW=SearchByBoW_cpu (like the reference's src/ORBMatcher.cpp in libdropin.so), so that
// `SearchByBoW` below names the reference's own CPU body; the rest of the harness sees the un-renamed class and therefore the GPU drop-in.
#include "ORBMatcher.h"
int call_cpu_search_by_bow(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& v, float nnratio, bool ori) {
    ORBMatcher m(nnratio, ori);
    return m.SearchByBoW(pKF, F, v);
}
