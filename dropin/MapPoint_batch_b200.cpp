// Batched MapPoint::ComputeDistinctiveDescriptors on libcubemap_b200.so. The reference calls the member once per MapPoint (a dozen 32-byte rows: far
// below one kernel launch), so the member itself stays the reference's; what moves to the device is the LOOP over MapPoints at the call sites that
// refresh many points at once - LocalMapping::ProcessNewKeyFrame (src/LocalMapping.cpp:128-160), SearchInNeighbors (:437-452) and
// Tracking::CreateInitialMapMonocular. INTEGRATION.md shows the call-site change (one added setter in MapPoint, one replaced loop).
//   reference: src/MapPoint.cpp:243-303
#include "MapPoint.h"
#include "KeyFrame.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "cubemap_b200.h"

namespace cubemap_b200 {

// For every MapPoint (NULL / bad / unobserved ones get an empty Mat) the descriptor ComputeDistinctiveDescriptors would store in mDescriptor.
void DistinctiveDescriptors(const std::vector<MapPoint*>& vpMPs, std::vector<cv::Mat>& out) {
    static thread_local cslam_mapper* m = nullptr;
    if (!m && cslam_mapper_create(&m, 0) != CSLAM_OK) { std::fprintf(stderr, "MapPoint (cubemap_b200): %s\n", cslam_last_error()); std::exit(EXIT_FAILURE); }
    const int n = (int)vpMPs.size();
    out.assign(n, cv::Mat());
    std::vector<int32_t> offset(n + 1, 0), best(n, -1);
    std::vector<uint8_t> desc; std::vector<cv::Mat> rows;
    for (int p = 0; p < n; p++) {
        MapPoint* pMP = vpMPs[p];
        if (pMP && !pMP->isBad()) {
            const std::map<KeyFrame*, size_t> obs = pMP->GetObservations();   // iterated in the map's order, like the reference
            for (std::map<KeyFrame*, size_t>::const_iterator it = obs.begin(); it != obs.end(); ++it) {
                KeyFrame* pKF = it->first;
                if (pKF->isBad()) continue;
                const cv::Mat row = pKF->mDescriptors.row(it->second);
                rows.push_back(row);
                desc.insert(desc.end(), row.ptr<uchar>(0), row.ptr<uchar>(0) + 32);
            }
        }
        offset[p + 1] = (int32_t)rows.size();
    }
    if (cslam_distinctive_descriptors(m, desc.data(), offset.data(), n, best.data()) != CSLAM_OK) { std::fprintf(stderr, "MapPoint (cubemap_b200): %s\n", cslam_last_error()); std::exit(EXIT_FAILURE); }
    for (int p = 0; p < n; p++) if (best[p] >= 0) out[p] = rows[offset[p] + best[p]].clone();
}

}  // namespace cubemap_b200
