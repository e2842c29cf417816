/* Optimizer.h — facade with the reference's static Optimizer signatures for the cubemap BA path (reference
 * include/Optimizer.h:42-63; src/Optimizer.cpp:48-451) on top of libcubemap_b200.so. Templates over the reference's
 * Frame / KeyFrame / MapPoint / Map types: the window collection (covisible KFs, local MapPoints, fixed KFs) is the
 * reference's own code path (src/Optimizer.cpp:194-243), the g2o graph is replaced by flat arrays handed to the C ABI. */
#ifndef CSLAM_OPTIMIZER_H
#define CSLAM_OPTIMIZER_H
#include <cstdio>
#include <cstdlib>
#include <list>
#include <map>
#include <mutex>
#include <type_traits>
#include <vector>
#include "cubemap_b200.h"
#include "cv_compat.h"

class Optimizer {
public:
    // reference src/Optimizer.cpp:48-190. CamT supplies GetCosFovTh(), GetCubeFaceWidth/Height() (CamModelGeneral singleton).
    template <class FrameT, class CamT>
    static int PoseOptimization(FrameT* pFrame, CamT* cam) {
        const int N = pFrame->N;
        std::vector<float> Xw, kp, w; std::vector<int> idx;
        for (int i = 0; i < N; i++) {
            if (pFrame->mvKeyRays[i](2) < cam->GetCosFovTh()) continue;
            if (!pFrame->mvpMapPoints[i]) continue;
            pFrame->mvbOutlier[i] = false;
            const cv::Mat P = pFrame->mvpMapPoints[i]->GetWorldPos();
            Xw.push_back(P.template at<float>(0)); Xw.push_back(P.template at<float>(1)); Xw.push_back(P.template at<float>(2));
            kp.push_back(pFrame->mvKeys[i].pt.x); kp.push_back(pFrame->mvKeys[i].pt.y);
            w.push_back(pFrame->mvInvLevelSigma2[pFrame->mvKeys[i].octave]); idx.push_back(i);
        }
        const int n = (int)idx.size();
        if (n < 3) return 0;
        float T[16];
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T[4 * r + c] = pFrame->mTcw.template at<float>(r, c);
        std::vector<uint8_t> out(n); int32_t off[2] = {0, n}, inl = 0;
        if (cslam_pose_optimization(handle(), 1, off, T, Xw.data(), kp.data(), w.data(), cam->GetCubeFaceWidth(), cam->GetCubeFaceHeight(), out.data(), &inl, NULL) != CSLAM_OK) fatal();
        for (int k = 0; k < n; k++) pFrame->mvbOutlier[idx[k]] = out[k] != 0;
        cv::Mat pose; pose.create(4, 4, CV_32F);
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) pose.template at<float>(r, c) = T[4 * r + c];
        pFrame->SetPose(pose);
        return inl;
    }

    // reference src/Optimizer.cpp:192-451
    template <class KeyFrameT, class MapT, class CamT>
    static void LocalBundleAdjustment(KeyFrameT* pKF, bool* pbStopFlag, MapT* pMap, CamT* cam) {
        typedef typename std::remove_pointer<typename std::decay<decltype(pKF->GetMapPointMatches()[0])>::type>::type MapPointT;
        std::list<KeyFrameT*> lLocalKeyFrames; lLocalKeyFrames.push_back(pKF); pKF->mnBALocalForKF = pKF->mnId;
        const std::vector<KeyFrameT*> vNeighKFs = pKF->GetVectorCovisibleKeyFrames();
        for (size_t i = 0; i < vNeighKFs.size(); i++) { vNeighKFs[i]->mnBALocalForKF = pKF->mnId; if (!vNeighKFs[i]->isBad()) lLocalKeyFrames.push_back(vNeighKFs[i]); }
        std::list<MapPointT*> lLocalMapPoints;
        for (auto lit = lLocalKeyFrames.begin(); lit != lLocalKeyFrames.end(); ++lit)
            for (MapPointT* pMP : (*lit)->GetMapPointMatches())
                if (pMP && !pMP->isBad() && pMP->mnBALocalForKF != pKF->mnId) { lLocalMapPoints.push_back(pMP); pMP->mnBALocalForKF = pKF->mnId; }
        std::list<KeyFrameT*> lFixedCameras;
        for (MapPointT* pMP : lLocalMapPoints)
            for (auto& ob : pMP->GetObservations()) {
                KeyFrameT* pKFi = ob.first;
                if (pKFi->mnBALocalForKF != pKF->mnId && pKFi->mnBAFixedForKF != pKF->mnId) { pKFi->mnBAFixedForKF = pKF->mnId; if (!pKFi->isBad()) lFixedCameras.push_back(pKFi); }
            }
        // vertices in g2o's order: keyframes by mnId, points by mnId
        std::map<unsigned long, KeyFrameT*> kfs; std::map<unsigned long, bool> fixedFlag;
        for (KeyFrameT* k : lLocalKeyFrames) { kfs[k->mnId] = k; fixedFlag[k->mnId] = (k->mnId == 0); }
        for (KeyFrameT* k : lFixedCameras) { kfs[k->mnId] = k; fixedFlag[k->mnId] = true; }
        std::map<unsigned long, MapPointT*> mps;
        for (MapPointT* m : lLocalMapPoints) mps[m->mnId] = m;
        std::map<KeyFrameT*, int> kfIndex; std::vector<float> Tcw; std::vector<uint8_t> fixed; std::vector<KeyFrameT*> kfList;
        for (auto& kv : kfs) {
            kfIndex[kv.second] = (int)kfList.size(); kfList.push_back(kv.second); fixed.push_back(fixedFlag[kv.first]);
            const cv::Mat T = kv.second->GetPose();
            for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) Tcw.push_back(T.template at<float>(r, c));
        }
        std::vector<float> pts, kp, w; std::vector<int32_t> eMP, eKF; std::vector<MapPointT*> mpList; std::vector<std::pair<KeyFrameT*, MapPointT*> > edgeOwner;
        for (auto& kv : mps) {
            MapPointT* pMP = kv.second; const int li = (int)mpList.size(); mpList.push_back(pMP);
            const cv::Mat P = pMP->GetWorldPos();
            pts.push_back(P.template at<float>(0)); pts.push_back(P.template at<float>(1)); pts.push_back(P.template at<float>(2));
            for (auto& ob : pMP->GetObservations()) {
                KeyFrameT* pKFi = ob.first;
                if (pKFi->isBad() || !kfIndex.count(pKFi)) continue;
                if (pKFi->mvKeyRays[ob.second](2) < cam->GetCosFovTh()) continue;       // reference :323-325
                const cv::KeyPoint& k = pKFi->mvKeys[ob.second];
                eMP.push_back(li); eKF.push_back(kfIndex[pKFi]); kp.push_back(k.pt.x); kp.push_back(k.pt.y);
                w.push_back(pKFi->mvInvLevelSigma2[k.octave]); edgeOwner.push_back(std::make_pair(pKFi, pMP));
            }
        }
        cslam_ba_problem p; p.n_kf = (int)kfList.size(); p.n_mp = (int)mpList.size(); p.n_edges = (int)eMP.size();
        p.Tcw = Tcw.data(); p.kf_fixed = fixed.data(); p.points = pts.data(); p.edge_mp = eMP.data(); p.edge_kf = eKF.data(); p.kp_xy = kp.data();
        p.inv_sigma2 = w.data(); p.face_w = cam->GetCubeFaceWidth(); p.face_h = cam->GetCubeFaceHeight();
        std::vector<uint8_t> outlier(eMP.size() + 1);
        cslam_ba_result r; r.outlier = outlier.data(); r.pose_fp64 = NULL; r.points_fp64 = NULL; r.lm_log = NULL; r.log_cap = 0;
        if (pbStopFlag && *pbStopFlag) return;
        if (cslam_local_ba(handle(), &p, reinterpret_cast<const volatile uint8_t*>(pbStopFlag), 5, 10, &r) != CSLAM_OK) fatal();
        std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);
        for (size_t e = 0; e < edgeOwner.size(); e++) if (outlier[e] && !edgeOwner[e].second->isBad()) {
            edgeOwner[e].first->EraseMapPointMatch(edgeOwner[e].second); edgeOwner[e].second->EraseObservation(edgeOwner[e].first);
        }
        for (KeyFrameT* k : lLocalKeyFrames) {
            cv::Mat T; T.create(4, 4, CV_32F); const float* s = &Tcw[16 * kfIndex[k]];
            for (int r2 = 0; r2 < 4; r2++) for (int c = 0; c < 4; c++) T.template at<float>(r2, c) = s[4 * r2 + c];
            k->SetPose(T);
        }
        for (size_t i = 0; i < mpList.size(); i++) {
            cv::Mat P; P.create(3, 1, CV_32F);
            for (int c = 0; c < 3; c++) P.template at<float>(c) = pts[3 * i + c];
            mpList[i]->SetWorldPos(P); mpList[i]->UpdateNormalAndDepth();
        }
    }

protected:
    static cslam_optimizer* handle() {
        static thread_local cslam_optimizer* o = nullptr;
        if (!o && cslam_optimizer_create(&o, 0) != CSLAM_OK) fatal();
        return o;
    }
    static void fatal() { std::fprintf(stderr, "Optimizer (cubemap_b200): %s\n", cslam_last_error()); std::exit(EXIT_FAILURE); }
};
#endif
