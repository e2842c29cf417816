"""cubemapslam_b200 — B200-native (sm_100a) hot path of CubemapSLAM: warp, ORB extraction, Hamming matching,
cubemap-edge bundle adjustment. The CUDA library is mandatory on the product path (no CPU fallback)."""
