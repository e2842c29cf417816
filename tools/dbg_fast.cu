// debug harness: runs k_fast on a random image and prints internals for one pixel
#define CSLAM_DEBUG_FAST 1
#include "../cubemapslam_b200/csrc/frontend.cu"
#include "../cubemapslam_b200/csrc/capi_common.cu"
#include <cstdlib>
#include <set>
#include <tuple>
#include "../oracle/cvprim.h"
int main() {
    const int w = 200, h = 200, pitch = 256;
    std::vector<uint8_t> img(pitch * h);
    srand(1);
    for (int y = 0; y < h; y++) for (int x = 0; x < pitch; x++) img[y * pitch + x] = (uint8_t)((x * 7 + y * 13 + (rand() % 40)) & 255);
    LevelGeom g{}; g.w = w; g.h = h; g.pitch = pitch; g.minB = 16; g.maxBX = w - 16; g.maxBY = h - 16;
    float width = g.maxBX - g.minB; int nCols = width / 30.f; g.wCell = (int)ceil(width / nCols); g.hCell = g.wCell;
    g.nColsEff = 0; for (int j = 0; j < nCols; j++) if (!((float)(g.minB + j * g.wCell) >= g.maxBX - 6)) g.nColsEff = j + 1;
    g.nRowsEff = g.nColsEff; g.candCap = 100000;
    uint8_t* d_img; uint32_t* d_cand; uint32_t* d_cnt; int* d_err;
    cudaMalloc(&d_img, img.size()); cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&d_cand, 4 * 100000); cudaMalloc(&d_cnt, 4); cudaMemset(d_cnt, 0, 4); cudaMalloc(&d_err, 4); cudaMemset(d_err, 0, 4);
    dim3 grid(cdiv(g.nColsEff, CG), g.nRowsEff, 1);
    k_fast<<<grid, 256>>>(d_img, g, 20, 7, d_cand, d_cnt, 1, d_err, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t n; cudaMemcpy(&n, d_cnt, 4, cudaMemcpyDeviceToHost);
    printf("sync %s, n=%u wCell=%d ncols=%d\n", cudaGetErrorString(e), n, g.wCell, g.nColsEff);
    // host ref for pixel (dbg) image coords (16+3+5, 16+3+4)
    int X = 16 + 3 + 5, Y = 16 + 3 + 4;
    printf("host pixel c=%d ring:", img[Y * pitch + X]);
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1}, dy[16] = {-3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3};
    for (int k = 0; k < 16; k++) printf(" %d", img[(Y + dy[k]) * pitch + X + dx[k]]);
    printf("\n");
    std::vector<uint32_t> cand(n); cudaMemcpy(cand.data(), d_cand, 4 * n, cudaMemcpyDeviceToHost);
    std::set<std::tuple<int,int,int>> G, R;
    for (uint32_t v : cand) G.insert({(int)(v & 0xfff), (int)((v >> 12) & 0xfff), (int)(v >> 24)});
    std::vector<orc::FastKp> kc;
    for (int i = 0; i < g.nRowsEff; i++) for (int j = 0; j < g.nColsEff; j++) {
        int iniY = g.minB + i * g.hCell, maxY = std::min(iniY + g.hCell + 6, g.maxBY), iniX = g.minB + j * g.wCell, maxX = std::min(iniX + g.wCell + 6, g.maxBX);
        orc::fast_nms(img.data() + iniY * pitch + iniX, maxX - iniX, maxY - iniY, pitch, 20, kc);
        if (kc.empty()) orc::fast_nms(img.data() + iniY * pitch + iniX, maxX - iniX, maxY - iniY, pitch, 7, kc);
        for (auto& k : kc) R.insert({k.x + j * g.wCell, k.y + i * g.hCell, k.response});
    }
    printf("gpu %zu ref %zu\n", G.size(), R.size());
    int shown = 0;
    for (auto& t : G) if (!R.count(t) && shown++ < 10) printf("only gpu (%d,%d,%d)\n", std::get<0>(t), std::get<1>(t), std::get<2>(t));
    shown = 0;
    for (auto& t : R) if (!G.count(t) && shown++ < 10) printf("only ref (%d,%d,%d)\n", std::get<0>(t), std::get<1>(t), std::get<2>(t));
    return 0;
}
