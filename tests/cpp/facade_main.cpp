// Drives the C++ facade (include/smt_b200.hpp) in the order the reference's main() calls its classes
// (AD-CensusV1/main.cpp:58-94 with the commented-out scanline / post-processing calls enabled, and
// CBLSM/CBLSM.cpp:138-143 for the cross aggregation).  Usage: facade_main in.bin out.bin
//   in.bin : int32 H, W, D, then gray_left[H*W] u8, gray_right[H*W] u8, bgr_left[H*W*3] u8, bgr_right[H*W*3] u8
//   out.bin: float disp_final[H*W], float disp_left_raw[H*W], float disp_right_raw[H*W]
#include <cstdio>
#include <vector>

#include "smt_b200.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 3;
    int32_t hdr[3];
    if (fread(hdr, 4, 3, f) != 3) return 4;
    const int H = hdr[0], W = hdr[1], D = hdr[2];
    const size_t n = (size_t)H * W;
    std::vector<uint8_t> gl(n), gr(n), bl(n * 3), br(n * 3);
    if (fread(gl.data(), 1, n, f) != n || fread(gr.data(), 1, n, f) != n || fread(bl.data(), 1, n * 3, f) != n * 3 ||
        fread(br.data(), 1, n * 3, f) != n * 3)
        return 5;
    fclose(f);
    try {
        using namespace smt;
        std::vector<float> leftptr(gl.begin(), gl.end()), rightptr(gr.begin(), gr.end());  // main.cpp:46-55
        ImageView leftImage(gl.data(), H, W), rightImage(gr.data(), H, W);
        AD_Census adcensus;
        adcensus.Initialize(leftptr.data(), rightptr.data(), D, H, W, leftImage, rightImage, 10, 30);  // main.cpp:58
        adcensus.ComputeADcensus();
        adcensus.ComputeADcensusRight();

        CrossAggregator aggL, aggR;  // CBLSM.cpp:138-143
        aggL.Initialize(W, H, 0, D);
        aggL.SetData(bl.data(), br.data(), adcensus.GetPtrLeft());
        aggL.SetParams(34, 17, 20, 6);
        aggL.Aggregate(4);
        aggR.Initialize(W, H, 0, D);
        aggR.SetData(br.data(), bl.data(), adcensus.GetPtrRight());
        aggR.SetParams(34, 17, 20, 6);
        aggR.Aggregate(4);

        std::vector<float> dispL(n), dispR(n), out(n);
        ScanlineOptimizer so;  // main.cpp:86-89
        so.Initialize(H, W, D, aggL.get_cost_ptr(), 10, 150);
        so.ScanLine(aggL.get_cost_ptr(), leftptr.data());
        so.WTA(dispL.data());
        CrossArmAggregation wta;  // main.cpp:84 -- first-minimum WTA of the right volume
        wta.Initialize(H, W, nullptr, nullptr, 30, D);
        wta.WTA(aggR.get_cost_ptr(), dispR.data());
        std::vector<float> rawL = dispL;

        std::vector<std::pair<int, int>> occ, mis;
        LeftRightConsistency(W, H, 2, dispL.data(), dispR.data(), occ, mis);  // main.cpp:92
        RemoveSpeckles(dispL.data(), W, H, 1, 30, Invalid_Float);             // main.cpp:93
        MedianFilter(dispL.data(), out.data(), W, H, 3);                      // main.cpp:94

        FILE* g = fopen(argv[2], "wb");
        if (!g) return 6;
        fwrite(out.data(), 4, n, g);
        fwrite(rawL.data(), 4, n, g);
        fwrite(dispR.data(), 4, n, g);
        fclose(g);
        printf("facade ok: %dx%dx%d, %zu occlusions, %zu mismatches\n", H, W, D, occ.size(), mis.size());
    } catch (const smt::Error& e) {
        fprintf(stderr, "smt error %d: %s\n", e.status, e.what());
        return 1;
    }
    return 0;
}
