/* ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C restatement of the reference's hot-path algorithms (Asherchi/Stereo_Match_Traditional).
 * Every function cites the reference file:line it restates.  Parity status: PINNED -- each function is
 * compared bit-for-bit (disparity-for-disparity for NCC/ASW) against the compiled, unmodified reference
 * (oracle/_ref/libsmt_ref.so) by tests/test_oracle_port_vs_ref.py in the build container, and against the
 * golden fixtures in tests/golden/ (outputs of the compiled reference, see oracle/make_golden.py) everywhere.
 *
 * The symbols mirror oracle/ref_harness.cpp's ref_* entry points one for one (prefix orc_), so the Python
 * loader (oracle/oracle.py) drives either library through the same code.
 */
#ifndef SMT_ORACLE_PORT_H_
#define SMT_ORACLE_PORT_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int orc_abi_version(void);
int orc_num_threads(void);

void orc_adcensus(const float* left, const float* right, int H, int W, int D, float sigmaC, float sigmaS, int do_left,
                  int do_right, float* costL, float* costR, float* adL, float* adR, float* censusL, float* censusR,
                  float* wtaL, float* wtaR);
void orc_wta(const float* vol, int H, int W, int D, float* disp);
void orc_scanline(const float* cost, const float* gray, int H, int W, int D, int p1, int p2, float* left,
                  float* right, float* up, float* down, float* sum, float* disp);
void orc_lrcheck(float* dispL, const float* dispR, int W, int H, int gate, int32_t* occ, int32_t* n_occ, int32_t* mis,
                 int32_t* n_mis);
void orc_lrcheck_and(const float* dispL, const float* dispR, float* last, int W, int H, float gate, int32_t* occ,
                     int32_t* n_occ, int32_t* mis, int32_t* n_mis);
void orc_remove_speckles_inf(float* disp, int W, int H, int diff, unsigned min_area);
void orc_remove_speckles(float* disp, int W, int H, int diff, unsigned min_area, int invalid_val);
void orc_median(const float* in, float* out, int W, int H, int k);
void orc_transform_to_show(const float* disp, int H, int W, uint8_t* out);
void orc_fill_the_hole_fixed(int row, int col, int dispRange, float* disp, const int32_t* occ, int32_t n_occ, int32_t* mis,
                             int32_t* n_mis, int32_t mis_capacity);
void orc_fill_the_hole(int row, int col, int dispRange, float* disp, const int32_t* occ, int32_t n_occ, int32_t* mis, int32_t* n_mis,
                       int32_t mis_capacity);
int orc_cross_aggregate(const uint8_t* bgrL, const uint8_t* bgrR, const float* cost_init, int W, int H, int D, int L1,
                        int L2, int t1, int t2, int iters, float* cost_out, uint8_t* arms, uint16_t* sup0,
                        uint16_t* sup1);
int orc_crossarm(const uint8_t* img, int ch, int H, int W, int D, int tao, int32_t* armL, int32_t* armR,
                 int32_t* armT, int32_t* armB, int mode, const float* vol_in, float* vol_out);
void orc_cblsm_ad(const uint8_t* left, const uint8_t* right, int H, int W, int D, float* adL, float* adR);
void orc_cblsm_arms(const uint8_t* img, int ch, int H, int W, int tao, int maxLength, int secLength, int32_t* armL,
                    int32_t* armR, int32_t* armUp, int32_t* armDown);
void orc_cblsm_aggregate_v5(const float* vol_in, float* vol_out, const int32_t* armL, const int32_t* armR,
                            const int32_t* armUp, const int32_t* armDown, int D, int H, int W);
void orc_cblsm_wta(const float* vol, float* disp, int D, int H, int W);
void orc_sad_left(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int D, int winsize, int32_t* disp);
void orc_sad_right(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int D, int winsize, int32_t* disp);
void orc_sad_costs_left(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int D, int winsize, float* costs);
float orc_sad_optimal_disparity(const float* sad, int D);
void orc_ncc(const uint8_t* L, const uint8_t* R, int H, int W, int winSize, int D, int32_t* disp);
double orc_ncc_cost(const uint8_t* L, const uint8_t* R, int H, int W, int winSize, int i, int j, int d);
void orc_asw(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int winSize, int D, double spaceSigma,
             double colorSigma, int T, int right_view, float* disp);
void orc_asw_masks(int winSize, double spaceSigma, double colorSigma, double* space_out, double* color_out);
void orc_asw_crosscheck(const float* dispL, const float* dispR, int H, int W, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
