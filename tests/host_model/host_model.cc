// tests/host_model/host_model.cc — compiles the product's shared host/device sources as plain C++
// (EORB_HOST_MODEL: one "thread", no atomics) so the CPU test-suite can check the data-parallel
// formulations and scalar arithmetic of the CUDA kernels against the serial oracle without a GPU.
#define EORB_HOST_MODEL 1
#include <cmath>
#include <cstring>
#include <vector>
#include "../../eorb_slam_b200/csrc/eorb_math.cuh"
#include "../../eorb_slam_b200/csrc/octree_core.cuh"

extern "C" {

int hm_octree(const uint32_t* keys, int n, int width, int height, int N, uint32_t* out, int out_cap) {
    int nIni = (int)std::round((float)width / (float)height);
    if (nIni <= 0) return 0;
    float hX = (float)width / (float)nIni;
    int nodeCap = (N + 3 > 4 * nIni ? N + 3 : 4 * nIni) + 1;
    std::vector<unsigned char> smem(eorb::oct_smem_bytes(nodeCap));
    std::vector<uint16_t> knode(n > 0 ? n : 1);
    std::vector<uint32_t> o(nodeCap);
    int r = eorb::oct_distribute(keys, knode.data(), n, width, height, nIni, hX, N, nodeCap, smem.data(), o.data());
    for (int i = 0; i < r && i < out_cap; i++) out[i] = o[i];
    return r;
}

int hm_fast_max_arc_min(int v, const int* ring16) { return eorb::fast_max_arc_min(v, ring16); }
float hm_fast_atan2(float y, float x) { return eorb::fast_atan2_deg(y, x); }
void hm_brief_offset(int px, int py, float a, float b, int* row, int* col) { eorb::brief_offset(px, py, a, b, *row, *col); }
int hm_resize_px(int p00, int p01, int p10, int p11, int a0, int a1, int b0, int b1) {
    return eorb::resize_vsum(eorb::resize_hsum(p00, p01, a0, a1), eorb::resize_hsum(p10, p11, a0, a1), b0, b1);
}
int hm_hamming(const uint32_t* a, const uint32_t* b) { return eorb::hamming256(a, b); }

}
