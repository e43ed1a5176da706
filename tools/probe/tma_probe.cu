// Standalone probe: 3-D u8 TMA tile load (param-space and global-space tensor maps), run on the GPU box.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s line %d\n", #x, cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// V: 0 = param map, 1 = global map; both with the tile at a 128-byte aligned dynamic-smem address and the barrier behind it
template <int V>
__global__ void probe(const __grid_constant__ CUtensorMap tmP, const CUtensorMap* tmG, int x, int y, int z, int bytes, uint8_t* out, unsigned* info, int stage) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned base = smem_u32(smem_dyn);
    unsigned tileA = (base + 127u) & ~127u;
    unsigned bar = tileA + 4096;
    unsigned char* tile = smem_dyn + (tileA - base);
    if (threadIdx.x == 0) {
        info[0] = base; info[1] = tileA;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (stage >= 1) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        const CUtensorMap* tm = V == 1 ? tmG : &tmP;
        if (stage >= 2) asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(tileA), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
    }
    __syncwarp();
    if (stage == 2) { for (int k = 0; k < 100000; k++) __nanosleep(100); }
    if (stage >= 3) asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes; i += 32) out[i] = tile[i];
}

int main(int argc, char** argv) {
    const int V = argc > 1 ? atoi(argv[1]) : 0;
    const int BW = argc > 2 ? atoi(argv[2]) : 48, BH = argc > 3 ? atoi(argv[3]) : 42;
    const int stage = argc > 4 ? atoi(argv[4]) : 3;
    const int X0 = argc > 5 ? atoi(argv[5]) : 0;
    const int W = 240, H = 180, F = 2, P = 240;
    printf("variant %d box %dx%d\n", V, BW, BH);
    std::vector<uint8_t> img((size_t)P * H * F);
    for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d_img, *d_out; CUtensorMap* d_tm; unsigned* d_info;
    CK(cudaMalloc(&d_img, img.size())); CK(cudaMalloc(&d_out, BW * BH)); CK(cudaMalloc(&d_tm, sizeof(CUtensorMap))); CK(cudaMalloc(&d_info, 16));
    CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[3] = {W, H, F}; cuuint64_t strides[2] = {P, (cuuint64_t)P * H};
    cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode r=%d\n", (int)r);
    CK(cudaMemcpy(d_tm, &tm, sizeof(tm), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    CK(cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    for (int t = 0; t < 3; t++) {
        const int x = t == 0 ? X0 : (t == 1 ? X0 + 192 : X0 - 16), y = t == 0 ? 6 : (t == 1 ? 150 : -2), z = t & 1;
        CK(cudaMemset(d_out, 0xEE, BW * BH));
        if (V == 1) probe<1><<<1, 32, 16384>>>(tm, d_tm, x, y, z, BW * BH, d_out, d_info, stage);
        else probe<0><<<1, 32, 16384>>>(tm, d_tm, x, y, z, BW * BH, d_out, d_info, stage);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("case %d: kernel error %s\n", t, cudaGetErrorString(e)); return 2; }
        std::vector<uint8_t> out(BW * BH); unsigned info[4];
        CK(cudaMemcpy(out.data(), d_out, out.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(info, d_info, 16, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int r2 = 0; r2 < BH; r2++) for (int c = 0; c < BW; c++) {
            const int gx = x + c, gy = y + r2;
            const uint8_t exp = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? img[(size_t)z * P * H + (size_t)gy * P + gx] : 0;
            bad += out[r2 * BW + c] != exp;
        }
        printf("case %d (x=%d y=%d z=%d): %d bad bytes; smem base 0x%x tile 0x%x\n", t, x, y, z, bad, info[0], info[1]);
    }
    return 0;
}
