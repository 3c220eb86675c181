// hbma_dev.cuh -- helpers shared by the HBMA kernels (k_hbma.cu, k_hbma_pool.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace svc {

// acc + sum |a.b[i] - b.b[i]| in ONE instruction (VABSDIFF4.U8.ACC with a live accumulator);
// the __vsadu4() + add form compiles to VABSDIFF4 ..., RZ plus an IADD3 tree, i.e. 1.5x
// the integer-ALU work.
__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t acc) {
  uint32_t d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
  return d;
}

// ---- host side: tensor maps ------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

inline bool encode_box(CUtensorMap* m, const uint8_t* base, uint32_t w, uint32_t h, uint32_t pitch,
                       uint64_t slot_bytes, uint32_t n_slots, uint32_t box_w, uint32_t box_h) {
  EncodeTiledFn fn = get_encode_tiled();
  if (!fn) return false;
  const cuuint64_t dims[3] = {w, h, n_slots};
  const cuuint64_t strides[2] = {pitch, slot_bytes};
  const cuuint32_t box[3] = {box_w, box_h, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


struct EbmaMaps {
  CUtensorMap t;  // tracked window box of one level
  CUtensorMap a;  // anchor block / anchor tile box of the same level
};

// mid ranges (r = 3..8), one level per launch, windows realigned in registers (k_hbma_rs.cu)
struct HbmaParams;
bool rs_level_supported(const HbmaParams& p);
cudaError_t launch_rs_level(const HbmaParams& p, uint32_t lvl, bool top, cudaStream_t st);

// the encoder default (16x16 blocks, 4 levels, r = 1): strip-per-lane tile kernel (k_hbma_strip.cu)
bool strip_supported(const HbmaParams& p);
cudaError_t launch_strip(const HbmaParams& p, cudaStream_t st, int* extra_launches);

// large-range 16x16 search with pooled work items and pre-shifted window copies (k_hbma_pool.cu);
// returns false when the configuration is outside its limits (the caller picks another kernel)
// (*extra_launches += launches beyond the first, for the level-synchronous path)
bool try_launch_pool(const HbmaParams& p, cudaStream_t st, cudaError_t* err, int* extra_launches);
// hbma_tile_kernel (k_hbma.cu) over the three coarsest levels of a 5-level pyramid (r = 3, 4): vectors
// and MADs of level 2 land in p.mv / p.mad for the refinement launches of k_hbma_pool.cu
cudaError_t launch_tile_upper3(const HbmaParams& p, cudaStream_t st);

}  // namespace svc
