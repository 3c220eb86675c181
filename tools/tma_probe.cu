// tma_probe.cu -- standalone probe of cp.async.bulk.tensor.3d on sm_100a (debug aid).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int z, int bytes, uint8_t* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"((uint32_t)bytes) : "memory");
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(d), "l"(&map), "r"(x), "r"(y), "r"(z), "r"(bar_addr) : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar_addr), "r"(0u) : "memory");
  }
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

int main(int argc, char** argv) {
  // args: W H pitch slots boxw boxh x y z
  if (argc < 10) return 2;
  int W = atoi(argv[1]), H = atoi(argv[2]), pitch = atoi(argv[3]), slots = atoi(argv[4]);
  int bw = atoi(argv[5]), bh = atoi(argv[6]), x = atoi(argv[7]), y = atoi(argv[8]), z = atoi(argv[9]);
  size_t slot_bytes = ((size_t)pitch * H + 255) & ~(size_t)255;
  std::vector<uint8_t> h(slot_bytes * slots);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 8) * 3 + 1);
  uint8_t *d, *o;
  cudaMalloc(&d, h.size());
  cudaMalloc(&o, bw * bh);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess) return 3;
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)slots};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, slot_bytes};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((EncodeTiledFn)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 4; }
  probe<<<1, 128, bw * bh + 128>>>(map, x, y, z, bw * bh, o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorName(e)); return 5; }
  std::vector<uint8_t> got(bw * bh);
  cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int j = 0; j < bh; ++j)
    for (int i = 0; i < bw; ++i) {
      int gx = x + i, gy = y + j;
      uint8_t exp = 0;
      if (gx >= 0 && gx < W && gy >= 0 && gy < H && z >= 0 && z < slots) exp = h[(size_t)z * slot_bytes + (size_t)gy * pitch + gx];
      if (got[j * bw + i] != exp) ++bad;
    }
  printf("ok, mismatches=%d\n", bad);
  return bad ? 6 : 0;
}
