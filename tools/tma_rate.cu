// tma_rate.cu -- how many cp.async.bulk.tensor.3d box loads per microsecond does one SM sustain
// for the small u8 boxes the HBMA window kernels use?  (design input for k_hbma_pool.cu)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tma_rate tools/tma_rate.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// every CTA: `rounds` rounds of `depth` box loads in flight (one mbarrier phase per round)
__global__ void rate(const __grid_constant__ CUtensorMap map, int W, int H, int bw, int bh, int rounds, int depth,
                     uint32_t* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
  const int slot = (bw * bh + 127) & ~127;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t parity = 0, h = blockIdx.x * 2654435761u;
  for (int r = 0; r < rounds; ++r) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr),
                   "r"((uint32_t)(bw * bh * depth)) : "memory");
      for (int d = 0; d < depth; ++d) {
        h = h * 1664525u + 1013904223u;
        const int x = (int)((h >> 8) % (uint32_t)(W - bw)) & ~15, y = (int)((h >> 20) % (uint32_t)(H - bh));
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem + d * slot);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(dst), "l"(&map), "r"(x), "r"(y), "r"(0), "r"(bar_addr) : "memory");
      }
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar_addr), "r"(parity) : "memory");
    }
    parity ^= 1u;
  }
  if (threadIdx.x == 0 && smem[0] == 0xee && smem[1] == 0x11) sink[0] = 1;
}

int main() {
  const int W = 1920, H = 1088, pitch = 1920;
  uint8_t* d;
  uint32_t* sink;
  cudaMalloc(&d, (size_t)pitch * H * 2);
  cudaMemset(d, 3, (size_t)pitch * H * 2);
  cudaMalloc(&sink, 4);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess) return 3;
  const int shapes[][2] = {{16, 16}, {48, 32}, {48, 8}, {48, 1}, {64, 48}, {96, 80}, {160, 144}, {256, 32}, {128, 64}};
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("box_w box_h depth ctas/SM | us/op/SM  ops/us/SM  GB/s(chip)  clk/row(1.9GHz)\n");
  for (auto& s : shapes) {
    const int bw = s[0], bh = s[1];
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, 2};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * H};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (((EncodeTiledFn)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 4;
    for (int depth : {1, 4}) {
      for (int cps : {1, 4}) {
        const int slot = (bw * bh + 127) & ~127;
        if (slot * depth > 200 * 1024 / cps - 1024) continue;
        const int rounds = 400, ctas = 148 * cps;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        rate<<<ctas, 32, slot * depth>>>(map, W, H, bw, bh, 20, depth, sink);
        cudaEventRecord(e0);
        rate<<<ctas, 32, slot * depth>>>(map, W, H, bw, bh, rounds, depth, sink);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("failed\n"); return 5; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops_per_sm = (double)rounds * depth * cps;
        const double us_per_op = ms * 1e3 / ops_per_sm;
        printf("%5d %5d %5d %5d   | %8.4f  %8.2f  %9.1f  %8.2f\n", bw, bh, depth, cps, us_per_op, 1.0 / us_per_op,
               (double)bw * bh * ops_per_sm * 148 / (ms * 1e-3) / 1e9, us_per_op * 1900.0 / bh);
      }
    }
  }
  return 0;
}
