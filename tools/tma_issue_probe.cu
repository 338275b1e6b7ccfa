// What does ONE cp.async.bulk.tensor (UTMALDG) cost the issuing thread, and how many bytes per clock does an SM pull through
// TMA?  The conv main loop is producer-bound (profiles/r02_timeline_*): a producer warp needs ~450 cycles per load.  This probe
// separates the pieces: W warps (one elected lane each) issue `loads` back-to-back box loads of [rows][64] bf16 (128-byte rows,
// SWIZZLE_128B) into a shared-memory ring, every load on its own mbarrier phase-0 barrier slot, with or without an
// expect_tx in front, and record clock64() around every issue; afterwards they wait for all bytes.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/tma_issue_probe tools/tma_issue_probe.cu
//   /tmp/tma_issue_probe            (prints a table; all 148 SMs run the same thing, SM 0 reports)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void wait0(uint32_t bar) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar) : "memory");
    if (!ok && clock64() - t0 > 4000000000ll) __trap();
  }
}

constexpr int kMaxLoads = 32;

// mode bit 0: expect_tx before every load (else one expect_tx for everything up front)
// mode bit 1: all loads of a warp land in the SAME shared-memory slot (no ring; shows whether smem write conflicts matter)
// mode bit 2: all SMs read the same rows
__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tm, int rows, int loads, int warps, int mode, long long* out, int row_span) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bars[4];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t box_bytes = (uint32_t)rows * 128u;
  const int slots = (200 * 1024) / (int)box_bytes / warps;          // ring slots per warp
  long long t[kMaxLoads + 2];
  if (warp < warps) {
    const uint32_t bar = smem_u32(&bars[warp]);
    if (elect_one()) {
      if (!(mode & 1)) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes * loads) : "memory");
      t[0] = clock64();
#pragma unroll 1
      for (int i = 0; i < loads; ++i) {
        const int slot = (mode & 2) ? 0 : (i % slots);
        const uint32_t dst = smem_u32(smem) + (uint32_t)((warp * slots + slot) * box_bytes);
        // every SM, warp and load reads different rows (row_span rows of the tensor per SM)
        // (mode bit 2: every SM reads the SAME rows, like the weight slabs of the conv kernel)
        const int r0 = (((mode & 4) ? 0 : (int)blockIdx.x * row_span) + (warp * loads + i) * rows) % (row_span * (int)gridDim.x);
        if (mode & 1) {
          if (i == loads - 1) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes) : "memory");
          else asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(box_bytes) : "memory");
        }
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst), "l"((uint64_t)&tm), "r"(bar), "r"(0), "r"(r0) : "memory");
        t[i + 1] = clock64();
      }
    }
    __syncwarp();
    wait0(bar);
    if (elect_one()) {
      t[loads + 1] = clock64();
      if (blockIdx.x == 0)
        for (int i = 0; i < loads + 2; ++i) out[warp * (kMaxLoads + 2) + i] = t[i];
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int sms = 148, row_span = 4096;                  // 4096 rows x 128 B = 512 KiB per SM, 74 MiB in all: L2 resident after the warm-up
  const size_t n_rows = (size_t)sms * row_span;
  __nv_bfloat16* d;
  long long* dout;
  CK(cudaMalloc(&d, n_rows * 128));
  CK(cudaMemset(d, 0, n_rows * 128));
  CK(cudaMalloc(&dout, 4 * (kMaxLoads + 2) * sizeof(long long)));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  printf("rows warps mode | issue cycles per load (median of loads 4..) | total cycles | bytes/clk/SM (all warps) | first issues\n");
  for (int rows : {64, 128, 256}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {64, (cuuint64_t)n_rows}, str[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)rows}, es[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      printf("encode failed\n");
      return 2;
    }
    for (int warps : {1, 2}) {
      for (int mode : {0, 2, 6}) {
        const int loads = rows == 256 ? 8 : 16;              // <= 512 KiB per SM and launch
        std::vector<long long> h(4 * (kMaxLoads + 2));
        for (int rep = 0; rep < 3; ++rep) {              // rep 0-1 warm the L2 and the descriptor cache
          probe<<<sms, 128, 210 * 1024>>>(tm, rows, loads, warps, mode, dout, row_span);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h.data(), dout, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        std::vector<long long> iss;
        for (int i = 4; i < loads; ++i) iss.push_back(h[i + 1] - h[i]);
        std::sort(iss.begin(), iss.end());
        long long total = 0;
        for (int w = 0; w < warps; ++w) {
          const long long tt = h[w * (kMaxLoads + 2) + loads + 1] - h[w * (kMaxLoads + 2)];
          if (tt > total) total = tt;
        }
        printf("%4d %5d %4d | %6lld | %7lld | %6.1f | %lld %lld %lld %lld\n", rows, warps, mode, iss[iss.size() / 2], total,
               (double)rows * 128 * loads * warps / (double)total, h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3]);
      }
    }
  }
  return 0;
}
