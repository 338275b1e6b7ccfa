// Bring-up probe for MN-major tcgen05 operands (what a weight-gradient GEMM needs: K = pixels is the strided dimension of
// NHWC tensors, so both operands arrive "MN-major": every shared-memory row is one pixel (k) holding 64 contiguous
// channels (m or n)).   D[m][n] = sum_p A[p][m] * B[p][n],  P = 64 pixels, M = 128, N = 128, bf16 in, fp32 out.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/mn_probe tools/mn_probe.cu
//   tools/mn_probe <lbo_bytes> <sbo_bytes> <kstep_bytes>       (canonical guess: 8192 1024 2048)
//
// A and B are TMA-loaded as two [64 px][64 ch] boxes each (128B swizzle), box j at +8192 bytes.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D, uint32_t lbo, uint32_t sbo,
      uint32_t kstep) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 2 boxes x 8 KiB
  uint8_t* sB = smem + 16384;         // 2 boxes x 8 KiB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 32768 + 64);
  const uint32_t full = smem_u32(bar), done = smem_u32(bar + 1);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(done) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(32768u) : "memory");
    for (int j = 0; j < 2; ++j) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(sA + j * 8192)), "l"((uint64_t)&tmA), "r"(full), "r"(j * 64), "r"(0) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(sB + j * 8192)), "l"((uint64_t)&tmB), "r"(full), "r"(j * 64), "r"(0) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(full) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    auto desc = [&](uint32_t addr) {
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
      d |= (uint64_t)(lbo >> 4) << 16;
      d |= (uint64_t)(sbo >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)2 << 61;
      return d;
    };
    // D f32, A/B bf16, both MN-major (bits 15, 16), N = 128, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = desc(smem_u32(sA) + k * kstep), db = desc(smem_u32(sB) + k * kstep);
      const uint32_t acc = k != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(done) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = threadIdx.x;           // TMEM lane = m
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) D[row * 128 + c0 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const uint32_t lbo = argc > 1 ? atoi(argv[1]) : 8192, sbo = argc > 2 ? atoi(argv[2]) : 1024, kstep = argc > 3 ? atoi(argv[3]) : 2048;
  const int P = 64, M = 128, N = 128;
  std::vector<__nv_bfloat16> hA(P * M), hB(P * N);
  std::vector<float> fA(P * M), fB(P * N), ref(M * N, 0.f), out(M * N);
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 9) & 0xffff) / 65536.f - 0.5f; };
  for (int i = 0; i < P * M; ++i) { hA[i] = __float2bfloat16(rnd()); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < P * N; ++i) { hB[i] = __float2bfloat16(rnd()); fB[i] = __bfloat162float(hB[i]); }
  for (int p = 0; p < P; ++p)
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) ref[m * N + n] += fA[p * M + m] * fB[p * N + n];
  __nv_bfloat16 *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, P * M * 2)); CK(cudaMalloc(&dB, P * N * 2)); CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), P * M * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), P * N * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, M * N * 4));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  CUtensorMap ta, tb;
  cuuint64_t dimsA[2] = {(cuuint64_t)M, (cuuint64_t)P}, strA[1] = {(cuuint64_t)M * 2};
  cuuint64_t dimsB[2] = {(cuuint64_t)N, (cuuint64_t)P}, strB[1] = {(cuuint64_t)N * 2};
  cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
  if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, strA, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
      enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, strB, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("tensor map encode failed\n");
    return 2;
  }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
  probe<<<1, 128, 34 * 1024 + 1024, 0>>>(ta, tb, dD, lbo, sbo, kstep);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
  double maxe = 0, maxr = 0;
  int bad = 0;
  for (int i = 0; i < M * N; ++i) {
    const double e = fabs((double)out[i] - ref[i]);
    if (!(e <= 1e-3 + 1e-3 * fabs(ref[i]))) ++bad;
    if (e > maxe || e != e) maxe = e;
    if (fabs(ref[i]) > maxr) maxr = fabs(ref[i]);
  }
  printf("lbo %u sbo %u kstep %u: %d / %d wrong, max err %.4g (max |ref| %.3g)  D[0][0..3] = %.4f %.4f %.4f %.4f  ref %.4f %.4f %.4f %.4f\n",
         lbo, sbo, kstep, bad, M * N, maxe, maxr, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
  return bad ? 1 : 0;
}
