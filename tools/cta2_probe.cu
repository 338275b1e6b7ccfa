// Bring-up probe for cta_group::2 ("2-SM") tcgen05 MMA -- the structural fix for the shared-memory operand bandwidth that
// bounds the BN <= 128 conv layers: a pair of CTAs on one TPC computes a 256 x N tile, each CTA stages its own
// 128 rows of A and only HALF of B.  NOT YET RUN ON A GPU (written at the end of round 1, after the GPU budget was spent);
// it is the first thing to run in round 2, before the conv kernel grows a paired variant.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/cta2_probe tools/cta2_probe.cu && tools/cta2_probe
//
// D[m][n] = sum_k A[m][k] * B[n][k],  M = 256 (2 CTAs x 128), N = 256, K = 64, bf16 in, fp32 out; A, B K-major.
//   * cluster of 2 CTAs; rank 0 is the leader and issues tcgen05.mma.cta_group::2 (instruction descriptor M = 256)
//   * both CTAs TMA-load their A rows and their half of B into their own shared memory with the .cta_group::2 load form,
//     completing bytes on the LEADER's mbarrier; the peer also arrives there remotely (barrier count 2)
//   * tcgen05.commit.cta_group::2 ... multicast::cluster signals the "done" mbarrier of both CTAs
//   * TMEM is allocated with tcgen05.alloc.cta_group::2 by warp 0 of both CTAs (same shared-memory slot offset)
//   * every CTA reads its own 128 TMEM lanes (its M half, all 256 columns)
// Sources for the instruction forms: /opt/skills/guides/blackwell_cuda_programming.md (2-CTA sections) and the CUTLASS headers
// vendored in the image (cute/arch/copy_sm100_tma.hpp SM100_TMA_2SM_LOAD_2D, cute/arch/mma_sm100_umma.hpp
// SM100_MMA_F16BF16_2x1SM_SS, cutlass/arch/barrier.h umma_arrive_multicast_2x1SM, cute/arch/tmem_allocator_sm100.hpp).
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
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void wait_parity0(uint32_t bar) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar) : "memory");
    if (!ok && clock64() - t0 > 4000000000ll) __trap();       // ~2 s: fail loudly instead of hanging the box
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);   // same offset in both CTAs (same kernel, same layout)
  uint8_t* sA = smem;                 // [128 rows][64 k] bf16, 128B swizzle
  uint8_t* sB = smem + 16384;         // this CTA's 128 of the 256 B rows
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 32768 + 64);
  const uint32_t full = smem_u32(bar), done = smem_u32(bar + 1);
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"(full) : "memory");     // leader's expect_tx arrive + peer's arrive
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(done) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync();                     // both CTAs' barriers exist before anybody signals them
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  const uint32_t full_leader = mapa(full, 0);          // the leader's barrier in the shared::cluster window
  if (threadIdx.x == 0) {
    if (rank == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(65536u) : "memory");
    // A rows [rank*128, +128), B rows [rank*128, +128): both land in THIS CTA's shared memory, bytes count on the leader
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sA)), "l"((uint64_t)&tmA), "r"(full_leader), "r"(0), "r"((int)rank * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sB)), "l"((uint64_t)&tmB), "r"(full_leader), "r"(0), "r"((int)rank * 128) : "memory");
    if (rank != 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(full_leader) : "memory");
  }
  if (rank == 0 && threadIdx.x == 0) {
    wait_parity0(full);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    auto desc = [&](uint32_t addr) {                   // K-major, SWIZZLE_128B, 8-row atoms 1 KiB apart
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
      d |= (uint64_t)(1024u >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)2 << 61;
      return d;
    };
    // D f32, A/B bf16, both K-major, N = 256, M = 256 (the pair)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = desc(smem_u32(sA)) + (uint64_t)(2 * k), db = desc(smem_u32(sB)) + (uint64_t)(2 * k);
      const uint32_t acc = k != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    // completion of the pair's MMAs -> the "done" barrier of BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(done), "h"((uint16_t)3) : "memory");
  }
  wait_parity0(done);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = (int)rank * 128 + threadIdx.x;        // TMEM lane = row of this CTA's M half
  for (int c0 = 0; c0 < 256; c0 += 32) {
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
    for (int i = 0; i < 32; ++i) D[row * 256 + c0 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();                     // nobody frees TMEM / exits while the peer still reads
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int M = 256, N = 256, K = 64;
  std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K), ref(M * N, 0.f), out(M * N);
  uint32_t s = 2024u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 9) & 0xffff) / 65536.f - 0.5f; };
  for (int i = 0; i < M * K; ++i) { hA[i] = __float2bfloat16(rnd()); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { hB[i] = __float2bfloat16(rnd()); fB[i] = __bfloat162float(hB[i]); }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float a = 0.f;
      for (int k = 0; k < K; ++k) a += fA[m * K + k] * fB[n * K + k];
      ref[m * N + n] = a;
    }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, M * N * 4));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  CUtensorMap ta, tb;
  cuuint64_t dimsA[2] = {(cuuint64_t)K, (cuuint64_t)M}, dimsB[2] = {(cuuint64_t)K, (cuuint64_t)N}, str[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
  if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
      enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("tensor map encode failed\n");
    return 2;
  }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
  probe<<<2, 128, 34 * 1024 + 1024, 0>>>(ta, tb, dD);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
  double maxe = 0;
  int bad = 0;
  for (int i = 0; i < M * N; ++i) {
    const double e = fabs((double)out[i] - ref[i]);
    if (!(e <= 1e-3 + 1e-3 * fabs(ref[i]))) ++bad;
    if (e > maxe || e != e) maxe = e;
  }
  printf("cta_group::2 probe: %d / %d wrong, max err %.4g  D[0][0..1] %.4f %.4f ref %.4f %.4f  D[128][0] %.4f ref %.4f\n", bad,
         M * N, maxe, out[0], out[1], ref[0], ref[1], out[128 * 256], ref[128 * 256]);
  return bad ? 1 : 0;
}
