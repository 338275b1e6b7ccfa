// Detection post-processing on the device: confidence threshold + class-aware greedy NMS.
//
// Replaces [yolox 0.3.0] yolox.utils.postprocess (called at /root/reference/exps/evaluators/onex_stream_evaluator.py:148,
// sAP/streamyolo/streamyolo_det.py:62-83): cxcywh -> xyxy, class_conf / class_pred = max over the class scores,
// keep obj * class_conf >= conf_thre, torchvision.ops.batched_nms(boxes, obj * class_conf, class, nms_thre), rows
// [x1, y1, x2, y2, obj, class_conf, class_pred] in decreasing score order.  The reference does this per image in Python
// with torchvision's NMS; here one CTA per image: composite-key bitonic sort in shared memory (score descending, anchor
// index ascending on ties: deterministic), greedy suppression with the IoU arithmetic of torchvision's kernel
// (inter / (area_i + area_j - inter) > thr, no +1), compaction.  Compiled with -fmad=false: every decision is bit-exact
// against the fp32 restatement in oracle/postprocess_oracle.py.
#include "common.cuh"

namespace sy {

constexpr int kNmsThreads = 1024;

struct NmsArgs {
  const float* pred;       // [B][A][5 + NC]
  int A, NC, Apad, max_det;
  float conf_thre, nms_thre;
  int class_agnostic;
  float* boxes;            // workspace [B][A][4] sorted xyxy
  int* cls;                // workspace [B][A]
  float* det;              // [B][max_det][7]
  int* count;              // [B]
};

__global__ void __launch_bounds__(kNmsThreads) nms_kernel(const NmsArgs q) {
  extern __shared__ unsigned long long keys[];                 // [Apad], then removed flags [Apad] bytes, then scan scratch
  unsigned char* removed = reinterpret_cast<unsigned char*>(keys + q.Apad);
  __shared__ int s_scan[kNmsThreads / 32];
  __shared__ int s_n;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int no = 5 + q.NC;
  const float* P = q.pred + (size_t)b * q.A * no;
  // ---- 1. scores -> composite keys (0 = rejected)
  for (int a = tid; a < q.Apad; a += kNmsThreads) {
    unsigned long long key = 0ull;
    if (a < q.A) {
      const float* r = P + (size_t)a * no;
      float best = r[5];
      for (int k = 1; k < q.NC; ++k) best = r[5 + k] > best ? r[5 + k] : best;      // first maximum, like torch.max
      const float score = r[4] * best;
      if (score >= q.conf_thre)
        key = ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)a);
    }
    keys[a] = key;
  }
  __syncthreads();
  // ---- 2. bitonic sort, descending
  for (int k = 2; k <= q.Apad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < q.Apad; i += kNmsThreads) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long x = keys[i], y = keys[l];
          const bool desc = (i & k) == 0;
          if (desc ? (x < y) : (x > y)) { keys[i] = y; keys[l] = x; }
        }
      }
      __syncthreads();
    }
  }
  // ---- 3. number of candidates (keys are sorted: first zero key), sorted boxes / classes
  if (tid == 0) s_n = 0;
  __syncthreads();
  for (int i = tid; i < q.Apad; i += kNmsThreads)
    if (keys[i] != 0ull && (i + 1 == q.Apad || keys[i + 1] == 0ull)) s_n = i + 1;
  __syncthreads();
  const int n = s_n;
  float* BX = q.boxes + (size_t)b * q.A * 4;
  int* CL = q.cls + (size_t)b * q.A;
  for (int j = tid; j < n; j += kNmsThreads) {
    const int a = (int)(0xFFFFFFFFu - (unsigned)(keys[j] & 0xFFFFFFFFull));
    const float* r = P + (size_t)a * no;
    const float hw = r[2] / 2.f, hh = r[3] / 2.f;
    BX[j * 4 + 0] = r[0] - hw; BX[j * 4 + 1] = r[1] - hh; BX[j * 4 + 2] = r[0] + hw; BX[j * 4 + 3] = r[1] + hh;
    int best = 0;
    float bv = r[5];
    for (int k = 1; k < q.NC; ++k)
      if (r[5 + k] > bv) { bv = r[5 + k]; best = k; }
    CL[j] = best;
    removed[j] = 0;
  }
  __syncthreads();
  // ---- 4. greedy suppression in score order, 32 candidates at a time.  (One block-wide barrier per candidate -- the first
  //         version -- cost ~0.25 ms for 2000 candidates; now two barriers per 32.)
  //   (a) warp 0 settles the chunk among its own members: lane l owns candidate c0 + l; for i = 0..31 in order, if candidate
  //       i is still alive, the later lanes of the same class test their box against it (the box of i comes by shuffle);
  //   (b) all threads apply the chunk's survivors to the candidates behind the chunk.
  //   Same decisions as the sequential loop: a candidate is removed iff a kept earlier candidate of its class overlaps it.
  __shared__ float s_kbox[32][4];
  __shared__ float s_karea[32];
  __shared__ int s_kcls[32];
  __shared__ int s_nk;
  for (int c0 = 0; c0 < n; c0 += 32) {
    if (tid < 32) {
      const int j = c0 + tid;
      const bool in = j < n;
      float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
      int cl = -1;
      bool alive = false;
      if (in) {
        x1 = BX[j * 4]; y1 = BX[j * 4 + 1]; x2 = BX[j * 4 + 2]; y2 = BX[j * 4 + 3];
        cl = CL[j];
        alive = removed[j] == 0;
      }
      const float area = (x2 - x1) * (y2 - y1);
      for (int i = 0; i < 32; ++i) {
        const unsigned live = __ballot_sync(0xffffffffu, alive);
        if (!((live >> i) & 1u)) continue;                   // candidate i was removed (or lies past n): uniform
        const float ix1 = __shfl_sync(0xffffffffu, x1, i), iy1 = __shfl_sync(0xffffffffu, y1, i);
        const float ix2 = __shfl_sync(0xffffffffu, x2, i), iy2 = __shfl_sync(0xffffffffu, y2, i);
        const float iarea = __shfl_sync(0xffffffffu, area, i);
        const int ic = __shfl_sync(0xffffffffu, cl, i);
        if (alive && tid > i && (q.class_agnostic || cl == ic)) {
          const float xx1 = fmaxf(ix1, x1), yy1 = fmaxf(iy1, y1);
          const float xx2 = fminf(ix2, x2), yy2 = fminf(iy2, y2);
          const float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
          const float inter = w * h;
          const float ovr = inter / (iarea + area - inter);
          if (ovr > q.nms_thre) alive = false;
        }
      }
      if (in && !alive) removed[j] = 1;
      // survivors of the chunk, compacted (order irrelevant for step (b))
      const unsigned live = __ballot_sync(0xffffffffu, alive);
      if (alive) {
        const int k = __popc(live & ((1u << tid) - 1u));
        s_kbox[k][0] = x1; s_kbox[k][1] = y1; s_kbox[k][2] = x2; s_kbox[k][3] = y2;
        s_karea[k] = area;
        s_kcls[k] = cl;
      }
      if (tid == 0) s_nk = __popc(live);
    }
    __syncthreads();
    const int nk = s_nk;
    if (nk > 0) {
      for (int j = c0 + 32 + tid; j < n; j += kNmsThreads) {
        if (removed[j]) continue;
        const float x1 = BX[j * 4], y1 = BX[j * 4 + 1], x2 = BX[j * 4 + 2], y2 = BX[j * 4 + 3];
        const float jarea = (x2 - x1) * (y2 - y1);
        const int cl = CL[j];
        for (int k = 0; k < nk; ++k) {
          if (!q.class_agnostic && s_kcls[k] != cl) continue;
          const float xx1 = fmaxf(s_kbox[k][0], x1), yy1 = fmaxf(s_kbox[k][1], y1);
          const float xx2 = fminf(s_kbox[k][2], x2), yy2 = fminf(s_kbox[k][3], y2);
          const float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
          const float inter = w * h;
          const float ovr = inter / (s_karea[k] + jarea - inter);
          if (ovr > q.nms_thre) { removed[j] = 1; break; }
        }
      }
    }
    __syncthreads();
  }
  // ---- 5. compaction in score order
  int base = 0;
  for (int j0 = 0; j0 < n; j0 += kNmsThreads) {
    const int j = j0 + tid;
    const int keep = (j < n && !removed[j]) ? 1 : 0;
    // block-wide exclusive scan of keep
    int v = keep;
    const int lane = tid & 31, warp = tid >> 5;
    for (int m = 1; m < 32; m <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, m);
      if (lane >= m) v += t;
    }
    if (lane == 31) s_scan[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int w = lane < kNmsThreads / 32 ? s_scan[lane] : 0;
      for (int m = 1; m < 32; m <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, m);
        if (lane >= m) w += t;
      }
      if (lane < kNmsThreads / 32) s_scan[lane] = w;
    }
    __syncthreads();
    const int pos = base + (warp ? s_scan[warp - 1] : 0) + v - keep;
    if (keep && pos < q.max_det) {
      const int a = (int)(0xFFFFFFFFu - (unsigned)(keys[j] & 0xFFFFFFFFull));
      const float* r = P + (size_t)a * no;
      float bv = r[5];
      for (int k = 1; k < q.NC; ++k) bv = r[5 + k] > bv ? r[5 + k] : bv;
      float* o = q.det + ((size_t)b * q.max_det + pos) * 7;
      o[0] = BX[j * 4]; o[1] = BX[j * 4 + 1]; o[2] = BX[j * 4 + 2]; o[3] = BX[j * 4 + 3];
      o[4] = r[4]; o[5] = bv; o[6] = (float)CL[j];
    }
    base += s_scan[kNmsThreads / 32 - 1];
    __syncthreads();
  }
  if (tid == 0) q.count[b] = base < q.max_det ? base : q.max_det;
}

static inline size_t nms_ws_bytes(int B, int A) { return ((size_t)B * A * 4 * sizeof(float) + 255) / 256 * 256 + (size_t)B * A * sizeof(int); }

}  // namespace sy

using namespace sy;

extern "C" size_t sy_postprocess_nms_workspace_bytes(int32_t b, int32_t a_total) {
  return b > 0 && a_total > 0 ? nms_ws_bytes(b, a_total) : 0;
}

extern "C" int sy_postprocess_nms(const SyNmsDesc* d, sy_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SY_REQUIRE(d != nullptr, SY_EINVAL, "null descriptor");
  SY_REQUIRE(d->b > 0 && d->a_total > 0 && d->num_classes >= 1 && d->max_det > 0, SY_EINVAL, "postprocess_nms: bad sizes");
  SY_REQUIRE(d->pred && d->det_out && d->count_out && d->workspace, SY_EINVAL, "postprocess_nms: null pointer");
  SY_REQUIRE(d->workspace_bytes >= nms_ws_bytes(d->b, d->a_total), SY_EWORKSPACE, "postprocess_nms: workspace %zu < %zu",
             d->workspace_bytes, nms_ws_bytes(d->b, d->a_total));
  SY_REQUIRE(((uintptr_t)d->workspace % 16) == 0, SY_EINVAL, "postprocess_nms: workspace must be 16B aligned");
  int apad = 32;
  while (apad < d->a_total) apad <<= 1;
  const size_t smem = (size_t)apad * 9;
  SY_REQUIRE(smem <= 200 * 1024, SY_EINVAL, "postprocess_nms: %d anchors exceed the shared-memory sort (max 16384)", d->a_total);
  if (smem > 48 * 1024) SY_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NmsArgs q{};
  q.pred = d->pred; q.A = d->a_total; q.NC = d->num_classes; q.Apad = apad; q.max_det = d->max_det;
  q.conf_thre = d->conf_thre; q.nms_thre = d->nms_thre; q.class_agnostic = d->class_agnostic;
  q.boxes = reinterpret_cast<float*>(d->workspace);
  q.cls = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(d->workspace) +
                                 ((size_t)d->b * d->a_total * 4 * sizeof(float) + 255) / 256 * 256);
  q.det = d->det_out; q.count = d->count_out;
  nms_kernel<<<d->b, kNmsThreads, smem, stream>>>(q);
  return launch_status("nms_kernel");
}
