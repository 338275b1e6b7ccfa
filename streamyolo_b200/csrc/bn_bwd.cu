// Backward of BatchNorm(train) + SiLU behind every [yolox] BaseConv (what autograd runs for loss.backward(),
// /root/reference/exps/train_utils/double_trainer.py:114):   y = silu(z),  z = gamma * xhat + beta,
// xhat = (raw - mean_g) * invstd_g  with the batch statistics of the pixel's statistics group g (current / support frames,
// see DESIGN.md section 3), raw = the conv output.
//
//   dz      = dy * silu'(z)                      silu'(z) = s (1 + z (1 - s)),  s = sigmoid(z)
//   dbeta   = sum dz             dgamma = sum dz * xhat               (over both groups)
//   draw    = gamma * invstd_g * (dz - mean_g(dz) - xhat * mean_g(dz * xhat))
//
// Three launches: per-(image, pixel chunk) partial sums (deterministic, fixed order), a finalize that produces dgamma,
// dbeta and the two per-group coefficients, and the element-wise pass that writes draw in bf16 for the conv's data /
// weight gradient kernels.  HBM-bound 16-byte accesses over NHWC bf16 views.
#include <math.h>

#include "common.cuh"

namespace sy {

constexpr int kBwdChunk = 512;     // pixels per partial row

__device__ __forceinline__ void unpack8b(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ float dsilu(float z) {
  const float s = 1.0f / (1.0f + __expf(-z));
  return s * (1.0f + z * (1.0f - s));
}

struct BwdArgs {
  const __nv_bfloat16* raw; long long raw_pitch;
  const __nv_bfloat16* dy; long long dy_pitch;
  const float* scale; const float* shift; const float* mean; const float* invstd;   // [2 groups][C]
  int HW, C, split_n, act;
};

// partial rows [n * chunks][2 (sum dz | sum dz*xhat)][C]
__global__ void bn_act_bwd_reduce_kernel(const BwdArgs q, float* partials) {
  __shared__ float red[256][17];
  const int chunks = cdiv(q.HW, kBwdChunk);
  const int n = blockIdx.x / chunks, ch = blockIdx.x % chunks;
  const int p0 = ch * kBwdChunk, p1 = min(q.HW, p0 + kBwdChunk);
  const int grp = n >= q.split_n ? 1 : 0;
  const int C = q.C, G = C / 8;
  const int lanes = G < 256 ? G : 256, PL = 256 / lanes;
  const int gl = threadIdx.x % lanes, pl = threadIdx.x / lanes;
  float* out = partials + (size_t)blockIdx.x * 2 * C;
  for (int g0 = 0; g0 < G; g0 += lanes) {
    const int g = g0 + gl;
    float s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = 0.f;
    if (g < G && pl < PL) {
      float sc[8], sh[8], mu[8], is[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = grp * C + g * 8 + i;
        sc[i] = q.scale[c]; sh[i] = q.shift[c]; mu[i] = q.mean[c]; is[i] = q.invstd[c];
      }
      for (int pp = p0 + pl; pp < p1; pp += PL) {
        const long long pix = (long long)n * q.HW + pp;
        float r[8], d[8];
        unpack8b(*reinterpret_cast<const uint4*>(q.raw + pix * q.raw_pitch + g * 8), r);
        unpack8b(*reinterpret_cast<const uint4*>(q.dy + pix * q.dy_pitch + g * 8), d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float z = r[i] * sc[i] + sh[i];
          const float dz = q.act ? d[i] * dsilu(z) : d[i];
          s[i] += dz;
          s[8 + i] += dz * ((r[i] - mu[i]) * is[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) red[threadIdx.x][i] = s[i];
    __syncthreads();
    if (pl == 0 && g < G) {
      for (int i = 0; i < 16; ++i) {
        float a = 0.f;
        for (int k = 0; k < PL; ++k) a += red[k * lanes + gl][i];
        out[(i >> 3) * C + g * 8 + (i & 7)] = a;
      }
    }
    __syncthreads();
  }
}

// rows [0, rows0) belong to group 0, [rows0, rows) to group 1.  coef [2 groups][2 (mean dz | mean dz*xhat)][C];
// dgamma / dbeta (+)= sums over both groups.  One thread per channel, fixed order, fp64.
__global__ void bn_act_bwd_finalize_kernel(const float* __restrict__ partials, int rows0, int rows, double cnt0, double cnt1, int C,
                                           float* dgamma, float* dbeta, int accumulate, float* coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  for (int r = 0; r < rows; ++r) {
    const int g = r >= rows0 ? 1 : 0;
    s[g][0] += (double)partials[(size_t)r * 2 * C + c];
    s[g][1] += (double)partials[(size_t)r * 2 * C + C + c];
  }
  coef[(0 * 2 + 0) * C + c] = (float)(s[0][0] / cnt0);
  coef[(0 * 2 + 1) * C + c] = (float)(s[0][1] / cnt0);
  coef[(1 * 2 + 0) * C + c] = cnt1 > 0.0 ? (float)(s[1][0] / cnt1) : 0.f;
  coef[(1 * 2 + 1) * C + c] = cnt1 > 0.0 ? (float)(s[1][1] / cnt1) : 0.f;
  const float db = (float)(s[0][0] + s[1][0]), dg = (float)(s[0][1] + s[1][1]);
  dbeta[c] = accumulate ? dbeta[c] + db : db;
  dgamma[c] = accumulate ? dgamma[c] + dg : dg;
}

__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(const BwdArgs q, const float* __restrict__ coef, long long npix,
                                                               __nv_bfloat16* draw, long long draw_pitch) {
  const int C = q.C, G = C / 8;
  const long long total = npix * G;
  const long long split_pix = (long long)q.split_n * q.HW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    const long long pix = idx / G;
    const int grp = pix >= split_pix ? 1 : 0;
    float r[8], d[8], o[8];
    unpack8b(*reinterpret_cast<const uint4*>(q.raw + pix * q.raw_pitch + g * 8), r);
    unpack8b(*reinterpret_cast<const uint4*>(q.dy + pix * q.dy_pitch + g * 8), d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = grp * C + g * 8 + i;
      const float z = r[i] * q.scale[c] + q.shift[c];
      const float dz = q.act ? d[i] * dsilu(z) : d[i];
      const float xh = (r[i] - q.mean[c]) * q.invstd[c];
      const float a = coef[(grp * 2 + 0) * C + g * 8 + i], b = coef[(grp * 2 + 1) * C + g * 8 + i];
      o[i] = q.scale[c] * (dz - a - xh * b);
    }
    *reinterpret_cast<uint4*>(draw + pix * draw_pitch + g * 8) =
        make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
  }
}

}  // namespace sy

using namespace sy;

extern "C" int sy_bn_act_bwd_rows(int32_t n, int32_t hw) { return n * cdiv(hw, kBwdChunk); }

extern "C" int sy_bn_act_backward(const SyBnActBwdDesc* d, sy_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SY_REQUIRE(d != nullptr, SY_EINVAL, "null descriptor");
  const SyTensor& raw = d->raw;
  const SyTensor& dy = d->dy;
  const SyTensor& dr = d->draw;
  SY_REQUIRE(view_ok(raw) && view_ok(dy) && view_ok(dr), SY_EINVAL, "bn_act_backward: bad view");
  SY_REQUIRE(dy.n == raw.n && dy.h == raw.h && dy.w == raw.w && dy.c == raw.c && dr.n == raw.n && dr.h == raw.h &&
                 dr.w == raw.w && dr.c == raw.c,
             SY_EINVAL, "bn_act_backward: shape mismatch");
  SY_REQUIRE(d->scale && d->shift && d->mean && d->invstd && d->dgamma && d->dbeta && d->partials && d->coef, SY_EINVAL,
             "bn_act_backward: null pointer");
  const int hw = raw.h * raw.w;
  const int rows = sy_bn_act_bwd_rows(raw.n, hw);
  SY_REQUIRE(d->n_partials >= rows, SY_EWORKSPACE, "bn_act_backward: %d partial rows, need %d", d->n_partials, rows);
  const int split = (d->split_n > 0 && d->split_n < raw.n) ? d->split_n : raw.n;
  BwdArgs q{};
  q.raw = reinterpret_cast<const __nv_bfloat16*>(raw.ptr); q.raw_pitch = raw.pitch;
  q.dy = reinterpret_cast<const __nv_bfloat16*>(dy.ptr); q.dy_pitch = dy.pitch;
  q.scale = d->scale; q.shift = d->shift; q.mean = d->mean; q.invstd = d->invstd;
  q.HW = hw; q.C = raw.c; q.split_n = split; q.act = d->act;
  bn_act_bwd_reduce_kernel<<<rows, 256, 0, stream>>>(q, d->partials);
  const int chunks = cdiv(hw, kBwdChunk);
  bn_act_bwd_finalize_kernel<<<cdiv(raw.c, 128), 128, 0, stream>>>(d->partials, split * chunks, rows, (double)split * hw,
                                                                   (double)(raw.n - split) * hw, raw.c, d->dgamma, d->dbeta,
                                                                   d->accumulate, d->coef);
  const long long npix = (long long)raw.n * hw;
  const long long total = npix * (raw.c / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  bn_act_bwd_apply_kernel<<<blocks, 256, 0, stream>>>(q, d->coef, npix, reinterpret_cast<__nv_bfloat16*>(dr.ptr), dr.pitch);
  return launch_status("bn_act_backward kernels");
}
