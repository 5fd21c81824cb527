// embed.cu - the token prologue of the ViT: PatchEmbed + CLS/pos-embed + pos_drop,
// /root/reference/src/models/vit.py:25-36 (Conv2d with kernel == stride, flatten, transpose) and vit.py:207-212
// (cat CLS, add pos_embed, dropout).
//
// A 16x16/stride-16 convolution is a GEMM over non-overlapping patches, so there is nothing to "im2col": the
// patch matrix is a pure re-ordering of the image.  gvit_patchify writes it straight into the (B, 1+Np, C*P*P) layout
// of the token tensor (row 0 of every image zero = the CLS slot), so the projection GEMM and its weight gradient run
// over all B*(1+Np) rows without slicing; gvit_embed_assemble then adds bias / pos_embed, inserts the CLS token and
// applies pos_drop in ONE pass.  Both are pure HBM streams (read once, write once).
#include "kernels.cuh"
#include "philox.cuh"

namespace gvit {
namespace {

// One thread = 8 horizontally adjacent pixels of one image row: a 32-byte (fp32) / 16-byte (bf16) read that is
// contiguous with its neighbours' (a warp reads 1 KB of one image row), and one 16/32-byte write into the patch row.
// Feature order inside a patch row is (c, i, j) - the order Conv2d's weight.view(D, C*P*P) expects.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256) patchify_kernel(const Tin* __restrict__ img, int B, int C, int H, int W, int P,
                                                       Tout* __restrict__ out) {
  const int gw = W / P, gh = H / P, Np = gw * gh;
  const int K = C * P * P;
  const int w8 = W / 8;
  const int64_t total = (int64_t)B * C * H * w8;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int x8 = (int)(t % w8);
    int64_t r = t / w8;
    const int y = (int)(r % H); r /= H;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    float v[8];
    load8(img + ((((int64_t)b * C + c) * H + y) * W + x8 * 8), v);
    const int x = x8 * 8, px = x / P, j = x % P, py = y / P, i = y % P;
    Tout* dst = out + ((int64_t)b * (Np + 1) + 1 + py * gw + px) * K + (c * P + i) * P + j;
    store8(dst, v);
  }
  // CLS slot: row 0 of every image is zero
  const int k8 = K / 8;
  const float z[8] = {};
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)B * k8; t += (int64_t)gridDim.x * blockDim.x)
    store8(out + (t / k8) * (int64_t)(Np + 1) * K + (t % k8) * 8, z);
}

// x[b,0,:] = cls + pos[0];  x[b,n,:] = y[b,n,:] + bias + pos[n] (n >= 1);  then dropout(p).
// y is the projection of the patch matrix over all 1+Np rows (row 0 of y is ignored).  cls / pos / bias are the fp32
// (or T) parameters themselves: Tp.  To: the token stream's dtype - T, or fp32 over a bf16 projection (the fp32 residual
// stream torch.autocast produces: cat / add with the fp32 cls_token / pos_embed promote, vit.py:207-211).
template <typename T, typename Tp, typename To>
__global__ void __launch_bounds__(256) embed_assemble_kernel(const T* __restrict__ y, const Tp* __restrict__ bias,
                                                             const Tp* __restrict__ cls, const Tp* __restrict__ pos, int B, int N,
                                                             int D, float p, uint64_t seed, uint64_t offset,
                                                             const uint64_t* __restrict__ offset_dev, To* __restrict__ out,
                                                             uint8_t* __restrict__ mask) {
  if (offset_dev) offset += __ldg(offset_dev);
  const float scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  const uint32_t th = dropout_thresh16(p);
  const int d8 = D / 8;
  const int64_t per_img = (int64_t)N * d8;
  const int64_t total = (int64_t)B * per_img;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = t % per_img;          // (row, column group) inside the image
    const int n = (int)(q / d8), c = (int)(q % d8) * 8;
    float a[8], e[8];
    load8(pos + (int64_t)n * D + c, e);
    if (n == 0) {
      load8(cls + c, a);
    } else {
      load8(y + t * 8, a);
      if (bias) {
        float bb[8];
        load8(bias + c, bb);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += bb[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] += e[i];
    if (p > 0.f) {
      const uint32_t bits = keep_bits8(seed, offset + (uint64_t)t, th);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = (bits >> i) & 1u ? a[i] * scale : 0.f;
      mask[t] = (uint8_t)bits;
    }
    store8(out + t * 8, a);
  }
}


}  // namespace

int patchify(const void* img, int B, int C, int H, int W, int P, int in_dtype, int out_dtype, void* out, cudaStream_t st) {
  using bf = __nv_bfloat16;
  const int64_t n8 = (int64_t)B * C * H * (W / 8);
  if (in_dtype == GVIT_F32 && out_dtype == GVIT_F32)
    stream_launch(patchify_kernel<float, float>, n8, st, static_cast<const float*>(img), B, C, H, W, P, static_cast<float*>(out));
  else if (in_dtype == GVIT_F32)
    stream_launch(patchify_kernel<float, bf>, n8, st, static_cast<const float*>(img), B, C, H, W, P, static_cast<bf*>(out));
  else
    stream_launch(patchify_kernel<bf, bf>, n8, st, static_cast<const bf*>(img), B, C, H, W, P, static_cast<bf*>(out));
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

int embed_assemble(const void* y, const void* bias, const void* cls, const void* pos, int B, int N, int D, float p,
                   uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int dtype, int param_dtype, int out_dtype, void* out,
                   uint8_t* keep_mask, cudaStream_t st) {
  using bf = __nv_bfloat16;
  const int64_t n8 = (int64_t)B * N * (D / 8);
  if (dtype == GVIT_BF16 && out_dtype == GVIT_F32)
    stream_launch(embed_assemble_kernel<bf, float, float>, n8, st, static_cast<const bf*>(y), static_cast<const float*>(bias), static_cast<const float*>(cls),
                                                                  static_cast<const float*>(pos), B, N, D, p, seed, offset, offset_dev, static_cast<float*>(out), keep_mask);
  else if (dtype == GVIT_F32)
    stream_launch(embed_assemble_kernel<float, float, float>, n8, st, static_cast<const float*>(y), static_cast<const float*>(bias), static_cast<const float*>(cls),
                                                              static_cast<const float*>(pos), B, N, D, p, seed, offset, offset_dev, static_cast<float*>(out), keep_mask);
  else if (param_dtype == GVIT_F32)
    stream_launch(embed_assemble_kernel<bf, float, bf>, n8, st, static_cast<const bf*>(y), static_cast<const float*>(bias), static_cast<const float*>(cls),
                                                           static_cast<const float*>(pos), B, N, D, p, seed, offset, offset_dev, static_cast<bf*>(out), keep_mask);
  else
    stream_launch(embed_assemble_kernel<bf, bf, bf>, n8, st, static_cast<const bf*>(y), static_cast<const bf*>(bias), static_cast<const bf*>(cls),
                                                        static_cast<const bf*>(pos), B, N, D, p, seed, offset, offset_dev, static_cast<bf*>(out), keep_mask);
  GVIT_CHECK_LAUNCH();
  return GVIT_OK;
}

}  // namespace gvit
