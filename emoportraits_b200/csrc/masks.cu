// Device side of the mask pre/post-processing around the EXTERNAL mask networks.  The networks
// themselves (BiSeNet: repos/face_par_off, MODNet: repos/MODNet) are separate checkouts that are not part of the reference
// tree; what the reference tree does around them is restated here as three exact-fp32 kernels:
//
//   emo_parsing_prepare  networks/volumetric_avatar/face_parcing.py:57-58   (x - mean) / std per channel, then
//                        F.interpolate(size=(512, 512), mode='bilinear') (align_corners=False, no antialias)
//   emo_parsing_masks    networks/volumetric_avatar/face_parcing.py:60-80   F.interpolate(y, size=(h, w), 'bilinear') ->
//                        argmax over the classes -> membership of the label in four label sets (mask, face_body, mask_body,
//                        mask_cloth): one pass, the up-sampled logits never exist in memory
//   emo_resize_area      notebooks/infer.py:651-657, 673, 681 (get_mask)    F.interpolate(mode='area') == adaptive average
//                        pooling, with the Normalize((0.5,)*3, (0.5,)*3) of :651-657 fused as a per-element affine applied
//                        to every input element before the average
#include "common.cuh"

namespace emo {

// F.interpolate bilinear, align_corners=False: src = (dst + 0.5) * (in / out) - 0.5, clamped at 0; the upper neighbour is
// clamped to the last index; lambda from the unclamped-at-the-top source coordinate (ATen area_pixel_compute_source_index)
__device__ __forceinline__ void bilinear_taps(int o, int in, int out, int& i0, int& i1, float& l1) {
  // two roundings, never contracted into an FMA: ATen's CPU path computes scale * (dst + 0.5) - 0.5 that way, and one ulp of the
  // source coordinate (6e-5 at index 700) times a steep image gradient is far above fp32 noise
  const float scale = __fdiv_rn((float)in, (float)out);
  float s = __fsub_rn(__fmul_rn(scale, (float)o + 0.5f), 0.5f);
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = fminf(fmaxf(s - (float)i0, 0.f), 1.f);
}

__global__ void parsing_prepare_kernel(const float* __restrict__ in, int N, int C, int Hin, int Win, int Hout, int Wout,
                                       const float* __restrict__ mean, const float* __restrict__ std, float* __restrict__ out) {
  const long long total = (long long)N * C * Hout * Wout;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx;
    const int ow = (int)(r % Wout); r /= Wout;
    const int oh = (int)(r % Hout); r /= Hout;
    const int c = (int)(r % C);
    const int n = (int)(r / C);
    const float m = mean ? __ldg(mean + c) : 0.f, s = std ? __ldg(std + c) : 1.f;
    const float* p = in + ((long long)n * C + c) * Hin * Win;
    if (Hin == Hout && Win == Wout) {  // F.interpolate to the same size is the identity
      out[idx] = (__ldg(p + (long long)oh * Win + ow) - m) / s;
      continue;
    }
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_taps(oh, Hin, Hout, y0, y1, ly);
    bilinear_taps(ow, Win, Wout, x0, x1, lx);
    const float v00 = (__ldg(p + (long long)y0 * Win + x0) - m) / s, v01 = (__ldg(p + (long long)y0 * Win + x1) - m) / s;
    const float v10 = (__ldg(p + (long long)y1 * Win + x0) - m) / s, v11 = (__ldg(p + (long long)y1 * Win + x1) - m) / s;
    // ATen upsample_bilinear2d: w00*v00 + w01*v01 + w10*v10 + w11*v11 with w = (1-ly|ly) * (1-lx|lx), summed in that order
    const float hy = 1.f - ly, hx = 1.f - lx;
    out[idx] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
  }
}

// one thread per output pixel: bilinear sample of the K class planes, running argmax (first maximum wins, as torch.argmax
// on a tie-free input), membership of the label in the four label bit sets
__global__ void parsing_masks_kernel(const float* __restrict__ logits, int N, int K, int Hin, int Win, int Hout, int Wout,
                                     unsigned set0, unsigned set1, unsigned set2, unsigned set3, unsigned char* __restrict__ out,
                                     unsigned char* __restrict__ labels) {
  const long long plane = (long long)Hout * Wout, total = (long long)N * plane;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(idx % Wout);
    const int oh = (int)((idx / Wout) % Hout);
    const int n = (int)(idx / plane);
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_taps(oh, Hin, Hout, y0, y1, ly);
    bilinear_taps(ow, Win, Wout, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const bool same = (Hin == Hout && Win == Wout);
    float best = 0.f;
    int lab = 0;
    for (int k = 0; k < K; ++k) {
      const float* p = logits + ((long long)n * K + k) * Hin * Win;
      float v;
      if (same) v = __ldg(p + (long long)oh * Win + ow);
      else
        v = hy * (hx * __ldg(p + (long long)y0 * Win + x0) + lx * __ldg(p + (long long)y0 * Win + x1)) +
            ly * (hx * __ldg(p + (long long)y1 * Win + x0) + lx * __ldg(p + (long long)y1 * Win + x1));
      if (k == 0 || v > best) { best = v; lab = k; }
    }
    const unsigned bit = 1u << lab;
    out[0 * total + idx] = (set0 & bit) ? 1 : 0;
    out[1 * total + idx] = (set1 & bit) ? 1 : 0;
    out[2 * total + idx] = (set2 & bit) ? 1 : 0;
    out[3 * total + idx] = (set3 & bit) ? 1 : 0;
    if (labels) labels[idx] = (unsigned char)lab;
  }
}

// adaptive average pooling (ATen adaptive_avg_pool2d: start = floor(o * in / out), end = ceil((o + 1) * in / out); rows outer,
// columns inner, one running fp32 sum divided by the window size), input elements mapped x -> x * scale + shift first
__global__ void resize_area_kernel(const float* __restrict__ in, int NC, int Hin, int Win, int Hout, int Wout, float scale, float shift,
                                   float* __restrict__ out) {
  const long long total = (long long)NC * Hout * Wout;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(idx % Wout);
    const int oh = (int)((idx / Wout) % Hout);
    const long long nc = idx / ((long long)Wout * Hout);
    const int y0 = (int)(((long long)oh * Hin) / Hout), y1 = (int)((((long long)oh + 1) * Hin + Hout - 1) / Hout);
    const int x0 = (int)(((long long)ow * Win) / Wout), x1 = (int)((((long long)ow + 1) * Win + Wout - 1) / Wout);
    const float* p = in + nc * Hin * Win;
    float acc = 0.f;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) acc += fmaf(__ldg(p + (long long)y * Win + x), scale, shift);
    out[idx] = acc / (float)((y1 - y0) * (x1 - x0));
  }
}

}  // namespace emo

using namespace emo;

static inline unsigned grid_for(long long total, int block, int sms_x) {
  long long b = cdivll(total, block);
  const long long cap = (long long)sms_x;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

extern "C" int emo_parsing_prepare(const float* in, int N, int C, int Hin, int Win, int Hout, int Wout, const float* mean,
                                   const float* std, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EMO_REQUIRE(in && out && N > 0 && C > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "emo_parsing_prepare: bad arguments");
  const long long total = (long long)N * C * Hout * Wout;
  launch_kernel(parsing_prepare_kernel, grid_for(total, 256, 148 * 16), 256, 0, stream, in, N, C, Hin, Win, Hout, Wout, mean, std, out);
  return check_launch("emo_parsing_prepare");
}

// label_sets: HOST array of four 32-bit class sets (bit k = class k belongs to the set)
extern "C" int emo_parsing_masks(const float* logits, int N, int K, int Hin, int Win, int Hout, int Wout, const unsigned* label_sets,
                                 unsigned char* out, unsigned char* labels, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EMO_REQUIRE(logits && label_sets && out && N > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "emo_parsing_masks: bad arguments");
  EMO_REQUIRE(K > 0 && K <= 32, "emo_parsing_masks: 1..32 classes (label sets are 32-bit masks), got %d", K);
  const long long total = (long long)N * Hout * Wout;
  launch_kernel(parsing_masks_kernel, grid_for(total, 128, 148 * 32), 128, 0, stream, logits, N, K, Hin, Win, Hout, Wout, label_sets[0],
                label_sets[1], label_sets[2], label_sets[3], out, labels);
  return check_launch("emo_parsing_masks");
}

extern "C" int emo_resize_area(const float* in, int N, int C, int Hin, int Win, int Hout, int Wout, float scale, float shift, float* out,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EMO_REQUIRE(in && out && N > 0 && C > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "emo_resize_area: bad arguments");
  const long long total = (long long)N * C * Hout * Wout;
  launch_kernel(resize_area_kernel, grid_for(total, 256, 148 * 16), 256, 0, stream, in, N * C, Hin, Win, Hout, Wout, scale, shift, out);
  return check_launch("emo_resize_area");
}
