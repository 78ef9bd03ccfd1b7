// Heatmap decoding for sm_100a: arg-max + log-domain sub-pixel refinement in one pass.
//
// Replaces (paths under /root/reference): val.py:151-164 (two-stage torch.max + 150 .item()
// syncs + full D2H per frame), inference.py:22-51 get_max_preds, :75-94 my_taylor,
// :136-152 get_final, :171-186 getPrediction.
//
// HBM-bound: every heatmap byte is read exactly once with 128-bit streaming loads
// (ld.global.nc.L1::no_allocate), 4 independent loads in flight per thread; the 9 samples of
// the sub-pixel stencil are re-read from L2.  Algorithmic traffic 4*H*W B per map in,
// 16 B per map out.  Small maps (<= 64x64) use one warp per keypoint map, larger maps one
// 256-thread CTA per map so that a single frame still spreads over many SMs.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace epb {

// torch.max / np.argmax semantics: NaN is the maximum; first occurrence wins.
__device__ __forceinline__ bool gt_nanmax(float v, float bv) {
  return v > bv || (v != v && bv == bv);
}
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
  if (gt_nanmax(v, bv)) return true;
  if (gt_nanmax(bv, v)) return false;
  return i < bi;  // equal (or both NaN): lower flat index
}

// inference.py:75-94 my_taylor on the clipped map (inference.py:141): returns true and the
// refined coordinate when the step is applied.  math.log of Python floats = FP64.
__device__ __forceinline__ bool subpixel_offset(const float* __restrict__ p, int H, int W, int px, int py,
                                                double& off_x, double& off_y) {
  if (!(px > 1 && px < W - 2 && py > 1 && py < H - 2)) return false;
  const float clipv = 1e-10f;  // np.maximum(hm_f32, 1e-10) stays float32
  auto lg = [&](int yy, int xx) { return log((double)fmaxf(__ldg(p + (size_t)yy * W + xx), clipv)); };
  const double c = lg(py, px);
  const double hx = 0.5 * (lg(py, px + 1) - lg(py, px - 1));
  const double hy = 0.5 * (lg(py + 1, px) - lg(py - 1, px));
  const double hxx = 0.25 * (lg(py, px + 2) - 2 * c + lg(py, px - 2));
  const double hyy = 0.25 * (lg(py + 2, px) - 2 * c + lg(py - 2, px));
  if (hxx != 0 && hyy != 0) {
    const double ox = -hx / hxx, oy = -hy / hyy;
    if (ox < 1 && oy < 1) {             // no lower bound, inference.py:92
      off_x = ox; off_y = oy;
      return true;
    }
  }
  return false;
}

template <int THREADS, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
decode_kernel(const float* __restrict__ hm, int n_maps, int H, int W, int flags,
              float* __restrict__ xy, float* __restrict__ maxval, int32_t* __restrict__ idx_out) {
  constexpr int MAPS_PER_BLOCK = BLOCK / THREADS;
  constexpr int WARPS_PER_MAP = THREADS / 32;
  const int local_map = threadIdx.x / THREADS;
  const int t = threadIdx.x % THREADS;
  const int map = blockIdx.x * MAPS_PER_BLOCK + local_map;
  const bool live = map < n_maps;
  const int n = H * W;
  const float* p = hm + (size_t)(live ? map : 0) * n;

  float bv = -INFINITY;
  int bi = 0x7fffffff;
  if (live) {
    if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      const float4* p4 = reinterpret_cast<const float4*>(p);
      const int n4 = n >> 2;
      int i = t;
      for (; i + 3 * THREADS < n4; i += 4 * THREADS) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ldg_stream_f4(p4 + i + u * THREADS);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = (i + u * THREADS) << 2;
          if (gt_nanmax(v[u].x, bv)) { bv = v[u].x; bi = e; }
          if (gt_nanmax(v[u].y, bv)) { bv = v[u].y; bi = e + 1; }
          if (gt_nanmax(v[u].z, bv)) { bv = v[u].z; bi = e + 2; }
          if (gt_nanmax(v[u].w, bv)) { bv = v[u].w; bi = e + 3; }
        }
      }
      for (; i < n4; i += THREADS) {
        const float4 v = ldg_stream_f4(p4 + i);
        const int e = i << 2;
        if (gt_nanmax(v.x, bv)) { bv = v.x; bi = e; }
        if (gt_nanmax(v.y, bv)) { bv = v.y; bi = e + 1; }
        if (gt_nanmax(v.z, bv)) { bv = v.z; bi = e + 2; }
        if (gt_nanmax(v.w, bv)) { bv = v.w; bi = e + 3; }
      }
    } else {
      for (int i = t; i < n; i += THREADS) {
        const float v = __ldg(p + i);
        if (gt_nanmax(v, bv)) { bv = v; bi = i; }
      }
    }
  }
  // warp reduce (value, first index)
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const float ov = __shfl_xor_sync(FULL, bv, m);
    const int oi = __shfl_xor_sync(FULL, bi, m);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if (WARPS_PER_MAP > 1) {
    __shared__ float sv[BLOCK / 32];
    __shared__ int si[BLOCK / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (t < 32) {
      const int w0 = local_map * WARPS_PER_MAP;
      bv = (lane < WARPS_PER_MAP) ? sv[w0 + lane] : -INFINITY;
      bi = (lane < WARPS_PER_MAP) ? si[w0 + lane] : 0x7fffffff;
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        const float ov = __shfl_xor_sync(FULL, bv, m);
        const int oi = __shfl_xor_sync(FULL, bi, m);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
    }
  }
  if (!live || t != 0) return;
  if (bi == 0x7fffffff) bi = 0;  // all -inf: np.argmax returns 0
  const int px = bi % W, py = bi / W;
  float fx = (float)px, fy = (float)py;
  double ox, oy;
  if ((flags & EPB_DECODE_REFINE) && subpixel_offset(p, H, W, px, py, ox, oy)) {
    fx = (float)((double)px + ox);  // numpy: float32 array += float64 list -> one rounding
    fy = (float)((double)py + oy);
  }
  if ((flags & EPB_DECODE_ZERO_NONPOS) && !(bv > 0.0f)) { fx = 0.f; fy = 0.f; }
  if (xy) { xy[2 * map] = fx; xy[2 * map + 1] = fy; }
  if (maxval) maxval[map] = bv;
  if (idx_out) idx_out[map] = bi;
}

// inference.py:136-152 get_final on caller-supplied integer peaks: coords [n_maps,2] in/out.
__global__ void refine_keypoints_kernel(const float* __restrict__ hm, int n_maps, int H, int W,
                                        float* __restrict__ xy) {
  const int map = blockIdx.x * blockDim.x + threadIdx.x;
  if (map >= n_maps) return;
  float fx = xy[2 * map], fy = xy[2 * map + 1];
  const int px = (int)fx, py = (int)fy;  // int(coord[0]) truncation, inference.py:79-80
  double ox, oy;
  if (subpixel_offset(hm + (size_t)map * H * W, H, W, px, py, ox, oy)) {
    xy[2 * map] = (float)((double)fx + ox);  // coord += offset on the float32 coordinate
    xy[2 * map + 1] = (float)((double)fy + oy);
  }
}

// inference.py:154-170 get_final2 (DARK-style decode) on caller-supplied integer peaks:
//   gaussian_blur(hm, 11) (:96-110: zero-padded copy in float64, cv2.GaussianBlur(11x11, sigma from the
//   kernel size = 2.0), cast back to float32, rescaled so that the map keeps its maximum),
//   clip at 1e-10, float32 log, then taylor (:54-73): full 2x2 Hessian Newton step at the peak,
//   applied whenever the Hessian determinant is non-zero.
// One CTA per map.  The whole map has to be blurred because the rescale needs the maximum of the
// BLURRED map; it is streamed through shared memory in row bands (separable 11-tap filter, float64
// like cv2 on the float64 copy), HBM traffic = one read of the map.
struct GaussKernel11 { double k[6]; };  // k[0] centre, k[i] = k[-i]

__global__ void __launch_bounds__(256)
dark_refine_kernel(const float* __restrict__ hm, int n_maps, int H, int W, int band_rows, GaussKernel11 g,
                   float* __restrict__ xy) {
  extern __shared__ __align__(16) unsigned char dark_smem[];
  const int map = blockIdx.x;
  const float* p = hm + (size_t)map * H * W;
  const int rows = band_rows + 10;
  double* tmp = reinterpret_cast<double*>(dark_smem);                 // [rows][W] horizontally blurred
  float* in = reinterpret_cast<float*>(dark_smem + (size_t)rows * W * 8);  // [rows][W]
  __shared__ float s_st[5][5];
  __shared__ float s_red[2][8];
  const float fx0 = xy[2 * map], fy0 = xy[2 * map + 1];
  const int px = (int)fx0, py = (int)fy0;
  float omax = -INFINITY, bmax = -INFINITY;
  for (int r0 = 0; r0 < H; r0 += band_rows) {
    for (int e = threadIdx.x; e < rows * W; e += 256) {
      const int rr = e / W, c = e - rr * W;
      const int r = r0 - 5 + rr;
      float v = 0.f;
      if (r >= 0 && r < H) {
        v = __ldg(p + (size_t)r * W + c);
        if (rr >= 5 && rr < 5 + band_rows) omax = fmaxf(omax, v);
      }
      in[e] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < rows * W; e += 256) {
      const int rr = e / W, c = e - rr * W;
      const float* row = in + rr * W;
      double acc = g.k[0] * (double)row[c];
#pragma unroll
      for (int i = 1; i <= 5; ++i) {
        const double a = c - i >= 0 ? (double)row[c - i] : 0.0, b = c + i < W ? (double)row[c + i] : 0.0;
        acc += g.k[i] * (a + b);
      }
      tmp[e] = acc;
    }
    __syncthreads();
    const int nb = min(band_rows, H - r0);
    for (int e = threadIdx.x; e < nb * W; e += 256) {
      const int rr = e / W, c = e - rr * W;
      const double* col = tmp + (size_t)(rr + 5) * W + c;
      double acc = g.k[0] * col[0];
#pragma unroll
      for (int j = 1; j <= 5; ++j) acc += g.k[j] * (col[-j * W] + col[j * W]);
      const float v = (float)acc;
      bmax = fmaxf(bmax, v);
      const int r = r0 + rr;
      if (r >= py - 2 && r <= py + 2 && c >= px - 2 && c <= px + 2) s_st[r - py + 2][c - px + 2] = v;
    }
    __syncthreads();
  }
  omax = fmaxf(omax, __shfl_xor_sync(FULL, omax, 16)); bmax = fmaxf(bmax, __shfl_xor_sync(FULL, bmax, 16));
#pragma unroll
  for (int m = 8; m > 0; m >>= 1) {
    omax = fmaxf(omax, __shfl_xor_sync(FULL, omax, m)); bmax = fmaxf(bmax, __shfl_xor_sync(FULL, bmax, m));
  }
  if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = omax; s_red[1][threadIdx.x >> 5] = bmax; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int w = 0; w < 8; ++w) { omax = fmaxf(omax, s_red[0][w]); bmax = fmaxf(bmax, s_red[1][w]); }
  if (!(px > 1 && px < W - 2 && py > 1 && py < H - 2)) return;       // inference.py:59
  const float scale = __fdiv_rn(omax, bmax);                            // hm *= origin_max / np.max(hm), float32
  auto L = [&](int dy, int dx) { return logf(fmaxf(__fmul_rn(s_st[dy + 2][dx + 2], scale), 1e-10f)); };
  const float c0 = L(0, 0);
  const float dx = 0.5f * (L(0, 1) - L(0, -1));
  const float dy = 0.5f * (L(1, 0) - L(-1, 0));
  const float dxx = 0.25f * (L(0, 2) - 2.f * c0 + L(0, -2));
  const float dxy = 0.25f * (L(1, 1) - L(-1, 1) - L(1, -1) + L(-1, -1));
  const float dyy = 0.25f * (L(2, 0) - 2.f * c0 + L(-2, 0));
  const float det = __fsub_rn(__fmul_rn(dxx, dyy), __fmul_rn(dxy, dxy));
  if (det != 0.f) {
    // offset = -inv(hessian) derivative (numpy.linalg.inv on float32; evaluated here in float64)
    const double d = (double)dxx * dyy - (double)dxy * dxy;
    const float ox = (float)(-((double)dyy * dx - (double)dxy * dy) / d);
    const float oy = (float)(-((double)dxx * dy - (double)dxy * dx) / d);
    xy[2 * map] = fx0 + ox;
    xy[2 * map + 1] = fy0 + oy;
  }
}

}  // namespace epb

static void decode_kernel_attributes() {
  using namespace epb;
  // decode_kernel keeps the DEFAULT L1 / shared split: it streams the heatmaps at the HBM rate only with a large
  // L1 behind its loads (measured on C3: 0.127 ms = 97 % of the HBM peak with the default or a 0 % carve-out,
  // 0.143 ms = 87 % once 30 % or more of the array is asked for shared memory).  It belongs to the heatmap path and
  // never shares an SM with vote_count, which is what the carve-out preference of common.cuh is for.
  prefer_max_shared(refine_keypoints_kernel); prefer_max_shared(dark_refine_kernel);
  cudaGetLastError();
}

extern "C" int epb_decode_heatmaps(const float* hm, int n_maps, int H, int W, int flags, float* xy,
                                   float* maxval, int32_t* idx, void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(decode_kernel_attributes);
  using namespace epb;
  if (!hm || n_maps < 0 || H <= 0 || W <= 0 || (long long)H * W > 0x7fffffffLL) return EPB_ERR_INVALID;
  if (n_maps == 0) return EPB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope ps(PROF_DECODE, s);
  if (H * W <= 64 * 64) {
    constexpr int BLOCK = 128;
    decode_kernel<32, BLOCK><<<(n_maps + 3) / 4, BLOCK, 0, s>>>(hm, n_maps, H, W, flags, xy, maxval, idx);
  } else {
    constexpr int BLOCK = 256;
    decode_kernel<256, BLOCK><<<n_maps, BLOCK, 0, s>>>(hm, n_maps, H, W, flags, xy, maxval, idx);
  }
  return check_launch();
}

extern "C" int epb_refine_keypoints(const float* hm, int n_maps, int H, int W, float* xy, void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(decode_kernel_attributes);
  using namespace epb;
  if (!hm || !xy || n_maps < 0 || H <= 0 || W <= 0) return EPB_ERR_INVALID;
  if (n_maps == 0) return EPB_OK;
  refine_keypoints_kernel<<<(n_maps + 127) / 128, 128, 0, (cudaStream_t)stream>>>(hm, n_maps, H, W, xy);
  return check_launch();
}

extern "C" int epb_refine_keypoints_dark(const float* hm, int n_maps, int H, int W, float* xy, void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(decode_kernel_attributes);
  using namespace epb;
  if (!hm || !xy || n_maps < 0 || H <= 0 || W <= 0 || W > 4096) return EPB_ERR_INVALID;
  if (n_maps == 0) return EPB_OK;
  // cv2.getGaussianKernel(11, sigma <= 0): sigma = 0.3*((11-1)*0.5 - 1) + 0.8 = 2.0, normalised to sum 1
  GaussKernel11 g;
  {
    const double sigma = 0.3 * ((11 - 1) * 0.5 - 1) + 0.8, scale2x = -0.5 / (sigma * sigma);
    double cf[11], sum = 0;
    for (int i = 0; i < 11; ++i) { const double x = i - 5; cf[i] = exp(scale2x * x * x); sum += cf[i]; }
    for (int i = 0; i <= 5; ++i) g.k[i] = cf[5 + i] * (1.0 / sum);
  }
  int band = (int)(200 * 1024 / ((size_t)W * 12)) - 10;
  if (band > H) band = H;
  if (band > 64) band = 64;
  if (band < 1) return EPB_ERR_INVALID;
  const size_t smem = (size_t)(band + 10) * W * 12;
  // per device and cheap: set on every call (one process may drive several devices)
  EPB_RETURN_IF(check_api(cudaFuncSetAttribute(dark_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)));
  dark_refine_kernel<<<n_maps, 256, smem, (cudaStream_t)stream>>>(hm, n_maps, H, W, band, g, xy);
  return check_launch();
}
