// PVNet-style RANSAC keypoint voting for sm_100a, batched and stream-ordered.
//
// Replaces (paths under /root/reference/lib/ransac_voting_gpu_layer):
//   src/ransac_voting_kernel.cu:11-49   generate_hypothesis_kernel
//   src/ransac_voting_kernel.cu:88-126  voting_for_hypothesis_kernel  (+ the hn*vn*tn byte tensor
//                                        and torch.sum that follow it, ransac_voting_gpu.py:557-561)
//   ransac_voting_gpu.py:514-598 v3, :669-761 v4, :763-858 v5, :218-261 hypothesis,
//   :263-331 / :333-406 voting distribution (the per-image Python loops, torch.nonzero /
//   masked_select compaction, uniform_/random_ draws, torch.max / topk / matmul / solve tails).
//
// Pipeline of one epb_voting_run() call (no host synchronisation anywhere):
//   mask_count -> mask_scan -> rng_offsets -> [mask_count/scan with the max_num subsample]
//   -> mask_scatter (stable row-major compaction of foreground pixel coordinates)
//   -> hypothesis (seeded pairs of foreground pixels -> ray intersections)
//   -> vote_count (dominant: hn*vn*tn inlier tests per image, never materialised)
//   -> winner_refine | distribution.
//
// Bit-exactness: every float operation of the reference kernels is issued through explicit
// round-to-nearest intrinsics in the order the reference SASS executes them, so
// nvcc cannot re-contract them.  vote_count decides each (hypothesis, pixel) pair with two affine
// forms of the hypothesis (no division, no square root) and a band whose width is derived in
// DESIGN.md section 5; only the pairs inside the band (~4e-4 of them at T = 0.999) are re-evaluated
// with the reference expression itself, so the counts equal the reference's integers.
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace epb {

// ------------------------------------------------------------------------------------------
// exact reference arithmetic
// ------------------------------------------------------------------------------------------

// ransac_voting_kernel.cu:22-48 as nvcc 12.9 / ptxas contracts the reference source for sm_100a
// (SASS of oracle/_ref/libref_voting.so; bitwise-checked on the GPU by tests/test_voting_gpu.py):
//   det1 = rn(nx1*ny0) - rn(nx0*ny1), det2 = -det1, guards in double,
//   s = fma(nx,cx,rn(ny*cy)), y = fma(nx1,s0,-rn(nx0*s1))/det1, x = fma(-ny0,s1,rn(ny1*s0))/det2.
// The cubin the reference shipped (CUDA 11.1, sm_86) contracts the x numerator the other way round
// (fma(ny1,s0,-rn(ny0*s1))): the two builds of the SAME source differ by 1 ulp in x on some pairs.
// We follow the sm_100a build, i.e. what the reference yields when built on the machine we replace.
__device__ __forceinline__ bool intersect_rays(float dx0, float dy0, float cx0, float cy0,
                                               float dx1, float dy1, float cx1, float cy1,
                                               float* x, float* y) {
  const float nx0 = dy0, ny0 = -dx0, nx1 = dy1, ny1 = -dx1;
  const float a = __fmul_rn(nx1, ny0);
  const float b = __fmul_rn(nx0, ny1);
  const float det1 = __fsub_rn(a, b);
  const float det2 = __fsub_rn(b, a);
  if (fabs((double)det1) < 1e-6) return false;
  if (fabs((double)det2) < 1e-6) return false;
  const float s0 = __fmaf_rn(nx0, cx0, __fmul_rn(ny0, cy0));
  const float s1 = __fmaf_rn(nx1, cx1, __fmul_rn(ny1, cy1));
  *y = __fdiv_rn(__fmaf_rn(nx1, s0, -__fmul_rn(nx0, s1)), det1);
  *x = __fdiv_rn(__fmaf_rn(-ny0, s1, __fmul_rn(ny1, s0)), det2);
  return true;
}

// ransac_voting_kernel.cu:100-125 as compiled (PTX pins the fma placement).
__device__ __forceinline__ bool vote_exact(float cx, float cy, float nx, float ny, float norm1,
                                           float hx, float hy, float thresh) {
  const float dx = __fsub_rn(hx, cx);
  const float dy = __fsub_rn(hy, cy);
  const float norm2 = __fsqrt_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
  if ((double)norm1 < 1e-6 || (double)norm2 < 1e-6) return false;
  const float ad = __fdiv_rn(__fmaf_rn(dx, nx, __fmul_rn(dy, ny)), __fmul_rn(norm1, norm2));
  return ad > thresh;
}
__device__ __forceinline__ float dir_norm(float nx, float ny) {
  return __fsqrt_rn(__fmaf_rn(nx, nx, __fmul_rn(ny, ny)));
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 in the layout of torch's CUDA distribution kernels
// (ATen/native/cuda/DistributionTemplates.h: block 256, grid = min(SMs * maxThreads/256,
//  ceil(numel/256)), unroll 4, curand_init(seed, thread, offset)).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
struct TorchRngLayout {
  unsigned total_threads;  // 256 * grid
  unsigned long long inc;  // generator offset consumed by one call
};
__host__ __device__ inline TorchRngLayout torch_rng_layout(unsigned long long numel, int sm_count,
                                                           int threads_per_sm) {
  unsigned long long grid = (numel + 255) / 256;
  const unsigned long long cap = (unsigned long long)sm_count * (unsigned)(threads_per_sm / 256);
  if (grid > cap) grid = cap;
  if (grid == 0) grid = 1;
  TorchRngLayout l;
  l.total_threads = (unsigned)(256 * grid);
  l.inc = numel ? ((numel - 1) / (256ull * grid * 4ull) + 1ull) * 4ull : 0ull;
  return l;
}
// raw 32-bit draw that element `li` of a torch CUDA tensor filled by one RNG call receives
__device__ __forceinline__ unsigned torch_philox_u32(unsigned long long seed, unsigned long long offset,
                                                     unsigned long long li, unsigned total_threads) {
  const unsigned long long per_iter = 4ull * total_threads;
  const unsigned long long k = li / per_iter;
  const unsigned rem = (unsigned)(li - k * per_iter);
  const unsigned ii = rem / total_threads;
  const unsigned idx = rem - ii * total_threads;
  const unsigned long long ctr = offset / 4ull + k;
  const uint4 c = make_uint4((unsigned)ctr, (unsigned)(ctr >> 32), idx, 0u);
  const uint4 r = philox4x32_10(c, make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  return ii == 0 ? r.x : ii == 1 ? r.y : ii == 2 ? r.z : r.w;
}
// torch uniform_(0,1) on float32: curand_uniform (0,1] then 1.0 -> 0.0
__device__ __forceinline__ float torch_uniform01(unsigned r) {
  const float u = __fmaf_rn((float)r, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  return u == 1.0f ? 0.0f : u;
}

// ------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------
constexpr int TILE_PX = 4096;  // 256 threads x 16 mask bytes (8 warps x 512 pixels)

struct Workspace {
  int32_t* tile_counts;   // [B][T]
  int32_t* tile_offsets;  // [B][T]
  int32_t* fg;            // [B] foreground before the subsample
  int32_t* tn;            // [B] after
  int32_t* sub;           // [B] 1 if the max_num subsample applies
  int32_t* live;          // [B] fg >= min_num
  float* ratio;           // [B] max_num / fg (float32 like the reference)
  unsigned long long* off_u;  // [B] generator offset of the uniform_ draw
  unsigned long long* off_r;  // [B] generator offset of the first random_ draw
  int32_t* ovf;           // [B] 1 if the subsample kept more pixels than the workspace holds (image dropped)
  uint32_t* fgpix;        // [B][cap] packed (y << 16 | x), row-major stable order
  float2* direct;         // [B][vn][cap] field vectors of the foreground pixels, in fgpix order
  int cap;                // voting pixels per image the workspace holds (multiple of the vote tile, see voting_cap)
  float* hyp;             // [B][vn][2][HNs] x plane then y plane, HNs = HN rounded up to 4 (hyp_stride)
  int32_t* counts;        // [B][vn][HN]
  int32_t* item_off;      // [B+1] exclusive scan of ceil(tn / item_px): work list of vote_count
  size_t bytes;
};
__host__ __device__ inline int hyp_stride(int HN) { return (HN + 3) & ~3; }
__device__ __forceinline__ float* hyp_plane(const Workspace& ws, int b, int vn, int v, int HN, int c) {
  return ws.hyp + (((size_t)b * vn + v) * 2 + c) * hyp_stride(HN);
}
__host__ inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }
// Voting pixels per image the workspace is sized for.  Without the subsample an image votes with all its
// foreground (<= H*W).  With it (torch-compatible Philox draws; ransac_voting_gpu.py:536-540) every
// foreground pixel is kept with probability max_num / fg, so tn ~ Binomial(fg, max_num / fg): mean max_num,
// variance < max_num.  max_num + 8 sigma + 1024 is exceeded with probability < 1e-15 per image; an image
// that does exceed it is dropped with status 3 (mask_scan_kernel), never written out of bounds.
// Caller-supplied draws (`selection`, parity tests) can select anything: full size there.
__host__ inline int voting_cap(const epb_voting_params& p) {
  const long long HW = (long long)p.H * p.W;
  long long cap = HW;
  if (p.rng_mode == EPB_RNG_PHILOX && p.max_num > 0 && (long long)p.max_num < HW) {
    const long long bound = (long long)p.max_num + (long long)(8.0 * sqrt((double)p.max_num)) + 1024;
    if (bound < cap) cap = bound;
  }
  return (int)((cap + 255) / 256 * 256);   // whole vote tiles (TMA copies never leave a row)
}
__host__ inline Workspace carve(const epb_voting_params& p, void* base) {
  Workspace w;
  const size_t B = p.B, T = ((size_t)p.H * p.W + TILE_PX - 1) / TILE_PX;
  const size_t HN = (size_t)p.hn * p.rounds;
  char* c = (char*)base;
  size_t o = 0;
  auto take = [&](size_t n) { char* r = c + o; o += align_up(n); return (void*)r; };
  w.tile_counts = (int32_t*)take(B * T * 4);
  w.tile_offsets = (int32_t*)take(B * T * 4);
  w.fg = (int32_t*)take(B * 4);
  w.tn = (int32_t*)take(B * 4);
  w.sub = (int32_t*)take(B * 4);
  w.live = (int32_t*)take(B * 4);
  w.ratio = (float*)take(B * 4);
  w.off_u = (unsigned long long*)take(B * 8);
  w.off_r = (unsigned long long*)take(B * 8);
  w.ovf = (int32_t*)take(B * 4);
  w.cap = voting_cap(p);
  w.fgpix = (uint32_t*)take(B * (size_t)w.cap * 4);
  w.direct = (float2*)take(B * p.vn * (size_t)w.cap * 8);
  w.hyp = (float*)take(B * p.vn * 2 * (size_t)hyp_stride((int)HN) * 4);
  w.counts = (int32_t*)take(B * p.vn * HN * 4);
  w.item_off = (int32_t*)take((B + 1) * 4);
  w.bytes = o;
  return w;
}

// Multi-class drivers (v1 / v2) run every (image, class) pair as one "virtual image": virtual index
// bv = image * classes + (class - 1); mask and field are addressed with bv / classes, everything in the
// workspace and the outputs with bv.
__device__ __forceinline__ bool mask_pred(unsigned m, int mask_mode, unsigned cls) {
  return mask_mode == EPB_MASK_CLASS ? (m == cls) : mask_mode == EPB_MASK_EQ1 ? (m == 1u) : (m != 0u);
}

// foreground predicate of pixel `p` of image `b` including the optional max_num subsample
// (ransac_voting_gpu.py:536-540: selection.uniform_(0,1) < max_num / fg.float())
struct SubsampleCtx {
  const float* selection;  // optional user draws [B][H*W]
  unsigned long long seed;
  unsigned total_threads;
  int use_philox;
};
__device__ __forceinline__ bool keep_pixel(const SubsampleCtx& sc, int b, size_t hw, size_t p,
                                           float ratio, unsigned long long off_u) {
  float u;
  if (sc.selection) u = __ldg(sc.selection + (size_t)b * hw + p);
  else if (sc.use_philox) u = torch_uniform01(torch_philox_u32(sc.seed, off_u, p, sc.total_threads));
  else return true;
  return u < ratio;
}

// ------------------------------------------------------------------------------------------
// 1. mask_count: per-tile foreground counts.  phase 0: plain mask; phase 1: with subsample,
//    only for images flagged in ws.sub.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mask_count_kernel(const uint8_t* __restrict__ mask, int HW, int T, int mask_mode, int classes, int phase,
                  Workspace ws, SubsampleCtx sc) {
  const int b = blockIdx.y, tile = blockIdx.x;
  if (phase == 1 && !ws.sub[b]) return;
  const uint8_t* m = mask + (size_t)(b / classes) * HW;
  const unsigned cls = (unsigned)(b % classes) + 1u;
  const int p0 = tile * TILE_PX + threadIdx.x * 16;
  int cnt = 0;
  float ratio = 0.f;
  unsigned long long off_u = 0;
  if (phase == 1) { ratio = ws.ratio[b]; off_u = ws.off_u[b]; }
  if (p0 < HW) {
    unsigned char bytes[16];
    if (p0 + 16 <= HW && ((reinterpret_cast<uintptr_t>(m + p0) & 15) == 0)) {
      *reinterpret_cast<uint4*>(bytes) = __ldg(reinterpret_cast<const uint4*>(m + p0));
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) bytes[j] = (p0 + j < HW) ? m[p0 + j] : 0;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      bool f = (p0 + j < HW) && mask_pred(bytes[j], mask_mode, cls);
      if (f && phase == 1) f = keep_pixel(sc, b, HW, p0 + j, ratio, off_u);
      cnt += f;
    }
  }
  cnt = warp_sum_i(cnt);
  __shared__ int s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i];
    ws.tile_counts[(size_t)b * T + tile] = t;
  }
}

// ------------------------------------------------------------------------------------------
// 2. mask_scan: one CTA per image, exclusive scan of the tile counts.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mask_scan_kernel(int T, int phase, int min_num, int max_num, int allow_sub, Workspace ws) {
  const int b = blockIdx.x;
  if (phase == 1 && !ws.sub[b]) return;
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int32_t* cnt = ws.tile_counts + (size_t)b * T;
  int32_t* off = ws.tile_offsets + (size_t)b * T;
  for (int base = 0; base < T; base += 256) {
    const int i = base + threadIdx.x;
    const int v = i < T ? cnt[i] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(FULL, incl, d);
      if ((threadIdx.x & 31) >= d) incl += o;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += s_warp[w];
    const int carry = s_carry;
    if (i < T) off[i] = carry + wbase + incl - v;
    __syncthreads();
    if (threadIdx.x == 255) s_carry = carry + wbase + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int total = s_carry;
    if (phase == 0) {
      ws.fg[b] = total;
      const int live = total >= min_num;
      const int sub = live && allow_sub && total > max_num;
      ws.live[b] = live;
      ws.sub[b] = sub;
      ws.ratio[b] = sub ? __fdiv_rn((float)max_num, (float)total) : 2.0f;
      ws.tn[b] = live ? total : 0;
      ws.ovf[b] = 0;
    } else if (total > ws.cap) {
      // the subsample kept more than the workspace holds (see voting_cap): drop the image, flagged.  The
      // generator offsets were assigned from `live` before this pass, so the other images are unaffected.
      ws.tn[b] = 0; ws.live[b] = 0; ws.ovf[b] = 1;
    } else {
      ws.tn[b] = total;
    }
  }
}

// ------------------------------------------------------------------------------------------
// 3. rng_offsets: sequential (batch-order) generator offsets, exactly the order in which the
//    reference's per-image loop consumes torch's generator: [uniform_ if fg > max_num],
//    then `rounds` x random_.
// ------------------------------------------------------------------------------------------
__global__ void rng_offsets_kernel(int B, int rounds, unsigned long long base, unsigned long long inc_u,
                                   unsigned long long inc_r, Workspace ws,
                                   unsigned long long* consumed, unsigned long long* state) {
  // one warp: exclusive scan over the batch of what each image consumes, 32 images per step
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  if (state) base = *state;   // offset chained on the device from the previous run (batch chunks)
  unsigned long long o = base;
  for (int b0 = 0; b0 < B; b0 += 32) {
    const int b = b0 + lane;
    const bool in = b < B;
    const unsigned long long du = (in && ws.sub[b]) ? inc_u : 0ull;
    const unsigned long long dr = (in && ws.live[b]) ? inc_r * (unsigned long long)rounds : 0ull;
    unsigned long long incl = du + dr;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long up = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl += up;
    }
    const unsigned long long before = o + incl - (du + dr);
    if (in) { ws.off_u[b] = before; ws.off_r[b] = before + du; }
    o += __shfl_sync(FULL, incl, 31);
  }
  if (lane == 0) {
    if (consumed) *consumed = o - base;
    if (state) *state = o;
  }
}

// ------------------------------------------------------------------------------------------
// 4. mask_scatter: stable compaction of foreground pixel coordinates.
// ------------------------------------------------------------------------------------------
// Pixel ownership inside a tile of 4096: warp w owns the 512 pixels [512 w, 512 w + 512) as 16 groups of
// 32 consecutive pixels, lane l owns pixel 32 g + l of every group g.  Every warp-level access is then
// one contiguous run: 32 mask bytes, one aligned 128-byte line of a field plane, and -- after the
// ballot-based compaction -- one contiguous run of the output.  Whole lines are the only access shape that
// keeps in-place PCIe reads of a host-resident field at full-size requests (tools/micro/zerocopy_bw.cu:
// 51 GB/s contiguous vs 12-30 GB/s with sector-sized pieces), and contiguous float2 stores keep the L2 write
// path at one transaction per sector.  Output order is the row-major rank of the pixel (torch.nonzero).
template <bool PLANAR>
__global__ void __launch_bounds__(256, 3)
mask_scatter_kernel(const uint8_t* __restrict__ mask, int H, int W, int T, int mask_mode,
                    Workspace ws, SubsampleCtx sc, const float* __restrict__ vertex, epb_voting_params p, int light) {
  __shared__ int s_warp[8];
  // grid-stride over (image, tile): a full grid for a device-resident field; for a host-resident field
  // the launcher caps the grid at one CTA per SM -- PCIe needs ~100 KB in flight, not the machine, and
  // a small resident footprint lets the voting kernel of the previous batch chunk keep its SMs
  for (int work = blockIdx.x; work < T * p.B; work += gridDim.x) {
  const int b = work / T, tile = work - b * T;
  __syncthreads();   // s_warp of the previous trip has been consumed
  if (!ws.live[b]) continue;
  const int HW = H * W;
  const uint8_t* m = mask + (size_t)(b / p.classes) * HW;
  const unsigned cls = (unsigned)(b % p.classes) + 1u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = tile * TILE_PX + warp * 512;
  const bool sub = ws.sub[b] != 0;
  const float ratio = ws.ratio[b];
  const unsigned long long off_u = ws.off_u[b];
  unsigned mine = 0;      // bit g: this lane's pixel of group g is foreground
  int pos[16];               // rank of the lane's pixel of group g inside the warp's segment
  int wtotal = 0;
#pragma unroll
  for (int g = 0; g < 16; ++g) {
    const int px = seg + g * 32 + lane;
    bool f = false;
    if (px < HW) {
      f = mask_pred(__ldg(m + px), mask_mode, cls);
      if (f && sub) f = keep_pixel(sc, b, HW, px, ratio, off_u);
    }
    const unsigned bal = __ballot_sync(FULL, f);
    mine |= (unsigned)f << g;
    pos[g] = wtotal + __popc(bal & ((1u << lane) - 1u));
    wtotal += __popc(bal);
  }
  if (lane == 0) s_warp[warp] = wtotal;
  __syncthreads();
  int wbase = ws.tile_offsets[(size_t)b * T + tile];
  for (int w = 0; w < warp; ++w) wbase += s_warp[w];
  // gridDim.y CTAs share one tile: each gathers a slice of the keypoints, the first also writes fgpix
  const int v_per = (p.vn + gridDim.y - 1) / gridDim.y;
  const int v_lo = blockIdx.y * v_per, v_hi = min(p.vn, v_lo + v_per);
  if (blockIdx.y == 0) {
    uint32_t* out = ws.fgpix + (size_t)b * ws.cap;
#pragma unroll
    for (int g = 0; g < 16; ++g)
      if ((mine >> g) & 1u) {
        const int px = seg + g * 32 + lane;
        const int y = px / W, x = px - y * W;
        out[wbase + pos[g]] = ((uint32_t)y << 16) | (uint32_t)x;
      }
  }
  if (PLANAR && __any_sync(FULL, mine != 0u)) {
    // planar field: every (image, keypoint, component) plane is one contiguous H*W array (the NCHW network
    // output): pixel px is element px of the plane.  Groups without foreground are not read at all.
    const float* base = vertex + (b / p.classes) * p.sb + seg + lane;
    float2* dst = ws.direct + (size_t)b * p.vn * ws.cap + wbase;
    // host-resident field (`light`): 8 loads in flight per thread instead of 32 -- plenty for PCIe
    for (int v = v_lo; v < v_hi; ++v) {
      const float* px = base + (long long)v * p.sv;
      const float* py = px + p.sc;
      float2* d = dst + (size_t)v * ws.cap;
      if (!light) {
        float fx[16], fy[16];
#pragma unroll
        for (int g = 0; g < 16; ++g)
          if ((mine >> g) & 1u) { fx[g] = ldg_stream(px + g * 32); fy[g] = ldg_stream(py + g * 32); }
#pragma unroll
        for (int g = 0; g < 16; ++g)
          if ((mine >> g) & 1u) d[pos[g]] = make_float2(fx[g], fy[g]);
      } else {
#pragma unroll
        for (int g0 = 0; g0 < 16; g0 += 4) {
          float fx[4], fy[4];
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if ((mine >> (g0 + g)) & 1u) { fx[g] = ldg_stream(px + (g0 + g) * 32); fy[g] = ldg_stream(py + (g0 + g) * 32); }
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if ((mine >> (g0 + g)) & 1u) d[pos[g0 + g]] = make_float2(fx[g], fy[g]);
        }
      }
    }
  }
  }
}

// ------------------------------------------------------------------------------------------
// 4b. field_gather: the vector field of the foreground pixels, compacted in fgpix order, one
//     float2 per (image, keypoint, pixel).  This is the only kernel that touches `vertex`, it
//     reads each needed element exactly once, and nothing else of the field is ever read -- so
//     `vertex` may just as well be page-locked HOST memory mapped into the device address space
//     (zero-copy over PCIe: only the foreground fraction of the field crosses the bus).
//     The reference materialises the same array with masked_select (ransac_voting_gpu.py:544-545).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
field_gather_kernel(const float* __restrict__ vertex, epb_voting_params p, Workspace ws) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (!ws.live[b] || t >= ws.tn[b]) return;
  const size_t HW = (size_t)ws.cap;
  const uint32_t q = ws.fgpix[(size_t)b * HW + t];
  const float* src = vertex + (b / p.classes) * p.sb + (long long)(q >> 16) * p.sy + (long long)(q & 0xffff) * p.sx;
  float2* dst = ws.direct + (size_t)b * p.vn * HW + t;
  int v = 0;
  for (; v + 4 <= p.vn; v += 4) {  // 8 independent loads in flight per thread (PCIe / HBM latency)
    float2 d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* s = src + (long long)(v + j) * p.sv;
      d[j].x = ldg_stream(s); d[j].y = ldg_stream(s + p.sc);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[(size_t)(v + j) * HW] = d[j];
  }
  for (; v < p.vn; ++v) {
    const float* s = src + (long long)v * p.sv;
    dst[(size_t)v * HW] = make_float2(ldg_stream(s), ldg_stream(s + p.sc));
  }
}

// ------------------------------------------------------------------------------------------
// 5. hypothesis generation (fused batched form of generate_hypothesis_kernel + idxs draw)
// ------------------------------------------------------------------------------------------
struct HypCtx {
  const int32_t* idxs;  // IDXS / RAW32
  int rng_mode;
  unsigned long long seed;
  unsigned total_threads_r;
  unsigned long long inc_r;
};
__global__ void __launch_bounds__(256)
hypothesis_kernel(epb_voting_params p, Workspace ws, HypCtx hc,
                  float* __restrict__ hyp_out) {
  const int HN = p.hn * p.rounds;
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long per_img = (long long)HN * p.vn;
  if (gid >= per_img * p.B) return;
  const int b = (int)(gid / per_img);
  const int rem = (int)(gid - (long long)b * per_img);
  const int hh = rem / p.vn;  // hypothesis index over all rounds
  const int v = rem - hh * p.vn;
  const int r = hh / p.hn, h = hh - r * p.hn;
  float x = 0.f, y = 0.f;
  const int tn = ws.tn[b];
  if (ws.live[b] && tn > 0) {
    const size_t e = ((size_t)h * p.vn + v) * 2;  // element index inside one [hn,vn,2] draw
    unsigned t0, t1;
    if (hc.rng_mode == EPB_RNG_PHILOX) {
      const unsigned long long off = ws.off_r[b] + hc.inc_r * (unsigned long long)r;
      t0 = torch_philox_u32(hc.seed, off, e, hc.total_threads_r) % (unsigned)tn;
      t1 = torch_philox_u32(hc.seed, off, e + 1, hc.total_threads_r) % (unsigned)tn;
    } else {
      const int32_t* q = hc.idxs + (((size_t)b * p.rounds + r) * p.hn) * p.vn * 2 + e;
      t0 = (unsigned)q[0];
      t1 = (unsigned)q[1];
      if (hc.rng_mode == EPB_RNG_RAW32) { t0 %= (unsigned)tn; t1 %= (unsigned)tn; }
      else { t0 = min(t0, (unsigned)tn - 1); t1 = min(t1, (unsigned)tn - 1); }
    }
    const uint32_t* fp = ws.fgpix + (size_t)b * ws.cap;
    const uint32_t q0 = fp[t0], q1 = fp[t1];
    const int x0 = q0 & 0xffff, y0 = q0 >> 16, x1 = q1 & 0xffff, y1 = q1 >> 16;
    const float2* dir = ws.direct + ((size_t)b * p.vn + v) * ws.cap;
    const float2 d0 = dir[t0], d1 = dir[t1];
    float ix, iy;
    if (intersect_rays(d0.x, d0.y, (float)x0, (float)y0, d1.x, d1.y, (float)x1, (float)y1, &ix, &iy)) {
      x = ix; y = iy;
    }
  }
  hyp_plane(ws, b, p.vn, v, HN, 0)[hh] = x;
  hyp_plane(ws, b, p.vn, v, HN, 1)[hh] = y;
  if (hyp_out) {
    float2* o = reinterpret_cast<float2*>(hyp_out) + ((size_t)b * HN + hh) * p.vn + v;
    *o = make_float2(x, y);
  }
}

// ------------------------------------------------------------------------------------------
// 6. vote_count: the dominant kernel (hn*vn*tn inlier tests per image).
//
// One CTA of 128 threads per work unit of a device-side list of equal-sized units
//   unit = (item of <= item_px foreground pixels of one image, keypoint, chunk of 128*R hypotheses);
// (short CTAs instead of persistent ones so that a higher-priority stream -- the zero-copy gather of
// the next batch chunk -- can slip in between them);
// each thread keeps R hypotheses and their counters in registers (lanes <-> hypotheses), pixels are
// staged through shared memory as records and broadcast to all lanes.
//
// The reference decides  c = dot/(|n| |d|) > T  with IEEE sqrt and div (~35 issue slots).  With
// theta = acos(T), k = tan(theta), d0 = h - c and the pixel direction scaled by a power of two so that
// max(|nx|,|ny|) is in [1,2), the exact decision function is
//   F0 = k (d0 . n) - |n x d0| = |d0||n| sin(theta - phi) / cos(theta)      (phi = angle(d0, n)),
// two AFFINE forms of the hypothesis.  Record = their coefficients:
//   a' = A1 hx + A2 hy + A3   (A = k n, A3 = -k n.c)          p = B1 hx + B2 hy + B3   (B = (-ny, nx), B3 = ny cx - nx cy)
//   m = a' - |p|,   w = gamma a' + EH[h]
//   |m| > w : the sign of m IS the reference's decision (error bound in DESIGN.md section 5:
//             the reference's c is within 9.04 ulp-units of the exact cosine; gamma and EH cover that
//             plus every rounding of the forms above, and the |d| < 1e-6 guard)
//   else    : the warp evaluates the reference expression itself (IEEE sqrt / div), rare.
// 6 FP32 lane-operations (5 of them as packed FFMA2 over two hypotheses) + 1 FSETP + 1 LEA.HI per test:
// ~7.1 issue slots instead of ~35; counts are exact integers.
// ------------------------------------------------------------------------------------------
// Blackwell packed FP32 (FADD2 / FMUL2 / FFMA2): two hypotheses per issue slot, each half with the
// same IEEE round-to-nearest result as the scalar instruction.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ float half2(f32x2 v, int hi) {
  float a, b; upk2(v, a, b); return hi ? b : a;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}

constexpr int VOTE_THREADS = 128;
constexpr int VOTE_TILE = 256;  // records per smem tile

// constants of the fast test, computed on the host in double (vote_consts())
struct VoteConsts {
  int fast_ok;      // 0: every pair takes the exact path (threshold outside [0.5, 1) or band too wide)
  float kf;         // tan(acos(T))
  float gamma;      // relative half-width of the undecided band, in units of a'
  float eh_scale;   // EH[h] = eh_scale * (|hx| + |hy|) + eh_abs
  float eh_abs;
  float s1_min;     // (vote_mma.cuh) smallest |n|^2 whose IEEE root passes the reference's guard (double)norm1 >= 1e-6
};
__host__ inline VoteConsts vote_consts(float thresh, int H, int W) {
  VoteConsts c;
  const double T = (double)thresh, u = 5.9604644775390625e-08;  // 2^-24
  c.fast_ok = 0; c.kf = 0.f; c.gamma = 0.f; c.eh_scale = 0.f; c.eh_abs = 0.f;
  if (!(T >= 0.5 && T < 1.0)) return c;
  const double k = sqrt((1.0 - T) * (1.0 + T)) / T;
  // |cos(phi) - T| >= sin(0.75 theta) sin|theta - phi| whenever |theta - phi| <= theta/2 (inside the cone) or
  // phi > theta; deeper inside the cone the margin is >= 2 sin(theta/2) sin(theta/4), checked here once
  const double theta = acos(T);
  if (!(2.0 * sin(0.5 * theta) * sin(0.25 * theta) >= 10.5 * u)) return c;
  const double beta = 10.5 * u / (T * sin(0.75 * theta));
  const double gamma = beta * (1.0 / k + 1.0) * 1.0001;
  if (!(gamma <= 0.25)) return c;
  // eps_a + eps_p <= (6.2 k + 4.1) u Nm (|h|_1 + |c|_1) with Nm < 2; factor 3 of slack
  const double es = 3.0 * (6.2 * k + 4.1) * u * 2.0;
  const double ea = es * (double)(H + W) + 3.5e-6 * (1.0 + k);
  c.fast_ok = 1;
  c.kf = (float)k;
  c.gamma = nextafterf((float)gamma, INFINITY);
  c.eh_scale = nextafterf((float)es, INFINITY);
  c.eh_abs = nextafterf((float)ea, INFINITY);
  return c;
}

constexpr int VOTE_QCAP = 1024;  // deferred undecided pairs per work unit (overflow is resolved in line)
constexpr int VOTE_RAW_STAGES = 1;   // consumed in the middle of the previous tile, refilled after its barrier
struct VoteSmem {
  float4 f[2][VOTE_TILE];  // A1 A2 B1 B2
  float2 g[2][VOTE_TILE];  // A3 B3
  // raw tiles (field vector + packed pixel of VOTE_TILE voting pixels) as the TMA delivers them
  alignas(128) float2 raw_dir[VOTE_RAW_STAGES][VOTE_TILE];
  alignas(128) uint32_t raw_pix[VOTE_RAW_STAGES][VOTE_TILE];
  unsigned long long raw_full[VOTE_RAW_STAGES];
  unsigned q[VOTE_QCAP];   // undecided (pixel, hypothesis) pairs: t_rel << 11 | thread << 4 | r << 1 | sign(m)
  unsigned qn;
};

template <int R>
__global__ void __launch_bounds__(VOTE_THREADS, R <= 4 ? 8 : 4)
vote_count_kernel(epb_voting_params p, Workspace ws, VoteConsts vc,
                  int item_px, int chunks) {
  __shared__ VoteSmem sm;
  static_assert(R == 2 || R == 4 || R == 8, "hypotheses are processed as packed pairs");
  constexpr int P = R / 2;
  constexpr int RECS = R >= 8 ? 1 : (R == 4 ? 2 : 4);   // records per trip of the inner loop (8 tests per lane)
  constexpr int PER_THREAD = VOTE_TILE / VOTE_THREADS;
  static_assert(VOTE_TILE % VOTE_THREADS == 0 && VOTE_TILE % 4 == 0, "tile shape");
  const int HN = p.hn * p.rounds;
  const int B = p.B;
  const int* __restrict__ item_off = ws.item_off;
  const long long total = (long long)item_off[B] * p.vn * chunks;
  const float T = p.inlier_thresh;
  const f32x2 GG = pk2(vc.gamma, vc.gamma);

  {
    const long long unit = blockIdx.x;
    if (unit >= total) return;
    const int per_item = p.vn * chunks;
    const int item = (int)(unit / per_item);
    const int rem = (int)(unit - (long long)item * per_item);
    const int v = rem / chunks, chunk = rem - v * chunks;
    int lo = 0, hi = B;  // largest b with item_off[b] <= item
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (item_off[mid] <= item) lo = mid; else hi = mid;
    }
    const int b = lo;
    const int tn = ws.tn[b];
    const int t_begin = (item - item_off[b]) * item_px, t_end = min(tn, t_begin + item_px);

    // thread t owns the hypothesis pairs {2t, 2t+1} + 256 q of its chunk: one 64-bit load per plane
    // yields a packed operand directly
    f32x2 HX[P], HY[P], EH[P];
    unsigned neg[R];   // pixels that do NOT vote for the hypothesis (sign bit of m)
    const float* hypx = hyp_plane(ws, b, p.vn, v, HN, 0);
    const float* hypy = hyp_plane(ws, b, p.vn, v, HN, 1);
#pragma unroll
    for (int q = 0; q < P; ++q) {
      const int h = chunk * VOTE_THREADS * R + q * 2 * VOTE_THREADS + 2 * threadIdx.x;
      float x0 = 0.f, x1 = 0.f, y0 = 0.f, y1 = 0.f;
      if (h + 1 < HN) {
        const float2 xx = *reinterpret_cast<const float2*>(hypx + h);
        const float2 yy = *reinterpret_cast<const float2*>(hypy + h);
        x0 = xx.x; x1 = xx.y; y0 = yy.x; y1 = yy.y;
      } else if (h < HN) {
        x0 = hypx[h]; y0 = hypy[h];
      }
      HX[q] = pk2(x0, x1); HY[q] = pk2(y0, y1);
      // non-finite or huge hypotheses: undecided against every pixel -> reference expression
      const float h0 = __fadd_ru(fabsf(x0), fabsf(y0)), h1 = __fadd_ru(fabsf(x1), fabsf(y1));
      EH[q] = pk2((h0 <= 1e15f) ? __fmaf_ru(vc.eh_scale, h0, vc.eh_abs) : INFINITY,
                  (h1 <= 1e15f) ? __fmaf_ru(vc.eh_scale, h1, vc.eh_abs) : INFINITY);
      neg[2 * q] = neg[2 * q + 1] = 0u;
    }

    const uint32_t* fp = ws.fgpix + (size_t)b * ws.cap;
    const float2* dir = ws.direct + ((size_t)b * p.vn + v) * ws.cap;
    // Tiles 1.. arrive by TMA: thread 0 issues two bulk copies per tile (cp.async.bulk global -> shared, completion
    // counted in bytes on an mbarrier), always whole tiles -- the workspace rows are padded to VOTE_TILE pixels, the
    // records beyond t_end are marked as padding by index.  No register staging: a tile is in flight for a whole
    // tile of compute (~11 k cycles) before anybody looks at it.  Tile 0 is loaded directly (a TMA round trip at the
    // start of every 25 us work unit cost 3 % of the kernel).
    const int ntiles = (t_end - t_begin + VOTE_TILE - 1) / VOTE_TILE;
    auto tma_issue = [&](int i) {   // thread 0; tile i >= 1 goes to raw stage (i - 1) % VOTE_RAW_STAGES
      const int rs = (i - 1) % VOTE_RAW_STAGES;
      mbar_expect_tx(&sm.raw_full[rs], VOTE_TILE * 12);
      bulk_g2s(sm.raw_dir[rs], dir + t_begin + (size_t)i * VOTE_TILE, VOTE_TILE * 8, &sm.raw_full[rs]);
      bulk_g2s(sm.raw_pix[rs], fp + t_begin + (size_t)i * VOTE_TILE, VOTE_TILE * 4, &sm.raw_full[rs]);
    };

    // records of tile i -> record buffer `buf` (tile 0 from global memory, the others from their TMA raw stage)
    auto stage = [&](int i, int buf) {
      const int rs = (i - 1) % VOTE_RAW_STAGES;
      if (i > 0) mbar_wait(&sm.raw_full[rs], (unsigned)((i - 1) / VOTE_RAW_STAGES) & 1u);
#pragma unroll
      for (int k = 0; k < PER_THREAD; ++k) {
        const int slot_in = k * VOTE_THREADS + threadIdx.x;
        const int t = t_begin + i * VOTE_TILE + slot_in;
        uint32_t q = 0xffffffffu;
        float2 dn = make_float2(0.f, 0.f);
        if (i == 0) { if (t < t_end) { q = __ldg(fp + t); dn = __ldg(dir + t); } }
        else { if (t < t_end) q = sm.raw_pix[rs][slot_in]; dn = sm.raw_dir[rs][slot_in]; }
        const float cx = (float)(q & 0xffff), cy = (float)(q >> 16);
        const float nx = dn.x, ny = dn.y;
        const float s1 = __fmaf_rn(nx, nx, __fmul_rn(ny, ny));
        const float n1 = __fsqrt_rn(s1);
        // reference guard :121 (zero direction), NaN / overflowed |n|^2 (c = NaN or 0): never an inlier
        const bool valid = q != 0xffffffffu && !((double)n1 < 1e-6) && (s1 <= FLT_MAX);
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 g = make_float2(-INFINITY, 0.f);
        if (valid) {
          if (vc.fast_ok) {
            // exact power-of-two scaling: max(|nx|,|ny|) -> [1,2)
            const float nm = fmaxf(fabsf(nx), fabsf(ny));
            const float sc = __uint_as_float((254u - ((__float_as_uint(nm) >> 23) & 0xffu)) << 23);
            const float nxs = __fmul_rn(nx, sc), nys = __fmul_rn(ny, sc);
            f.x = __fmul_rn(vc.kf, nxs);
            f.y = __fmul_rn(vc.kf, nys);
            f.z = -nys;
            f.w = nxs;
            g.x = -__fmul_rn(vc.kf, __fmaf_rn(nxs, cx, __fmul_rn(nys, cy)));
            g.y = __fmaf_rn(nys, cx, -__fmul_rn(nxs, cy));
          } else {
            g.x = __int_as_float(0x7fc00000);  // NaN: every pair undecided
          }
        }
        const int slot = k * VOTE_THREADS + threadIdx.x;
        sm.f[buf][slot] = f; sm.g[buf][slot] = g;
      }
    };

    // one record against the thread's R hypotheses: m (sign = decision), w (band); returns "some pair undecided"
    auto test_record = [&](const float4 f, const float2 g, float (&m)[R], float (&w)[R]) -> bool {
      const f32x2 A1 = pk2(f.x, f.x), A2 = pk2(f.y, f.y), B1 = pk2(f.z, f.z), B2 = pk2(f.w, f.w);
      const f32x2 A3 = pk2(g.x, g.x), B3 = pk2(g.y, g.y);
      bool amb = false;
#pragma unroll
      for (int q = 0; q < P; ++q) {
        const f32x2 AP = fma2(A1, HX[q], fma2(A2, HY[q], A3));
        const f32x2 PP = fma2(B1, HX[q], fma2(B2, HY[q], B3));
        const f32x2 WW = fma2(GG, AP, EH[q]);
        float a0, a1, p0, p1;
        upk2(AP, a0, a1); upk2(PP, p0, p1); upk2(WW, w[2 * q], w[2 * q + 1]);
        m[2 * q] = __fsub_rn(a0, fabsf(p0));
        m[2 * q + 1] = __fsub_rn(a1, fabsf(p1));
        amb |= !(fabsf(m[2 * q]) > w[2 * q]) || !(fabsf(m[2 * q + 1]) > w[2 * q + 1]);
      }
      return amb;
    };
    // the reference expression (IEEE sqrt / div) for pixel t of this image against hypothesis (hx, hy)
    auto exact_vote = [&](int t, float hx, float hy) -> bool {
      if (t >= t_end) return false;                         // padding record
      const uint32_t q = __ldg(fp + t);
      const float2 d = __ldg(dir + t);
      return vote_exact((float)(q & 0xffff), (float)(q >> 16), d.x, d.y, dir_norm(d.x, d.y), hx, hy, T);
    };
    // Undecided pairs are rare (~6e-4 of the pairs) and scattered over the lanes: resolving them in line
    // would run the ~75-instruction reference expression with one or two active lanes.  They are queued
    // instead (with the sign the fast loop counted) and resolved densely, one pair per lane, after the
    // tile loop; the owner's count is corrected with an atomic.  A full queue falls back to in-line.
    unsigned* const qn_ptr = &sm.qn;
    unsigned* const q_ptr = sm.q;
    auto defer = [&](int t, int r, float& m, float hx, float hy) {
      const unsigned pos = atomicAdd(qn_ptr, 1u);
      if (pos < (unsigned)VOTE_QCAP)
        q_ptr[pos] = ((unsigned)(t - t_begin) << 11) | (threadIdx.x << 4) | ((unsigned)r << 1) | (__float_as_uint(m) >> 31);
      else
        m = exact_vote(t, hx, hy) ? 1.0f : -1.0f;
    };

    int buf = 0;
    if (threadIdx.x == 0) {
      sm.qn = 0u;
#pragma unroll
      for (int i = 0; i < VOTE_RAW_STAGES; ++i) mbar_init(&sm.raw_full[i], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      for (int i = 1; i <= min(VOTE_RAW_STAGES, ntiles - 1); ++i) tma_issue(i);
    }
    stage(0, 0);
    __syncthreads();
    for (int t0 = t_begin, ti = 0; t0 < t_end; t0 += VOTE_TILE, ++ti) {
      const bool more = t0 + VOTE_TILE < t_end;
      const int nrec = min(VOTE_TILE, t_end - t0);
      const float4* fr = sm.f[buf];
      const float2* gr = sm.g[buf];
      // RECS records per trip share one branch to the side path; an odd tail record is a padding
      // record of the tile (never an inlier) or, in a full tile, handled by the even tile size
      const int ntrip = (nrec + RECS - 1) / RECS;
      const float4* frp = fr;
      const float2* grp = gr;
      // The records of the NEXT tile are built in the middle of this one (its raw stage landed a whole tile ago): the
      // mbarrier check, the raw loads and the record arithmetic of a warp then run under the trips of the other
      // warps instead of in front of the end-of-tile barrier, where the whole CTA would wait for them.
      auto trips = [&](int it_begin, int it_end) {
#pragma unroll 1
      for (int it = it_begin; it < it_end; ++it, frp += RECS, grp += RECS) {
        const int i = it * RECS;
        float m[RECS][R], w[RECS][R];
        bool ambj[RECS], amb = false;
#pragma unroll
        for (int j = 0; j < RECS; ++j) { ambj[j] = test_record(frp[j], grp[j], m[j], w[j]); amb |= ambj[j]; }
        if (amb) {
#pragma unroll
          for (int j = 0; j < RECS; ++j)
            if (ambj[j]) {
#pragma unroll
              for (int r = 0; r < R; ++r)
                if (!(fabsf(m[j][r]) > w[j][r]))
                  defer(t0 + i + j, r, m[j][r], half2(HX[r >> 1], r & 1), half2(HY[r >> 1], r & 1));
            }
        }
#pragma unroll
        for (int j = 0; j < RECS; ++j)
#pragma unroll
          for (int r = 0; r < R; ++r) neg[r] += __float_as_uint(m[j][r]) >> 31;
      }
      };
      trips(0, ntrip / 2);
      if (more) stage(ti + 1, buf ^ 1);
      trips(ntrip / 2, ntrip);
      __syncthreads();
      // everybody has staged tile ti + 1: its raw stage takes the tile VOTE_RAW_STAGES further on
      if (threadIdx.x == 0 && more && ti + 1 + VOTE_RAW_STAGES < ntiles) tma_issue(ti + 1 + VOTE_RAW_STAGES);
      buf ^= 1;
    }

    int32_t* out = ws.counts + ((size_t)b * p.vn + v) * HN;
    // deferred pairs: one per lane, reference expression, correction of the count the fast loop produced
    {
      const unsigned nq = min(sm.qn, (unsigned)VOTE_QCAP);   // every enqueue precedes the loop's last barrier
      for (unsigned e = threadIdx.x; e < nq; e += VOTE_THREADS) {
        const unsigned ent = sm.q[e];
        const int t = t_begin + (int)(ent >> 11), owner = (ent >> 4) & 127, r = (ent >> 1) & 7;
        const int h = chunk * VOTE_THREADS * R + (r >> 1) * 2 * VOTE_THREADS + 2 * owner + (r & 1);
        if (h >= HN) continue;
        const bool fast_inlier = (ent & 1u) == 0u;
        const bool exact = exact_vote(t, __ldg(hypx + h), __ldg(hypy + h));
        if (exact != fast_inlier) atomicAdd(out + h, exact ? 1 : -1);
      }
    }
    // records visited by the trips above (tiles are rounded up to a multiple of RECS)
    int visited = 0;
    for (int t0 = t_begin; t0 < t_end; t0 += VOTE_TILE)
      visited += (min(VOTE_TILE, t_end - t0) + RECS - 1) / RECS * RECS;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int h = chunk * VOTE_THREADS * R + (r >> 1) * 2 * VOTE_THREADS + 2 * threadIdx.x + (r & 1);
      // every record visited beyond t_end is a padding record (m = -inf), so neg over-counts by exactly
      // the padding and the difference below is the number of voting pixels
      const int c = visited - (int)neg[r];
      if (h < HN && c != 0) atomicAdd(out + h, c);
    }
  }
}

// Thresholds the band test cannot serve (outside [0.5, 1), or so close to 1 that the band is wider than a
// quarter of a'): the reference expression for every pair, one thread per hypothesis, pixels staged through
// shared memory.  ~35 issue slots per pair; never on the default thresholds.
__global__ void __launch_bounds__(256)
vote_exact_kernel(epb_voting_params p, Workspace ws) {
  const int HN = p.hn * p.rounds;
  const int b = blockIdx.z, v = blockIdx.y, h = blockIdx.x * 256 + threadIdx.x;
  if (!ws.live[b]) return;
  const int tn = ws.tn[b];
  __shared__ float4 s_px[256];   // cx, cy, nx, ny
  __shared__ float s_n1[256];
  const uint32_t* fp = ws.fgpix + (size_t)b * ws.cap;
  const float2* dir = ws.direct + ((size_t)b * p.vn + v) * ws.cap;
  float hx = 0.f, hy = 0.f;
  if (h < HN) { hx = hyp_plane(ws, b, p.vn, v, HN, 0)[h]; hy = hyp_plane(ws, b, p.vn, v, HN, 1)[h]; }
  int cnt = 0;
  for (int t0 = 0; t0 < tn; t0 += 256) {
    const int t = t0 + threadIdx.x;
    if (t < tn) {
      const uint32_t q = __ldg(fp + t);
      const float2 d = __ldg(dir + t);
      s_px[threadIdx.x] = make_float4((float)(q & 0xffff), (float)(q >> 16), d.x, d.y);
      s_n1[threadIdx.x] = dir_norm(d.x, d.y);
    }
    __syncthreads();
    const int n = min(256, tn - t0);
    for (int r = 0; r < n; ++r) {
      const float4 px = s_px[r];
      cnt += vote_exact(px.x, px.y, px.z, px.w, s_n1[r], hx, hy, p.inlier_thresh);
    }
    __syncthreads();
  }
  if (h < HN) ws.counts[((size_t)b * p.vn + v) * HN + h] = cnt;
}

// work list of vote_count: item_off[b] = number of pixel items of the images before b.
__global__ void __launch_bounds__(256)
vote_items_kernel(int B, int item_px, Workspace ws) {
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 256) {
    const int i = base + threadIdx.x;
    const int v = (i < B && ws.live[i]) ? (ws.tn[i] + item_px - 1) / item_px : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(FULL, incl, d);
      if ((threadIdx.x & 31) >= d) incl += o;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += s_warp[w];
    const int carry = s_carry;
    if (i < B) ws.item_off[i] = carry + wbase + incl - v;
    __syncthreads();
    if (threadIdx.x == 255) s_carry = carry + wbase + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) ws.item_off[B] = s_carry;
}

#ifdef EPB_TUNING
#include "vote_mma.cuh"   // 6b. the tensor-core form of the vote test: measured, not adopted
#endif

// counts_ws [B][vn][HN] -> user layout [B][HN][vn]
__global__ void counts_export_kernel(epb_voting_params p, Workspace ws, int32_t* __restrict__ out) {
  const int HN = p.hn * p.rounds;
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long n = (long long)p.B * HN * p.vn;
  if (gid >= n) return;
  const int v = (int)(gid % p.vn);
  const long long r = gid / p.vn;
  const int h = (int)(r % HN);
  const int b = (int)(r / HN);
  // degenerate images: the reference emits ones (ransac_voting_gpu.py:231)
  out[gid] = ws.live[b] ? ws.counts[((size_t)b * p.vn + v) * HN + h] : 1;
}

// ------------------------------------------------------------------------------------------
// block reductions
// ------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void block_sum_d(double (&v)[N], double* smem /* [N*8] */) {
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < N; ++i) smem[i * 8 + warp] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += smem[i * 8 + w];
    v[i] = s;
  }
}

// ------------------------------------------------------------------------------------------
// 7. winner selection + least-squares refinement (+ variance / confidence)
//    ransac_voting_gpu.py:561-569 (first max, ratio = count/tn, strict '<' update from zeros)
//    and :578-595 (re-vote the winner, 2x2 normal equations).  Sums in FP64.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 5)
winner_refine_kernel(epb_voting_params p, Workspace ws,
                     float* __restrict__ pts, float* __restrict__ var_or_conf,
                     int32_t* __restrict__ status) {
  const int v = blockIdx.x, b = blockIdx.y;
  const int HN = p.hn * p.rounds;
  float2* out = reinterpret_cast<float2*>(pts) + (size_t)b * p.vn + v;
  float* aux = var_or_conf ? var_or_conf + (size_t)b * p.vn + v : nullptr;
  int32_t* st = status ? status + (size_t)b * p.vn + v : nullptr;
  if (!ws.live[b]) {
    if (threadIdx.x == 0) {
      const bool ovf = ws.ovf[b] != 0;   // dropped by the workspace bound (voting_cap): NaN, status 3
      *out = ovf ? make_float2(NAN, NAN) : make_float2(0.f, 0.f);
      if (aux) *aux = ovf ? NAN : (p.mode == EPB_VOTE_V4 ? 1.0f : 0.0f);
      if (st) *st = ovf ? 3 : 1;
    }
    return;
  }
  __shared__ int s_cnt[8], s_idx[8];
  __shared__ double s_red[7 * 8];
  __shared__ float2 s_pt;
  const int tn = ws.tn[b];
  // first maximum of the counts
  const int32_t* cnt = ws.counts + ((size_t)b * p.vn + v) * HN;
  int bc = -1, bi = 0x7fffffff;
  for (int h = threadIdx.x; h < HN; h += 256) {
    const int c = cnt[h];
    if (c > bc) { bc = c; bi = h; }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const int oc = __shfl_xor_sync(FULL, bc, m), oi = __shfl_xor_sync(FULL, bi, m);
    if (oc > bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_cnt[threadIdx.x >> 5] = bc; s_idx[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_cnt[w] > bc || (s_cnt[w] == bc && s_idx[w] < bi)) { bc = s_cnt[w]; bi = s_idx[w]; }
    const float ratio = __fdiv_rn((float)bc, (float)tn);
    // all_win_ratio starts at 0 and is replaced only where 0 < ratio (:566-569)
    s_pt = (0.0f < ratio) ? make_float2(hyp_plane(ws, b, p.vn, v, HN, 0)[bi], hyp_plane(ws, b, p.vn, v, HN, 1)[bi])
                          : make_float2(0.f, 0.f);
  }
  __syncthreads();
  float2 win = s_pt;
  const uint32_t* fp = ws.fgpix + (size_t)b * ws.cap;
  const float2* dir = ws.direct + ((size_t)b * p.vn + v) * ws.cap;
  if (p.mode == EPB_VOTE_V1) {   // ransac_voting_layer (:10-97): the winning hypothesis itself
    if (threadIdx.x == 0) { *out = win; if (st) *st = 0; }
    return;
  }
  // inliers of the current point and the normal equations; v2 repeats this refine_iters times (:171-200)
  const int passes = p.mode == EPB_VOTE_V2 ? max(p.refine_iters, 0) : 1;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};  // a00 a01 a11 b0 b1 sum(bb^2) count
  bool singular = false;
  double px = win.x, py = win.y;
  for (int pass = 0; pass < passes; ++pass) {
#pragma unroll
    for (int i = 0; i < 7; ++i) acc[i] = 0;
    // four pixels per trip with their loads issued up front: the loop is bound by the latency of the
    // (L2-resident) fgpix / direct loads, not by arithmetic
    for (int t0 = threadIdx.x; t0 < tn; t0 += 4 * 256) {
      uint32_t q4[4]; float2 d4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int t = t0 + j * 256;
        q4[j] = t < tn ? __ldg(fp + t) : 0u;
        d4[j] = t < tn ? __ldg(dir + t) : make_float2(0.f, 0.f);   // zero direction: never an inlier
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t q = q4[j];
        const float cx = (float)(q & 0xffff), cy = (float)(q >> 16);
        const float dx = d4[j].x, dy = d4[j].y;
        if (vote_exact(cx, cy, dx, dy, dir_norm(dx, dy), win.x, win.y, p.inlier_thresh)) {
          const double nx = dy, ny = -(double)dx;  // normal = (d_y, -d_x), :580-581
          const double bb = nx * cx + ny * cy;
          acc[0] += nx * nx; acc[1] += nx * ny; acc[2] += ny * ny;
          acc[3] += nx * bb; acc[4] += ny * bb; acc[5] += bb * bb; acc[6] += 1.0;
        }
      }
    }
    block_sum_d<7>(acc, s_red);
    const double det = acc[0] * acc[2] - acc[1] * acc[1];
    const double tr = acc[0] + acc[2];
    if (p.mode == EPB_VOTE_V2) {
      // torch.pinverse semantics (:200): no inliers -> zeros (:193-195); rank-deficient normals ->
      // minimum-norm least squares x = A^T b / ||A||_F^2; else the normal equations
      singular = false;
      if (acc[6] == 0.0 || !(tr > 0.0)) { px = 0.0; py = 0.0; }
      else if (!(det > 1e-30 * tr * tr)) { px = acc[3] / tr; py = acc[4] / tr; }
      else { px = (acc[2] * acc[3] - acc[1] * acc[4]) / det; py = (acc[0] * acc[4] - acc[1] * acc[3]) / det; }
      win = make_float2((float)px, (float)py);
    } else {
      singular = !(fabs(det) > 0.0) || !isfinite(det);
      px = NAN; py = NAN;
      if (!singular) {
        px = (acc[2] * acc[3] - acc[1] * acc[4]) / det;
        py = (acc[0] * acc[4] - acc[1] * acc[3]) / det;
      }
    }
    __syncthreads();   // s_red is reused by the next pass
  }
  const float fxp = (float)px, fyp = (float)py;
  if (threadIdx.x == 0) {
    *out = make_float2(fxp, fyp);
    if (st) *st = singular ? 2 : 0;
    if (p.mode == EPB_VOTE_V4 && aux) {
      // var = sum(residual^2) / sum(inlier), residual = n.pt - b   (:752-753)
      const double rss = px * (acc[0] * px + acc[1] * py) + py * (acc[1] * px + acc[2] * py) -
                         2.0 * (px * acc[3] + py * acc[4]) + acc[5];
      *aux = (float)(fmax(rss, 0.0) / acc[6]);
    }
  }
  if (p.mode == EPB_VOTE_V5 && aux) {
    // confidence = share of pixels voting for the refined point at the literal 0.999 (:848-850)
    int c = 0;
    for (int t0 = threadIdx.x; t0 < tn; t0 += 4 * 256) {
      uint32_t q4[4]; float2 d4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int t = t0 + j * 256;
        q4[j] = t < tn ? __ldg(fp + t) : 0u;
        d4[j] = t < tn ? __ldg(dir + t) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        c += vote_exact((float)(q4[j] & 0xffff), (float)(q4[j] >> 16), d4[j].x, d4[j].y, dir_norm(d4[j].x, d4[j].y),
                        fxp, fyp, 0.999f);
    }
    double cc[1] = {(double)c};
    block_sum_d<1>(cc, s_red);
    if (threadIdx.x == 0) *aux = __fdiv_rn((float)(int)cc[0], (float)tn);
  }
}

// ------------------------------------------------------------------------------------------
// 8. voting distribution: top-k by inlier ratio, weighted mean / covariance
//    (ransac_voting_gpu.py:318-329; _with_mean :392-402).  The k-th largest count is found with
//    a 4-pass radix select over shared-memory histograms; among hypotheses tied at the k-th
//    value the lowest indices are kept (torch.topk leaves this unspecified).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
distribution_kernel(epb_voting_params p, Workspace ws, const float* __restrict__ mean_in,
                    float* __restrict__ mean_out, float* __restrict__ cov_out) {
  const int v = blockIdx.x, b = blockIdx.y;
  const int HN = p.hn * p.rounds;
  float* mo = mean_out + ((size_t)b * p.vn + v) * 2;
  float* co = cov_out + ((size_t)b * p.vn + v) * 4;
  const bool with_mean = p.mode == EPB_VOTE_DISTRIBUTION_WITH_MEAN;
  if (!ws.live[b]) {
    // hypotheses 0, ratios 1 (:273-278): mean 0 (or the given mean), cov = mean mean^T * sum/(sum[+1e-3])
    if (threadIdx.x == 0 && ws.ovf[b]) {
      mo[0] = mo[1] = NAN; co[0] = co[1] = co[2] = co[3] = NAN;
    } else if (threadIdx.x == 0) {
      if (with_mean) {
        const double mx = mean_in[((size_t)b * p.vn + v) * 2], my = mean_in[((size_t)b * p.vn + v) * 2 + 1];
        const double s = (double)HN, den = s + 1e-3;
        mo[0] = (float)mx; mo[1] = (float)my;
        co[0] = (float)(mx * mx * s / den); co[1] = co[2] = (float)(mx * my * s / den);
        co[3] = (float)(my * my * s / den);
      } else {
        mo[0] = mo[1] = 0.f; co[0] = co[1] = co[2] = co[3] = 0.f;
      }
    }
    return;
  }
  __shared__ int hist[256];
  __shared__ int s_sel[3];
  __shared__ double s_red[6 * 8];
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  const int32_t* cnt = ws.counts + ((size_t)b * p.vn + v) * HN;
  const float* hypx = hyp_plane(ws, b, p.vn, v, HN, 0);
  const float* hypy = hyp_plane(ws, b, p.vn, v, HN, 1);
  const float tnf = (float)ws.tn[b];

  unsigned thr_val = 0;  // k-th largest count
  int need_eq = 0;       // how many of the ties at thr_val to keep
  float ratio_floor = 0.f;
  if (!with_mean) {
    const int k = min(p.topk, HN);
    unsigned prefix = 0, prefix_mask = 0;
    int remaining = k;
    for (int pass = 3; pass >= 0; --pass) {
      hist[threadIdx.x] = 0;
      __syncthreads();
      const int shift = pass * 8;
      for (int h = threadIdx.x; h < HN; h += 256) {
        const unsigned c = (unsigned)cnt[h];
        if ((c & prefix_mask) == prefix) atomicAdd(&hist[(c >> shift) & 255], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int acc = 0, bin = 255;
        for (; bin > 0; --bin) {
          if (acc + hist[bin] >= remaining) break;
          acc += hist[bin];
        }
        s_sel[0] = bin; s_sel[1] = remaining - acc;
      }
      __syncthreads();
      prefix |= (unsigned)s_sel[0] << shift;
      prefix_mask |= 255u << shift;
      remaining = s_sel[1];
      __syncthreads();
    }
    thr_val = prefix;
    need_eq = remaining;
  } else {
    int mx = 0;
    for (int h = threadIdx.x; h < HN; h += 256) mx = max(mx, cnt[h]);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, m));
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = mx;
    __syncthreads();
    for (int w = 0; w < 8; ++w) mx = max(mx, s_warp[w]);
    __syncthreads();
    ratio_floor = __fsub_rn(__fdiv_rn((float)mx, tnf), 0.1f);  // max - 0.1 (:394)
  }

  // weights: pass A accumulates sum w, sum w*h; ordered so ties keep the lowest indices
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  double acc[6] = {0, 0, 0, 0, 0, 0};  // sw, swx, swy, sxx, sxy, syy
  double mx_ = 0, my_ = 0;
  for (int phase = 0; phase < 2; ++phase) {
    if (phase == 1) {
      block_sum_d<6>(acc, s_red);
      if (with_mean) {
        mx_ = mean_in[((size_t)b * p.vn + v) * 2]; my_ = mean_in[((size_t)b * p.vn + v) * 2 + 1];
      } else {
        mx_ = acc[1] / acc[0]; my_ = acc[2] / acc[0];
      }
      if (threadIdx.x == 0) s_carry = 0;
      __syncthreads();
    }
    for (int base = 0; base < HN; base += 256) {
      const int h = base + threadIdx.x;
      const bool valid = h < HN;
      const unsigned c = valid ? (unsigned)cnt[h] : 0u;
      float w = 0.f;
      if (with_mean) {  // block-uniform branch
        if (valid) {
          const float r = __fdiv_rn((float)c, tnf);
          w = (r < ratio_floor) ? 0.f : r;
        }
      } else {
        const bool eq = valid && c == thr_val;
        // rank among ties, in index order (all threads take every barrier)
        const unsigned bal = __ballot_sync(FULL, eq);
        const int before_in_warp = __popc(bal & ((1u << (threadIdx.x & 31)) - 1));
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = __popc(bal);
        __syncthreads();
        int before = s_carry + before_in_warp;
        for (int wv = 0; wv < (int)(threadIdx.x >> 5); ++wv) before += s_warp[wv];
        const bool keep = valid && (c > thr_val || (eq && before < need_eq));
        w = keep ? __fdiv_rn((float)c, tnf) : 0.f;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int wv = 0; wv < 8; ++wv) t += s_warp[wv]; s_carry += t; }
        __syncthreads();
      }
      if (w != 0.f) {
        const float2 q = make_float2(hypx[h], hypy[h]);
        if (phase == 0) { acc[0] += w; acc[1] += (double)w * q.x; acc[2] += (double)w * q.y; }
        else {
          const double ddx = q.x - mx_, ddy = q.y - my_;
          acc[3] += w * ddx * ddx; acc[4] += w * ddx * ddy; acc[5] += w * ddy * ddy;
        }
      }
    }
  }
  const double sw = acc[0];  // reduced in phase switch
  double c2[3] = {acc[3], acc[4], acc[5]};
  block_sum_d<3>(c2, s_red);
  if (threadIdx.x == 0) {
    const double den = with_mean ? sw + 1e-3 : sw;
    mo[0] = (float)mx_; mo[1] = (float)my_;
    co[0] = (float)(c2[0] / den); co[1] = co[2] = (float)(c2[1] / den); co[3] = (float)(c2[2] / den);
  }
}

// ------------------------------------------------------------------------------------------
// 9. ransac_motion_voting (ransac_voting_gpu.py:960-981): mean over the foreground of (vector + pixel
//    coordinate) per keypoint; zeros for an empty mask.  CTA per (image, keypoint), FP64 sums.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
motion_mean_kernel(epb_voting_params p, Workspace ws, float* __restrict__ pts) {
  const int v = blockIdx.x, b = blockIdx.y;
  __shared__ double s_red[2 * 8];
  const int tn = ws.live[b] ? ws.tn[b] : 0;
  const uint32_t* fp = ws.fgpix + (size_t)b * ws.cap;
  const float2* dir = ws.direct + ((size_t)b * p.vn + v) * ws.cap;
  double acc[2] = {0, 0};
  for (int t = threadIdx.x; t < tn; t += 256) {
    const uint32_t q = __ldg(fp + t);
    const float2 d = __ldg(dir + t);
    // float32 "cur_vert + coords" like the reference (:977), accumulated in double
    acc[0] += (double)__fadd_rn(d.x, (float)(q & 0xffff));
    acc[1] += (double)__fadd_rn(d.y, (float)(q >> 16));
  }
  block_sum_d<2>(acc, s_red);
  if (threadIdx.x == 0) {
    float2* out = reinterpret_cast<float2*>(pts) + (size_t)b * p.vn + v;
    *out = tn > 0 ? make_float2((float)(acc[0] / tn), (float)(acc[1] / tn)) : make_float2(0.f, 0.f);
  }
}

// ------------------------------------------------------------------------------------------
// pybind-level primitives (reference tensor layouts)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
generate_hypothesis_kernel(const float* __restrict__ direct, const float* __restrict__ coords,
                           const int32_t* __restrict__ idxs, float* __restrict__ hypo, int tn, int vn,
                           int hn) {
  const int hvi = blockIdx.x * 256 + threadIdx.x;
  if (hvi >= hn * vn) return;
  const int vi = hvi % vn;
  const int t0 = idxs[hvi * 2], t1 = idxs[hvi * 2 + 1];
  float x = 0.f, y = 0.f, ix, iy;
  if (intersect_rays(direct[(size_t)t0 * vn * 2 + vi * 2], direct[(size_t)t0 * vn * 2 + vi * 2 + 1],
                     coords[t0 * 2], coords[t0 * 2 + 1], direct[(size_t)t1 * vn * 2 + vi * 2],
                     direct[(size_t)t1 * vn * 2 + vi * 2 + 1], coords[t1 * 2], coords[t1 * 2 + 1], &ix, &iy)) {
    x = ix; y = iy;
  }
  hypo[hvi * 2] = x;
  hypo[hvi * 2 + 1] = y;
}

__global__ void __launch_bounds__(256)
voting_for_hypothesis_kernel(const float* __restrict__ direct, const float* __restrict__ coords,
                             const float* __restrict__ hypo, uint8_t* __restrict__ inliers, int tn,
                             int vn, int hn, float thresh) {
  const int ti = blockIdx.x * 256 + threadIdx.x;
  const int hv = blockIdx.y;  // hi*vn + vi
  if (ti >= tn) return;
  const int vi = hv % vn;
  const float nx = direct[(size_t)ti * vn * 2 + vi * 2], ny = direct[(size_t)ti * vn * 2 + vi * 2 + 1];
  if (vote_exact(coords[ti * 2], coords[ti * 2 + 1], nx, ny, dir_norm(nx, ny), hypo[hv * 2],
                 hypo[hv * 2 + 1], thresh))
    inliers[(size_t)hv * tn + ti] = 1;
}

// ransac_voting_kernel.cu:170-229 with the contraction of the reference source as built for sm_100a
// (SASS of oracle/_ref/libref_voting.so; bitwise-checked on the GPU by tests/test_voting_gpu.py).
__global__ void __launch_bounds__(256)
generate_hypothesis_vp_kernel(const float* __restrict__ direct, const float* __restrict__ coords,
                              const int32_t* __restrict__ idxs, float* __restrict__ hypo, int tn,
                              int vn, int hn) {
  const int hvi = blockIdx.x * 256 + threadIdx.x;
  if (hvi >= hn * vn) return;
  const int vi = hvi % vn;
  const int id0 = idxs[hvi * 2], id1 = idxs[hvi * 2 + 1];
  const float dx0 = direct[(size_t)id0 * vn * 2 + vi * 2], dy0 = direct[(size_t)id0 * vn * 2 + vi * 2 + 1];
  const float cx0 = coords[id0 * 2], cy0 = coords[id0 * 2 + 1];
  const float dx1 = direct[(size_t)id1 * vn * 2 + vi * 2], dy1 = direct[(size_t)id1 * vn * 2 + vi * 2 + 1];
  const float cx1 = coords[id1 * 2], cy1 = coords[id1 * 2 + 1];
  const float lz0 = __fmaf_rn(dx0, cy0, -__fmul_rn(dy0, cx0));
  const float lz1 = __fmaf_rn(dx1, cy1, -__fmul_rn(dy1, cx1));
  float z = __fmaf_rn(dx0, dy1, -__fmul_rn(dy0, dx1));
  float x = __fmaf_rn(dx1, lz0, -__fmul_rn(dx0, lz1));
  float y = __fmaf_rn(dy1, lz0, -__fmul_rn(dy0, lz1));
  const float val_x0 = __fmul_rn(dx0, __fmaf_rn(-cx0, z, x)), val_x1 = __fmul_rn(dx1, __fmaf_rn(-cx1, z, x));
  const float val_y0 = __fmul_rn(dy0, __fmaf_rn(-cy0, z, y)), val_y1 = __fmul_rn(dy1, __fmaf_rn(-cy1, z, y));
  if (val_x0 < 0 && val_x1 < 0 && val_y0 < 0 && val_y1 < 0) { z = -z; x = -x; y = -y; }
  if (__fmul_rn(val_x0, val_x1) < 0 || __fmul_rn(val_y0, val_y1) < 0) { x = 0.f; y = 0.f; z = 0.f; }
  hypo[hvi * 3] = x; hypo[hvi * 3 + 1] = y; hypo[hvi * 3 + 2] = z;
}

// ransac_voting_kernel.cu:268-310 as compiled: diff = fma(-c, hz, h); cosine numerator = the un-fused sum
// of the two products the sign test reuses.
__global__ void __launch_bounds__(256)
voting_for_hypothesis_vp_kernel(const float* __restrict__ direct, const float* __restrict__ coords,
                                const float* __restrict__ hypo, uint8_t* __restrict__ inliers, int tn,
                                int vn, int hn, float thresh) {
  const int ti = blockIdx.x * 256 + threadIdx.x;
  const int hv = blockIdx.y;
  if (ti >= tn) return;
  const int vi = hv % vn;
  const float cx = coords[ti * 2], cy = coords[ti * 2 + 1];
  const float hx = hypo[hv * 3], hy = hypo[hv * 3 + 1], hz = hypo[hv * 3 + 2];
  const float ddx = direct[(size_t)ti * vn * 2 + vi * 2], ddy = direct[(size_t)ti * vn * 2 + vi * 2 + 1];
  const float fx = __fmaf_rn(-cx, hz, hx), fy = __fmaf_rn(-cy, hz, hy);
  const float norm1 = __fsqrt_rn(__fmaf_rn(ddx, ddx, __fmul_rn(ddy, ddy)));
  const float norm2 = __fsqrt_rn(__fmaf_rn(fx, fx, __fmul_rn(fy, fy)));
  if ((double)norm1 < 1e-6 || (double)norm2 < 1e-6) return;
  const float vx = __fmul_rn(fx, ddx), vy = __fmul_rn(fy, ddy);
  const float ad = __fdiv_rn(__fadd_rn(vx, vy), __fmul_rn(norm1, norm2));
  if (vx < 0 || vy < 0) return;
  if (fabsf(ad) > thresh) inliers[(size_t)hv * tn + ti] = 1;
}

}  // namespace epb

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace epb;

static void voting_kernel_attributes() {
  prefer_max_shared(mask_count_kernel); prefer_max_shared(mask_scan_kernel); prefer_max_shared(rng_offsets_kernel);
  prefer_max_shared(mask_scatter_kernel<true>); prefer_max_shared(mask_scatter_kernel<false>);
  prefer_max_shared(field_gather_kernel); prefer_max_shared(hypothesis_kernel); prefer_max_shared(vote_items_kernel);
  // vote_count: 19.6 KB per CTA (records + one TMA raw stage + queue).  The library-wide carve-out (60 %, common.cuh)
  // holds 6 of them per SM; asking for more (75 %, 100 %: 8 CTAs) measured the same device-resident time (1.75 ms)
  // and a slower host pipeline (24.8 k vs 26.3 k poses/s end to end), where the gather and pose kernels share the SMs
  prefer_max_shared(vote_count_kernel<2>); prefer_max_shared(vote_count_kernel<4>); prefer_max_shared(vote_count_kernel<8>);
  prefer_max_shared(counts_export_kernel); prefer_max_shared(winner_refine_kernel); prefer_max_shared(distribution_kernel);
  prefer_max_shared(motion_mean_kernel);
  prefer_max_shared(generate_hypothesis_kernel); prefer_max_shared(voting_for_hypothesis_kernel);
  prefer_max_shared(generate_hypothesis_vp_kernel); prefer_max_shared(voting_for_hypothesis_vp_kernel);
  cudaGetLastError();
}

static int g_vote_r_large = 4;

// Work units of the vote kernel: (item of <= item_px voting pixels, keypoint, hypothesis chunk), listed on the
// device (vote_items_kernel) so that ragged batches balance without a host sync; CTAs beyond the list exit at once.
static int launch_vote_ffma(const epb_voting_params& p, const Workspace& ws, const VoteConsts& vc, bool headroom,
                            cudaStream_t s) {
  const int HN = p.hn * p.rounds;
  int R = HN <= 256 ? 2 : (HN <= 512 ? 4 : g_vote_r_large);
  { const int forced = tuning_int("EPB_VOTE_R", 0); if (forced == 2 || forced == 4 || forced == 8) R = forced; }
  const int chunks = (HN + VOTE_THREADS * R - 1) / (VOTE_THREADS * R);
  const long long slots = (long long)device_sm_count() * (R <= 4 ? 8 : 4);
  int item_px = 4 * VOTE_TILE;
  while (item_px > VOTE_TILE &&
         (long long)p.B * p.vn * chunks * ((ws.cap + item_px - 1) / item_px) < 4 * slots) item_px >>= 1;
  const long long max_units = (long long)p.B * p.vn * chunks * ((ws.cap + item_px - 1) / item_px);
  if (max_units > 0x7fffffffLL) return EPB_ERR_INVALID;
  const unsigned grid = (unsigned)max_units;
  vote_items_kernel<<<1, 256, 0, s>>>(p.B, item_px, ws);
  EPB_RETURN_IF(check_launch());
  EPB_RETURN_IF(check_api(cudaMemsetAsync(ws.counts, 0, (size_t)p.B * p.vn * HN * 4, s)));
  // Split runs (host pipeline) overlap this kernel with the gather stage of the next batch chunk on a high-priority
  // stream.  Its small kernels can only start when an SM has registers to spare, so the vote CTAs are held two
  // below full occupancy there (unused dynamic shared memory is the occupancy knob; measured in round 1: gather
  // 0.75 -> 0.58 ms per chunk under a concurrent vote, vote 0.68 -> 0.61 ms).  Only a host-resident field keeps
  // the gather stage busy long enough to matter: it reads over PCIe.
  size_t pad = 0;
  const int occ = R <= 4 ? 8 : 4;
  if (headroom && occ >= 6) {
    const int want = occ - 2;
    const size_t per_sm = 227 * 1024, used = sizeof(VoteSmem) + 1024;
    if (per_sm / (size_t)(want + 1) + 1024 > used) pad = per_sm / (size_t)(want + 1) + 1024 - used;
    if (pad + used > 48 * 1024) pad = 0;       // stay within the default dynamic shared memory limit
  }
  if (R == 2) vote_count_kernel<2><<<grid, VOTE_THREADS, pad, s>>>(p, ws, vc, item_px, chunks);
  else if (R == 4) vote_count_kernel<4><<<grid, VOTE_THREADS, pad, s>>>(p, ws, vc, item_px, chunks);
  else vote_count_kernel<8><<<grid, VOTE_THREADS, pad, s>>>(p, ws, vc, item_px, chunks);
  return check_launch();
}
  // hypotheses per thread when HN > 512 (measured at HN = 2048: 1.48 ms with 4, 1.66 ms with 8)

static bool params_ok(const epb_voting_params* p) {
  if (!p) return false;
  if (p->B <= 0 || p->H <= 0 || p->W <= 0 || p->vn <= 0 || p->hn <= 0 || p->rounds <= 0) return false;
  if (p->H > 65535 || p->W > 65535) return false;
  if ((long long)p->H * p->W > 0x7fffffffLL) return false;
  if (p->mode < EPB_VOTE_V3 || p->mode > EPB_VOTE_MOTION) return false;
  if (p->classes < 0 || (long long)p->B * (p->classes < 1 ? 1 : p->classes) > 65535) return false;
  if (p->mask_mode < EPB_MASK_NONZERO || p->mask_mode > EPB_MASK_CLASS) return false;
  if (p->rng_mode < EPB_RNG_IDXS || p->rng_mode > EPB_RNG_PHILOX) return false;
  if (p->stage < EPB_STAGE_ALL || p->stage > EPB_STAGE_VOTE) return false;
  if ((long long)p->hn * p->rounds > (1 << 24)) return false;
  return true;
}

// classes (image, class) pairs per image become one virtual batch (see mask_pred)
static epb_voting_params virtual_batch(epb_voting_params p) {
  if (p.classes < 1) p.classes = 1;
  p.B *= p.classes;
  return p;
}

extern "C" size_t epb_voting_workspace_bytes(const epb_voting_params* p) {
  if (!params_ok(p)) return 0;
  return carve(virtual_batch(*p), nullptr).bytes;
}

extern "C" int epb_voting_run(const epb_voting_params* pp, const epb_voting_io* io, void* workspace,
                              size_t workspace_bytes, void* stream) {
  if (!params_ok(pp) || !io || !workspace) return EPB_ERR_INVALID;
  const epb_voting_params p = virtual_batch(*pp);
  if (p.stage != EPB_STAGE_VOTE && (!io->mask || !io->vertex)) return EPB_ERR_INVALID;
  EPB_INIT_ONCE_PER_DEVICE(voting_kernel_attributes);
  if (p.rng_mode != EPB_RNG_PHILOX && !io->idxs && p.mode != EPB_VOTE_MOTION) return EPB_ERR_INVALID;
  const bool is_layer = p.mode <= EPB_VOTE_V5 || p.mode == EPB_VOTE_V1 || p.mode == EPB_VOTE_V2 ||
                        p.mode == EPB_VOTE_MOTION;
  const bool is_dist = p.mode == EPB_VOTE_DISTRIBUTION || p.mode == EPB_VOTE_DISTRIBUTION_WITH_MEAN;
  if (p.stage != EPB_STAGE_GATHER) {   // the gather half produces nothing but the workspace
    if (is_layer && !io->pts) return EPB_ERR_INVALID;
    if ((p.mode == EPB_VOTE_V4 || p.mode == EPB_VOTE_V5) && !io->var_or_conf) return EPB_ERR_INVALID;
    if (p.mode == EPB_VOTE_HYPOTHESIS && (!io->hyp || !io->counts)) return EPB_ERR_INVALID;
    if (is_dist && (!io->mean || !io->cov)) return EPB_ERR_INVALID;
    if (p.mode == EPB_VOTE_DISTRIBUTION_WITH_MEAN && !io->mean_in) return EPB_ERR_INVALID;
  }
  if (is_dist && p.topk <= 0 && p.mode == EPB_VOTE_DISTRIBUTION) return EPB_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return EPB_ERR_INVALID;
  Workspace ws = carve(p, workspace);
  if (workspace_bytes < ws.bytes) return EPB_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;

  const int HW = p.H * p.W;
  const int T = (HW + TILE_PX - 1) / TILE_PX;
  const int HN = p.hn * p.rounds;
  const TorchRngLayout lu = torch_rng_layout((unsigned long long)HW, p.philox_sm_count, p.philox_threads_per_sm);
  const TorchRngLayout lr = torch_rng_layout((unsigned long long)p.hn * p.vn * 2, p.philox_sm_count,
                                             p.philox_threads_per_sm);
  SubsampleCtx sc;
  sc.selection = io->selection;
  sc.seed = p.philox_seed;
  sc.total_threads = lu.total_threads;
  sc.use_philox = (p.rng_mode == EPB_RNG_PHILOX) && !io->selection;
  // the subsample can only trigger when an image can hold more than max_num foreground pixels
  const int allow_sub = (HW > p.max_num) && (io->selection || sc.use_philox);

  if (p.stage != EPB_STAGE_VOTE) {
  prof_begin(PROF_COMPACT, s);
  mask_count_kernel<<<dim3(T, p.B), 256, 0, s>>>(io->mask, HW, T, p.mask_mode, p.classes, 0, ws, sc);
  EPB_RETURN_IF(check_launch());
  mask_scan_kernel<<<p.B, 256, 0, s>>>(T, 0, p.min_num, p.max_num, allow_sub, ws);
  EPB_RETURN_IF(check_launch());
  rng_offsets_kernel<<<1, 32, 0, s>>>(p.B, p.rounds, p.philox_offset, lu.inc, lr.inc, ws,
                                      io->philox_consumed, io->philox_state);
  EPB_RETURN_IF(check_launch());
  if (allow_sub) {
    mask_count_kernel<<<dim3(T, p.B), 256, 0, s>>>(io->mask, HW, T, p.mask_mode, p.classes, 1, ws, sc);
    EPB_RETURN_IF(check_launch());
    mask_scan_kernel<<<p.B, 256, 0, s>>>(T, 1, p.min_num, p.max_num, allow_sub, ws);
    EPB_RETURN_IF(check_launch());
  }
  // planar field (NCHW network output seen through vertex_layer_reshape): gather fused into the scatter
  const bool planar = p.sx == 1 && p.sy == p.W;   // pixel px of a plane is element px: scalar, line-coalesced reads
  const bool host_field = is_host_pointer(io->vertex);
  // The fused form exists for host-resident fields (whole 128-byte lines over PCIe).  In HBM the plain pair --
  // compaction of the coordinates, then one thread per compacted pixel gathering its keypoints -- is faster
  // (compaction class 0.113 -> 0.082 ms on config[1]): the fused kernel spends its instructions on the 3/4 of the
  // 32-pixel groups that hold no foreground.
  if (planar && !(tuning_int("EPB_GATHER_SPLIT", 1) && !host_field)) {
    const int sm_n = device_sm_count();
    const long long work = (long long)T * p.B;
    const int per_sm = tuning_int("EPB_GATHER_CTAS_PER_SM", 1), total_cap = tuning_int("EPB_GATHER_CTAS", 0);
    const long long cap = total_cap > 0 ? total_cap : (long long)sm_n * (per_sm > 0 ? per_sm : 1);
    const unsigned g = (unsigned)(host_field && work > cap ? cap : work);
    const int light_knob = tuning_int("EPB_GATHER_LIGHT", 1);
    // device-resident field: split the keypoints over gridDim.y so that the grid is several waves deep
    // (the gather is HBM-bound and each CTA is register-heavy); host field: one slim persistent wave
    const int vsplit_knob = tuning_int("EPB_GATHER_VSPLIT", 0);
    int vsplit = 1;
    if (!host_field) {
      vsplit = vsplit_knob > 0 ? vsplit_knob : (work < 16LL * sm_n ? 2 : 1);
      if (vsplit > p.vn) vsplit = p.vn;
    }
    // The scatter is the kernel that shares SMs with the vote kernel in the host pipeline, where it must ask for a
    // large shared-memory carve-out (common.cuh); alone on the device it is 20 % faster with the default
    // split (88 vs 111 us), so the preference follows the kind of field (attribute read at launch).
    {
      static int current[64];   // per device: 0 unknown, 1 default split, 2 large carve-out (the call is not free)
      const int want = host_field ? 2 : 1;
      int dev_id = 0;
      cudaGetDevice(&dev_id);
      if (dev_id >= 0 && dev_id < 64 && __atomic_load_n(&current[dev_id], __ATOMIC_ACQUIRE) != want) {
        // (two threads racing here both set the same attribute value for their `want`; the last one wins and the
        // cache then says so -- the attribute is a preference, never a correctness matter)
        cudaFuncSetAttribute(reinterpret_cast<const void*>(mask_scatter_kernel<true>),
                             cudaFuncAttributePreferredSharedMemoryCarveout,
                             host_field ? carveout_percent() : (int)cudaSharedmemCarveoutDefault);
        __atomic_store_n(&current[dev_id], want, __ATOMIC_RELEASE);
      }
    }
    mask_scatter_kernel<true><<<dim3(g, vsplit), 256, 0, s>>>(io->mask, p.H, p.W, T, p.mask_mode, ws, sc, io->vertex, p,
                                                              host_field && light_knob);
    EPB_RETURN_IF(check_launch());
  } else {
    mask_scatter_kernel<false><<<(unsigned)((long long)T * p.B), 256, 0, s>>>(io->mask, p.H, p.W, T, p.mask_mode, ws, sc, io->vertex, p, 0);
    EPB_RETURN_IF(check_launch());
    field_gather_kernel<<<dim3((HW + 255) / 256, p.B), 256, 0, s>>>(io->vertex, p, ws);
    EPB_RETURN_IF(check_launch());
  }
  prof_end(PROF_COMPACT, s);
  }
  if (p.stage == EPB_STAGE_GATHER) return EPB_OK;
  if (io->tn_out)
    EPB_RETURN_IF(check_api(cudaMemcpyAsync(io->tn_out, ws.tn, (size_t)p.B * 4, cudaMemcpyDeviceToDevice, s)));

  if (p.mode == EPB_VOTE_MOTION) {   // no hypotheses: the mean of (vector + coordinate) over the foreground
    ProfScope ps_m(PROF_REFINE, s);
    motion_mean_kernel<<<dim3(p.vn, p.B), 256, 0, s>>>(p, ws, io->pts);
    return check_launch();
  }
  HypCtx hc;
  hc.idxs = io->idxs;
  hc.rng_mode = p.rng_mode;
  hc.seed = p.philox_seed;
  hc.total_threads_r = lr.total_threads;
  hc.inc_r = lr.inc;
  {
    const long long n = (long long)p.B * HN * p.vn;
    ProfScope ps(PROF_HYPOTHESIS, s);
    hypothesis_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, ws, hc, io->hyp);
    EPB_RETURN_IF(check_launch());
  }
  {
    const VoteConsts vc = vote_consts(p.inlier_thresh, p.H, p.W);
    ProfScope ps(PROF_VOTE_COUNT, s);
    bool done = false;
#ifdef EPB_TUNING
    if (tuning_int("EPB_VOTE_IMPL", 0) == 1) {   // the tensor-core form (vote_mma.cuh), A/B measurements only
      VoteConsts vm = vote_consts_mma(p.inlier_thresh, p.H, p.W);
      if (vm.fast_ok) {
        vm.fast_ok |= tuning_int("EPB_VM_DEBUG", 0) << 8;
        EPB_RETURN_IF(launch_vote_mma(p, ws, vm, s));
        done = true;
      }
    }
#endif
    if (!done) {
      const bool headroom = p.stage == EPB_STAGE_VOTE && is_host_pointer(io->vertex);
      if (vc.fast_ok) EPB_RETURN_IF(launch_vote_ffma(p, ws, vc, headroom, s));
      else {   // thresholds the band test cannot serve: the reference expression for every pair
        vote_exact_kernel<<<dim3((HN + 255) / 256, p.vn, p.B), 256, 0, s>>>(p, ws);
        EPB_RETURN_IF(check_launch());
      }
    }
  }
  if (io->counts) {
    const long long n = (long long)p.B * HN * p.vn;
    counts_export_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, ws, io->counts);
    EPB_RETURN_IF(check_launch());
  }
  ProfScope ps_tail(PROF_REFINE, s);
  if (is_layer) {
    winner_refine_kernel<<<dim3(p.vn, p.B), 256, 0, s>>>(p, ws, io->pts, io->var_or_conf,
                                                         io->status);
    EPB_RETURN_IF(check_launch());
  } else if (is_dist) {
    distribution_kernel<<<dim3(p.vn, p.B), 256, 0, s>>>(p, ws, io->mean_in, io->mean, io->cov);
    EPB_RETURN_IF(check_launch());
  }
  return EPB_OK;
}

extern "C" int epb_generate_hypothesis(const float* direct, const float* coords, const int32_t* idxs,
                                       float* hypo, int tn, int vn, int hn, void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(voting_kernel_attributes);
  if (!direct || !coords || !idxs || !hypo || tn <= 0 || vn <= 0 || hn <= 0) return EPB_ERR_INVALID;
  generate_hypothesis_kernel<<<(hn * vn + 255) / 256, 256, 0, (cudaStream_t)stream>>>(direct, coords, idxs,
                                                                                  hypo, tn, vn, hn);
  return check_launch();
}

extern "C" int epb_voting_for_hypothesis(const float* direct, const float* coords, const float* hypo,
                                         uint8_t* inliers, int tn, int vn, int hn, float thresh,
                                         void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(voting_kernel_attributes);
  if (!direct || !coords || !hypo || !inliers || tn <= 0 || vn <= 0 || hn <= 0) return EPB_ERR_INVALID;
  if ((long long)hn * vn > 65535) {
    // grid.y limit: process in slabs of hypotheses
    const int per = 65535 / vn;
    for (int h0 = 0; h0 < hn; h0 += per) {
      const int n = hn - h0 < per ? hn - h0 : per;
      voting_for_hypothesis_kernel<<<dim3((tn + 255) / 256, n * vn), 256, 0, (cudaStream_t)stream>>>(
          direct, coords, hypo + (size_t)h0 * vn * 2, inliers + (size_t)h0 * vn * tn, tn, vn, n, thresh);
      EPB_RETURN_IF(check_launch());
    }
    return EPB_OK;
  }
  voting_for_hypothesis_kernel<<<dim3((tn + 255) / 256, hn * vn), 256, 0, (cudaStream_t)stream>>>(
      direct, coords, hypo, inliers, tn, vn, hn, thresh);
  return check_launch();
}

extern "C" int epb_generate_hypothesis_vanishing_point(const float* direct, const float* coords,
                                                       const int32_t* idxs, float* hypo3, int tn, int vn,
                                                       int hn, void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(voting_kernel_attributes);
  if (!direct || !coords || !idxs || !hypo3 || tn <= 0 || vn <= 0 || hn <= 0) return EPB_ERR_INVALID;
  generate_hypothesis_vp_kernel<<<(hn * vn + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      direct, coords, idxs, hypo3, tn, vn, hn);
  return check_launch();
}

extern "C" int epb_voting_for_hypothesis_vanishing_point(const float* direct, const float* coords,
                                                         const float* hypo3, uint8_t* inliers, int tn,
                                                         int vn, int hn, float thresh, void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(voting_kernel_attributes);
  if (!direct || !coords || !hypo3 || !inliers || tn <= 0 || vn <= 0 || hn <= 0) return EPB_ERR_INVALID;
  if ((long long)hn * vn > 65535) {
    // grid.y limit: process in slabs of hypotheses (like epb_voting_for_hypothesis)
    const int per = 65535 / vn;
    if (per < 1) return EPB_ERR_INVALID;
    for (int h0 = 0; h0 < hn; h0 += per) {
      const int n = hn - h0 < per ? hn - h0 : per;
      voting_for_hypothesis_vp_kernel<<<dim3((tn + 255) / 256, n * vn), 256, 0, (cudaStream_t)stream>>>(
          direct, coords, hypo3 + (size_t)h0 * vn * 3, inliers + (size_t)h0 * vn * tn, tn, vn, n, thresh);
      EPB_RETURN_IF(check_launch());
    }
    return EPB_OK;
  }
  voting_for_hypothesis_vp_kernel<<<dim3((tn + 255) / 256, hn * vn), 256, 0, (cudaStream_t)stream>>>(
      direct, coords, hypo3, inliers, tn, vn, hn, thresh);
  return check_launch();
}
