// Tensor-core form of the vote test (split-TF32 mma.sync).  TUNING BUILDS ONLY (-DEPB_TUNING, EPB_VOTE_IMPL=1):
// the counts are bit-exact (all of tests/test_voting_gpu.py pass with it), but on config[1] it runs at 2.4 ms per
// launch against 1.7 ms of the packed-FP32 kernel it was meant to replace -- see DESIGN.md section 5b and
// profiles/r2_vote_mma_history.md for the measurements and why.  Included by voting.cu inside namespace epb.
#pragma once

// Constants of the tensor-core form (vote_mma_kernel; error budget in DESIGN.md section 5b).
//   direction: n^ = n / |n| carries two independent final roundings -> rotated by <= u; beta uses 12.5 u
//   a' terms:  5 u each (k rounded, k n^ rounded, hi/lo residual of A, of h, dropped lo*lo); A3: 4 u
//   p  terms:  3 u each (residual of B, of h, dropped lo*lo); B3: 2 u of the FMA + 1 u residual
//   tensor core: accumulating the 8 products in FP32 with truncation: <= CM u sum|terms| (CM = 2x the worst
//              case measured by tools/micro/mma_tf32_probe.cu, rounded up)
//   band w:    computed from the high halves only and with gamma A truncated to TF32: relative 2^-9 of the
//              terms of a', i.e. gamma k 2^-8 (|h|_1 + |c|_1) in absolute terms, folded into EH; EH itself is
//              rounded up by 2^-9 because the tensor core truncates it to TF32
constexpr double VOTE_MMA_CM = 8.0;
__host__ inline VoteConsts vote_consts_mma(float thresh, int H, int W) {
  VoteConsts c;
  const double T = (double)thresh, u = 5.9604644775390625e-08;  // 2^-24
  c.fast_ok = 0; c.kf = 0.f; c.gamma = 0.f; c.eh_scale = 0.f; c.eh_abs = 0.f;
  if (!(T >= 0.5 && T < 1.0)) return c;
  const double k = sqrt((1.0 - T) * (1.0 + T)) / T;
  const double theta = acos(T);
  if (!(2.0 * sin(0.5 * theta) * sin(0.25 * theta) >= 12.5 * u)) return c;
  const double beta = 12.5 * u / (T * sin(0.75 * theta));
  const double gamma = beta * (1.0 / k + 1.0) * 1.0001 * (1.0 + 1.0 / 256.0);
  if (!(gamma <= 0.25)) return c;
  const double es = (1.5 * ((5.0 + VOTE_MMA_CM) * k + (3.0 + VOTE_MMA_CM)) * u + gamma * k / 256.0) * (1.0 + 1.0 / 256.0);
  const double ea = es * (double)(H + W) + 3.5e-6 * (1.0 + k) * (1.0 + 1.0 / 256.0);
  c.fast_ok = 1;
  c.kf = (float)k;
  c.gamma = nextafterf((float)gamma, INFINITY);
  c.eh_scale = nextafterf((float)es, INFINITY);
  c.eh_abs = nextafterf((float)ea, INFINITY);
  // reference guard (ransac_voting_kernel.cu:121): norm1 = sqrt.rn(s1) as float, compared in double with 1e-6.
  // sqrt.rn is monotone, so the pixels that pass are exactly { s1 >= s1_min }: bisection over the float bit patterns
  // (sqrtf on the host is correctly rounded, like sqrt.rn)
  {
    unsigned lo = 0x00800000u, hi = 0x3f800000u;     // sqrt(lo) ~ 1e-19 fails, sqrt(1) passes
    while (hi - lo > 1) {
      const unsigned mid = lo + (hi - lo) / 2;
      float f; memcpy(&f, &mid, 4);
      if ((double)sqrtf(f) < 1e-6) lo = mid; else hi = mid;
    }
    memcpy(&c.s1_min, &hi, 4);
  }
  return c;
}


// ------------------------------------------------------------------------------------------
// 6b. vote_mma: the same exact counts with the two affine forms on the tensor cores.
//
// a' and p are rank-3 bilinear forms of (record, hypothesis).  Each is evaluated as ONE K = 8 TF32 MMA with
// both operands split in a high and a low half (hi*hi + lo*hi + hi*lo; the dropped lo*lo term is < 2^-24 of
// the product), accumulated in FP32:
//   record side   (A1h, A2h, A3h, A1h | A1l, A2l, A3l, A2h)         A = k n^, A3 = -k n^.c   (n^ = n / |n|)
//   hypothesis    (hxh, hyh,  1 , hxl | hxh, hyh,  1 , hyl)         same vector for the p form (B = (-n^y, n^x), B3)
// and the band w = gamma a' + EH[h] -- itself affine in the hypothesis -- as a K = 4 MMA of the high halves
// (gamma A1h, gamma A2h, gamma A3h, 1) x (hxh, hyh, 1, EH[h]).  One m16n8k8 tile = 8 records x {a', p} rows x 8
// hypotheses, so a lane finds a'(t,h) and p(t,h) of the same pair in its own accumulator registers; what is
// left for the FP32 pipe per pair is  m = a' - |p|,  |m| > w,  and the sign-bit count: 3 instructions instead
// of 7.  Pairs inside the band take the reference expression, exactly as in vote_count (queue + dense
// resolution), so the counts remain the reference's integers; DESIGN.md section 5b has the error budget of
// the split (5 u per A term, 3 u per B term) and of the tensor core's accumulation (measured,
// tools/micro/mma_tf32_probe.cu) that EH covers.
//
// CTA = 8 consumer warps (64 hypotheses each: 8 n-blocks whose B fragments, bands and counters stay in
// registers for the whole unit) + 1 producer warp.  Work unit = (item of pixels, keypoint, chunk of 512
// hypotheses).  The producer's lane 0 streams the raw tiles (128 x float2 direction + 128 x packed pixel)
// global -> shared with cp.async.bulk on mbarriers (TMA, no register staging); the producer warp turns a raw
// tile into 128 48-byte records (normalisation, k-scaling, hi/lo split: ~70 instructions per record, once per
// record and chunk) in a 3-deep ring that the consumers read as MMA A fragments.
// ------------------------------------------------------------------------------------------
constexpr int VM_WARPS = 8;                      // warps per CTA; every warp converts AND consumes
constexpr int VM_NB = 4;                         // n-blocks of 8 hypotheses per consumer warp
constexpr int VM_CHUNK = VM_WARPS * VM_NB * 8;   // 256 hypotheses per work unit
constexpr int VM_TILE = 128;                     // records per stage
constexpr int VM_REC_STAGES = 3;
constexpr int VM_RAW_STAGES = 4;
constexpr int VM_THREADS = VM_WARPS * 32;        // 256: two CTAs per SM own the whole register file at 128 registers
#ifndef VM_PREDICATED
#define VM_PREDICATED 0
#endif
constexpr int VM_QCAP = 2048;                    // deferred undecided pairs per work unit

constexpr int VM_REC_F4 = VM_TILE * 4;          // float4 per record stage: [8-pixel group][g][j] = one MMA A-fragment lane

// what the out-of-line slow path needs, written once per work unit
struct VmCtx {
  const uint32_t* fp; const float2* dir; const float* hypx; const float* hypy; int32_t* out;
  int t_begin, t_end, h0, HN;
  float T;
};
struct VmSmem {
  // records in FRAGMENT ORDER: for every group of 8 records, lane (g, j) of a consumer warp finds its four A
  // registers (a' row k=j, p row k=j, a' row k=j+4, p row k=j+4 of record g) as one float4: a single
  // conflict-free LDS.128 per 8-record MMA tile (lane j=3 holds copies of the hi halves of j=0/1: 64 B/record)
  alignas(128) float4 rec[VM_REC_STAGES][VM_REC_F4];
  alignas(128) float2 raw_dir[VM_RAW_STAGES][VM_TILE];
  alignas(128) uint32_t raw_pix[VM_RAW_STAGES][VM_TILE];
  unsigned long long raw_full[VM_RAW_STAGES], raw_empty[VM_RAW_STAGES], rec_full[VM_REC_STAGES], rec_empty[VM_REC_STAGES];
  VmCtx ctx;
  unsigned q[VM_QCAP];   // lane-trips (8 pair tests) with an undecided pair: t_rel0 << 9 | h_rel0
  unsigned qn;
  unsigned overflow;     // the queue was full at least once: the unit is recounted with the reference expression
};

__device__ __forceinline__ float tf32_rna(float x) {
  unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r);
}
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
  hi = tf32_rna(x);
  lo = (fabsf(hi) <= FLT_MAX) ? tf32_rna(__fsub_rn(x, hi)) : 0.f;   // x - hi is exact; -inf markers keep lo = 0
}
__device__ __forceinline__ void mma_tf32_k8(float (&d)[4], const float4& a, float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
                 "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)), "f"(0.f));
}
__device__ __forceinline__ void mma_tf32_k4(float (&d)[4], float a0, float a1, float b0) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)), "r"(__float_as_uint(b0)), "f"(0.f));
}

// One raw pixel -> its record (four fragment-lane float4, j = 0..3, of VmSmem::rec), by ONE lane.
__device__ __forceinline__ void make_record(uint32_t q, float nx, float ny, bool in_range, const VoteConsts& vc, float4* dst) {
  const float cx = (float)(q & 0xffff), cy = (float)(q >> 16);
  const float s1 = __fmaf_rn(nx, nx, __fmul_rn(ny, ny));     // the reference's |n|^2 (:113-116 as compiled)
  // reference guard :121 (zero direction), NaN / overflowed |n|^2 (c = NaN or 0): never an inlier
  const bool valid = in_range && (s1 >= vc.s1_min) && (s1 <= FLT_MAX);
  // exact power-of-two scaling first (no over/underflow in the squares), then one common factor ~1/|n| (its own
  // error only scales both forms): the direction of (ux, uy) differs from n's by the two final roundings (<= u)
  const float nm = fmaxf(fabsf(nx), fabsf(ny));
  const float sc = __uint_as_float((254u - ((__float_as_uint(nm) >> 23) & 0xffu)) << 23);
  const float nxs = __fmul_rn(nx, sc), nys = __fmul_rn(ny, sc);
  const float r = rsqrtf(__fmaf_rn(nxs, nxs, __fmul_rn(nys, nys)));
  const float ux = __fmul_rn(nxs, r), uy = __fmul_rn(nys, r);
  float A1 = __fmul_rn(vc.kf, ux), A2 = __fmul_rn(vc.kf, uy);
  float A3 = -__fmul_rn(vc.kf, __fmaf_rn(ux, cx, __fmul_rn(uy, cy)));
  float B1 = -uy, B2 = ux, B3 = __fmaf_rn(uy, cx, -__fmul_rn(ux, cy));
  if (!valid) { A1 = A2 = B1 = B2 = B3 = 0.f; A3 = -INFINITY; }
  float A1h, A1l, A2h, A2l, A3h, A3l, B1h, B1l, B2h, B2l, B3h, B3l;
  tf32_split(A1, A1h, A1l); tf32_split(A2, A2h, A2l); tf32_split(A3, A3h, A3l);
  tf32_split(B1, B1h, B1l); tf32_split(B2, B2h, B2l); tf32_split(B3, B3h, B3l);
  // k = j:     (A1h, A2h, A3h, A1h)      k = j + 4: (A1l, A2l, A3l, A2h)        (same for B)
  dst[0] = make_float4(A1h, B1h, A1l, B1l);
  dst[1] = make_float4(A2h, B2h, A2l, B2l);
  dst[2] = make_float4(A3h, B3h, A3l, B3l);
  dst[3] = make_float4(A1h, B1h, A2h, B2h);
}

// One raw pixel -> its record, by a PAIR of lanes (half = 0 / 1): each writes two of the four fragment-lane
// float4 (j = 0..3) of VmSmem::rec.
//   k = j:     (A1h, A2h, A3h, A1h)      k = j + 4: (A1l, A2l, A3l, A2h)        (same for B)
//   j = 0: (A1h,B1h,A1l,B1l)  j = 1: (A2h,B2h,A2l,B2l)  j = 2: (A3h,B3h,A3l,B3l)  j = 3: (A1h,B1h,A2h,B2h)
// half 0 produces j = 0 and j = 2, half 1 produces j = 1 and j = 3 (its A1h, B1h come from the partner lane).
__device__ __forceinline__ void make_record_half(uint32_t q, float nx, float ny, bool in_range, const VoteConsts& vc,
                                                 int half, float4* dst /* &rec[record * 4] */) {
  const float cx = (float)(q & 0xffff), cy = (float)(q >> 16);
  const float s1 = __fmaf_rn(nx, nx, __fmul_rn(ny, ny));     // the reference's |n|^2 (:113-116 as compiled)
  // reference guard :121 (zero direction), NaN / overflowed |n|^2 (c = NaN or 0): never an inlier
  const bool valid = in_range && (s1 >= vc.s1_min) && (s1 <= FLT_MAX);
  // exact power-of-two scaling first (no over/underflow in the squares), then one common factor ~1/|n| (its own
  // error only scales both forms): the direction of (ux, uy) differs from n's by the two final roundings (<= u)
  const float nm = fmaxf(fabsf(nx), fabsf(ny));
  const float sc = __uint_as_float((254u - ((__float_as_uint(nm) >> 23) & 0xffu)) << 23);
  const float nxs = __fmul_rn(nx, sc), nys = __fmul_rn(ny, sc);
  const float r = rsqrtf(__fmaf_rn(nxs, nxs, __fmul_rn(nys, nys)));
  const float ux = __fmul_rn(nxs, r), uy = __fmul_rn(nys, r);
  // first vector: (A1, B1) for half 0, (A2, B2) for half 1
  float a = __fmul_rn(vc.kf, half ? uy : ux);
  float bb = half ? ux : -uy;
  if (!valid) { a = 0.f; bb = 0.f; }
  float ah, al, bh, bl;
  tf32_split(a, ah, al); tf32_split(bb, bh, bl);
  dst[half] = make_float4(ah, bh, al, bl);
  const float pah = __shfl_xor_sync(FULL, ah, 1), pbh = __shfl_xor_sync(FULL, bh, 1);   // partner's high halves
  if (half) {
    dst[3] = make_float4(pah, pbh, ah, bh);
  } else {
    float a3 = -__fmul_rn(vc.kf, __fmaf_rn(ux, cx, __fmul_rn(uy, cy)));
    float b3 = __fmaf_rn(uy, cx, -__fmul_rn(ux, cy));
    if (!valid) { a3 = -INFINITY; b3 = 0.f; }
    float a3h, a3l, b3h, b3l;
    tf32_split(a3, a3h, a3l); tf32_split(b3, b3h, b3l);
    dst[2] = make_float4(a3h, b3h, a3l, b3l);
  }
}

// the reference expression for pixel t against hypothesis h, everything fetched from global memory
__device__ __forceinline__ bool vote_exact_at(const VmCtx& c, int t, int h) {
  if (t >= c.t_end || h >= c.HN) return false;          // padding record / padding hypothesis
  const uint32_t q = __ldg(c.fp + t);
  const float2 d = __ldg(c.dir + t);
  return vote_exact((float)(q & 0xffff), (float)(q >> 16), d.x, d.y, dir_norm(d.x, d.y), __ldg(c.hypx + h),
                    __ldg(c.hypy + h), c.T);
}

// 256 threads at 128 registers: two CTAs fill the register file of an SM (16 warps, 4 per scheduler).  A separate
// producer warp would make 9 warps, which the register file allocates as 12 (granularity 4): one CTA per SM at 112
// registers or two at 80 with 30+ spilled registers -- both measured slower (profiles/r2_vote_mma_history.md).
// So every warp is a consumer AND converts 1/8 of each raw tile into records, one tile ahead of its own MMA
// work; the conversion's long dependency chain (normalise, scale, split) fills issue slots the MMA epilogue
// leaves empty.
__global__ void __launch_bounds__(VM_THREADS, 2)
vote_mma_kernel(epb_voting_params p, Workspace ws, VoteConsts vc, int item_px, int chunks) {
  __shared__ VmSmem sm;
  const int HN = p.hn * p.rounds;
  const int B = p.B;
  const int* __restrict__ item_off = ws.item_off;
  const long long total = (long long)item_off[B] * p.vn * chunks;
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
  const int ntiles = (t_end - t_begin + VM_TILE - 1) / VM_TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t* fp = ws.fgpix + (size_t)b * ws.cap;
  const float2* dir = ws.direct + ((size_t)b * p.vn + v) * ws.cap;
  const float* hypx = hyp_plane(ws, b, p.vn, v, HN, 0);
  const float* hypy = hyp_plane(ws, b, p.vn, v, HN, 1);
#ifdef EPB_TUNING
  const int dbg = vc.fast_ok >> 8;   // bisecting knobs (EPB_VM_DEBUG): 1 no consumer math, 2 no record math, 4 no rare path
#else
  constexpr int dbg = 0;
#endif

  // TMA: lane 0 of warp 0 streams the raw tiles (128 x float2 direction + 128 x packed pixel) global -> shared.
  // Whole tiles: the workspace rows are padded to VM_TILE pixels.
  auto tma_issue = [&](int i) {
    const int rs = i % VM_RAW_STAGES;
    mbar_expect_tx(&sm.raw_full[rs], VM_TILE * 12);
    bulk_g2s(sm.raw_dir[rs], dir + t_begin + (size_t)i * VM_TILE, VM_TILE * 8, &sm.raw_full[rs]);
    bulk_g2s(sm.raw_pix[rs], fp + t_begin + (size_t)i * VM_TILE, VM_TILE * 4, &sm.raw_full[rs]);
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < VM_RAW_STAGES; ++i) { mbar_init(&sm.raw_full[i], 1); mbar_init(&sm.raw_empty[i], VM_WARPS / 2); }
#pragma unroll
    for (int i = 0; i < VM_REC_STAGES; ++i) { mbar_init(&sm.rec_full[i], VM_WARPS / 2); mbar_init(&sm.rec_empty[i], VM_WARPS); }
    sm.qn = 0u; sm.overflow = 0u;
    sm.ctx.fp = fp; sm.ctx.dir = dir; sm.ctx.hypx = hypx; sm.ctx.hypy = hypy;
    sm.ctx.out = ws.counts + ((size_t)b * p.vn + v) * HN;
    sm.ctx.t_begin = t_begin; sm.ctx.t_end = t_end; sm.ctx.h0 = chunk * VM_CHUNK; sm.ctx.HN = HN;
    sm.ctx.T = p.inlier_thresh;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < min(VM_RAW_STAGES, ntiles); ++i) tma_issue(i);
  }
  __syncthreads();

  // convert(i): raw tile i -> record stage i % VM_REC_STAGES.  The four warps whose parity matches the tile's
  // convert 32 records each (one per lane); the other four skip, so every warp converts every other tile.
  auto convert = [&](int i) {
    if (((warp ^ i) & 1) != 0) return;
    const int rs = i % VM_RAW_STAGES, s = i % VM_REC_STAGES;
    mbar_wait(&sm.raw_full[rs], (unsigned)(i / VM_RAW_STAGES) & 1u);
    if (i >= VM_REC_STAGES) mbar_wait(&sm.rec_empty[s], (unsigned)(i / VM_REC_STAGES - 1) & 1u);
    const int r = (warp >> 1) * 32 + lane;
    const float2 d = sm.raw_dir[rs][r];
    float4* dst = sm.rec[s] + r * 4;          // record r = group r / 8, row g = r % 8: four fragment lanes j
    if (dbg & 2) { dst[0] = dst[1] = dst[3] = make_float4(d.x, d.y, 0.f, 0.f); dst[2] = make_float4(-INFINITY, 0.f, 0.f, 0.f); }
    else make_record(sm.raw_pix[rs][r], d.x, d.y, t_begin + i * VM_TILE + r < t_end, vc, dst);
    __syncwarp();     // every lane has read raw stage rs and written its record
    if (lane == 0) { mbar_arrive(&sm.rec_full[s]); mbar_arrive(&sm.raw_empty[rs]); }
  };

  {
    const int g = lane >> 2, j = lane & 3;
    const int hbase = chunk * VM_CHUNK + warp * (VM_NB * 8);
    const bool active = hbase < HN && !(dbg & 1);    // warp-uniform
    // B fragments (n = g: hypothesis hbase + 8 nb + g, k = j / j + 4) and, for the two hypotheses whose columns
    // this lane reads from the accumulators (hbase + 8 nb + 2 j + {0, 1}), the absolute part EH of the band
    float fb0[VM_NB], fb1[VM_NB], eh[VM_NB][2];
    unsigned neg[VM_NB][2];
#pragma unroll
    for (int nb = 0; nb < VM_NB; ++nb) {
      const int h = hbase + 8 * nb + g;
      float hx = 0.f, hy = 0.f;
      if (h < HN) { hx = hypx[h]; hy = hypy[h]; }
      float hxh, hxl, hyh, hyl;
      tf32_split(hx, hxh, hxl); tf32_split(hy, hyh, hyl);
      fb0[nb] = j == 0 ? hxh : j == 1 ? hyh : j == 2 ? 1.f : hxl;
      fb1[nb] = j == 0 ? hxh : j == 1 ? hyh : j == 2 ? 1.f : hyl;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int hc = hbase + 8 * nb + 2 * j + c;
        float cx = 0.f, cy = 0.f;
        if (hc < HN) { cx = hypx[hc]; cy = hypy[hc]; }
        // non-finite or huge hypotheses: undecided against every pixel -> reference expression
        const float habs = __fadd_ru(fabsf(cx), fabsf(cy));
        eh[nb][c] = (habs <= 1e15f) ? __fmaf_ru(vc.eh_scale, habs, vc.eh_abs) : INFINITY;
        neg[nb][c] = 0u;
      }
    }
    const float gam = vc.gamma;
    const int h_rel_lane = warp * (VM_NB * 8) + 2 * j;
    const uint32_t qn_addr = smem_u32(&sm.qn);
    unsigned* const q_ptr = sm.q;
    static_assert((VM_QCAP & (VM_QCAP - 1)) == 0, "the queue index wraps with a mask");

    // One TRIP = the 8 pair tests of a lane on 16 records x 1 n-block pair... see below: 2 MMA tiles (records g and
    // g + 8) x 2 n-blocks.  issue(): its four MMAs; finish(): m, band, queue, sign counts.  The loop below keeps the
    // MMAs of the next trip in flight while the FP32 pipe finishes the current one (two accumulator sets).
    auto issue = [&](float (&d)[4][4], const float4& A0, const float4& A1, int nb) {
      mma_tf32_k8(d[0], A0, fb0[nb], fb1[nb]);
      mma_tf32_k8(d[1], A1, fb0[nb], fb1[nb]);
      mma_tf32_k8(d[2], A0, fb0[nb + 1], fb1[nb + 1]);
      mma_tf32_k8(d[3], A1, fb0[nb + 1], fb1[nb + 1]);
    };
    // finish(): BRANCH-FREE.  A lane-trip with an undecided pair (~0.5 % of them) does not count its 8 pairs at
    // all (the sign accumulation is predicated off, i.e. they stand as 8 inliers) and appends itself to the queue;
    // after the tile loop the queued lane-trips are re-evaluated with the reference expression, one pair per
    // thread, and every pair that is NOT an inlier takes its vote back.  No divergent branch, no reconvergence
    // before the next mma.sync: the whole step is one basic block that ptxas can interleave freely.
    auto finish = [&](const float (&d)[4][4], int t_rel0, int nb) {
      // pair e: bit 0 = column (hypothesis 2j / 2j + 1), bit 1 = record (g / g + 8), bit 2 = n-block
      float m[8], w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float (&dd)[4] = d[e >> 1];
        m[e] = __fsub_rn(dd[e & 1], fabsf(dd[2 + (e & 1)]));              // a' - |p|
        w[e] = __fmaf_rn(gam, dd[e & 1], eh[nb + (e >> 2)][e & 1]);         // gamma a' + EH[h]
      }
      bool clear = (fabsf(m[0]) > w[0]);
#pragma unroll
      for (int e = 1; e < 8; ++e) clear = clear & (fabsf(m[e]) > w[e]);
      if (dbg & 4) clear = true;
      // sign counts of a clear trip (predicated adds), queue push of an unclear one (predicated atomic + store;
      // the queue index wraps, an overflow shows as qn > VM_QCAP afterwards)
#if VM_PREDICATED
      // everything predicated inside one asm block: no branch, no reconvergence point before the next mma.sync
      asm volatile(
          "{ .reg .pred c; .reg .u32 t, pos, a;\n"
          "  setp.ne.u32 c, %4, 0;\n"
          "  shr.u32 t, %5, 31;\n  @c add.u32 %0, %0, t;\n  shr.u32 t, %6, 31;\n  @c add.u32 %1, %1, t;\n"
          "  shr.u32 t, %7, 31;\n  @c add.u32 %0, %0, t;\n  shr.u32 t, %8, 31;\n  @c add.u32 %1, %1, t;\n"
          "  shr.u32 t, %9, 31;\n  @c add.u32 %2, %2, t;\n  shr.u32 t, %10, 31;\n @c add.u32 %3, %3, t;\n"
          "  shr.u32 t, %11, 31;\n @c add.u32 %2, %2, t;\n  shr.u32 t, %12, 31;\n @c add.u32 %3, %3, t;\n"
          "  @!c atom.shared.add.u32 pos, [%13], 1;\n"
          "  @!c shl.b32 a, pos, 2;\n  @!c and.b32 a, a, %16;\n  @!c add.u32 a, a, %14;\n"
          "  @!c st.shared.u32 [a], %15;\n"
          "}\n"
          : "+r"(neg[nb][0]), "+r"(neg[nb][1]), "+r"(neg[nb + 1][0]), "+r"(neg[nb + 1][1])
          : "r"((unsigned)clear), "r"(__float_as_uint(m[0])), "r"(__float_as_uint(m[1])), "r"(__float_as_uint(m[2])),
            "r"(__float_as_uint(m[3])), "r"(__float_as_uint(m[4])), "r"(__float_as_uint(m[5])), "r"(__float_as_uint(m[6])),
            "r"(__float_as_uint(m[7])), "r"(qn_addr), "r"(smem_u32(q_ptr)),
            "r"(((unsigned)t_rel0 << 9) | (unsigned)(h_rel_lane + 8 * nb)), "r"((unsigned)(VM_QCAP * 4 - 1))
          : "memory");
#else
      if (clear) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          asm("{ .reg .u32 t; shr.u32 t, %1, 31; add.u32 %0, %0, t; }" : "+r"(neg[nb + (e >> 2)][e & 1]) : "r"(__float_as_uint(m[e])));
      } else {
        unsigned pos;
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(qn_addr) : "memory");
        q_ptr[pos & (unsigned)(VM_QCAP - 1)] = ((unsigned)t_rel0 << 9) | (unsigned)(h_rel_lane + 8 * nb);
      }
#endif
    };

    constexpr int TRIPS = VM_NB / 2;       // per 16-record step
    convert(0);
    for (int i = 0; i < ntiles; ++i) {
      const int s = i % VM_REC_STAGES;
      if (i + 1 < ntiles) convert(i + 1);           // one tile ahead of this warp's own MMA work
      mbar_wait(&sm.rec_full[s], (unsigned)(i / VM_REC_STAGES) & 1u);
      // every warp has converted tile i, i.e. raw stage i % VM_RAW_STAGES is free: refill it (never blocks)
      if (threadIdx.x == 0 && i + VM_RAW_STAGES < ntiles) {
        mbar_wait(&sm.raw_empty[i % VM_RAW_STAGES], (unsigned)(i / VM_RAW_STAGES) & 1u);
        tma_issue(i + VM_RAW_STAGES);
      }
      __syncwarp();
      if (active) {
        const float4* rec = sm.rec[s] + lane;
        int t_rel0 = i * VM_TILE + g;
        float da[4][4], db[4][4];
        float4 A0 = rec[0], A1 = rec[32];
        issue(da, A0, A1, 0);
#pragma unroll 1
        for (int st = 0; st < VM_TILE / 16; ++st, t_rel0 += 16) {
          // software pipeline over the trips of this step and the first trip of the next one
          static_assert(TRIPS == 2, "the pipeline below is written for two trips per step");
          issue(db, A0, A1, 2);
          const float4* nx = rec + ((st + 1 < VM_TILE / 16) ? (st + 1) * 64 : st * 64);   // (last step: reload, unused)
          const float4 N0 = nx[0], N1 = nx[32];
          finish(da, t_rel0, 0);
          if (st + 1 < VM_TILE / 16) issue(da, N0, N1, 0);
          finish(db, t_rel0, 2);
          A0 = N0; A1 = N1;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.rec_empty[s]);
    }
    __syncthreads();   // every enqueue precedes this barrier
    const unsigned nq_raw = sm.qn;
    int32_t* out = ws.counts + ((size_t)b * p.vn + v) * HN;
    if (nq_raw <= (unsigned)VM_QCAP) {
      if (active) {
        // fast counts: every record visited beyond t_end is a padding record (m = -inf: sign counted, or its
        // lane-trip queued), every pair of a queued lane-trip stands as an inlier until the pass below
        const int visited = ntiles * VM_TILE;
#pragma unroll
        for (int nb = 0; nb < VM_NB; ++nb)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            unsigned n = neg[nb][c];
            n += __shfl_xor_sync(FULL, n, 4);
            n += __shfl_xor_sync(FULL, n, 8);
            n += __shfl_xor_sync(FULL, n, 16);
            const int h = hbase + 8 * nb + 2 * j + c;
            const int cnt = visited - (int)n;
            if (g == 0 && h < HN && cnt != 0) atomicAdd(out + h, cnt);
          }
      }
      // queued lane-trips: their 8 pairs with the reference expression, one pair per thread; a pair that is not an
      // inlier (or a padding record / padding hypothesis) takes back the vote it was given
      const VmCtx& c = sm.ctx;
      for (unsigned k = threadIdx.x; k < nq_raw * 8u; k += VM_THREADS) {
        const unsigned ent = sm.q[k >> 3], e = k & 7u;
        const int t = t_begin + (int)(ent >> 9) + ((e & 2u) ? 8 : 0);
        const int h = c.h0 + (int)(ent & 511u) + ((e & 4u) ? 8 : 0) + (int)(e & 1u);
        if (h < HN && !vote_exact_at(c, t, h)) atomicAdd(out + h, -1);
      }
    } else {
      // the queue overflowed (non-finite fields or hypotheses, thresholds with a wide band): recount the whole unit
      // with the reference expression, one thread per hypothesis
      const VmCtx& c = sm.ctx;
      for (int hr = threadIdx.x; hr < VM_CHUNK; hr += VM_THREADS) {
        const int h = c.h0 + hr;
        if (h >= HN) continue;
        int cnt = 0;
        for (int t = t_begin; t < t_end; ++t) cnt += vote_exact_at(c, t, h);
        if (cnt) atomicAdd(out + h, cnt);
      }
    }
  }
}


// Work units of the vote kernels: (item of <= item_px voting pixels, keypoint, hypothesis chunk), listed on the
// device (vote_items_kernel) so that ragged batches balance without a host sync; CTAs beyond the list exit at once.
static int launch_vote_mma(const epb_voting_params& p, const Workspace& ws, const VoteConsts& vc, cudaStream_t s) {
  const int HN = p.hn * p.rounds;
  const int chunks = (HN + VM_CHUNK - 1) / VM_CHUNK;
  const long long slots = 2LL * device_sm_count();          // two 8-warp CTAs per SM (128 registers each)
  // items small enough that the work list is several waves long, large enough to amortise a unit's prologue
  // (64 B-fragment registers per lane, barrier set-up) and its tail (queue resolution, 16 count atomics per lane)
  const int tn_max = ws.cap;
  int item_px = 16 * VM_TILE;
  while (item_px > VM_TILE && (long long)p.B * p.vn * chunks * ((tn_max + item_px - 1) / item_px) < 4 * slots) item_px >>= 1;
  { const int forced = tuning_int("EPB_VOTE_ITEM", 0); if (forced >= VM_TILE && forced % VM_TILE == 0) item_px = forced; }
  const long long max_units = (long long)p.B * p.vn * chunks * ((tn_max + item_px - 1) / item_px);
  if (max_units > 0x7fffffffLL) return EPB_ERR_INVALID;
  vote_items_kernel<<<1, 256, 0, s>>>(p.B, item_px, ws);
  EPB_RETURN_IF(check_launch());
  EPB_RETURN_IF(check_api(cudaMemsetAsync(ws.counts, 0, (size_t)p.B * p.vn * HN * 4, s)));
  vote_mma_kernel<<<(unsigned)max_units, VM_THREADS, 0, s>>>(p, ws, vc, item_px, chunks);
  return check_launch();
}

