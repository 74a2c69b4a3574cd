// mn_edge.cuh -- edge construction for the merge segmenter (sm_100a).
//
// Kernel 1  mn_edge_pass_kernel   (HBM-bound streaming pass; THE roofline kernel)
//   reference: Object ctor cc:5-21 (+ h:295-297) and the (pixel,offset) AdjacencyRecord ctor
//   cc:24-36, i.e. per pixel  clp[c] = logf(class[c,p]), cls = first argmax, and per pixel x offset
//   same = logf(s), diff = (float)log(1.0 - (double)s) with s read at the SOURCE pixel (h:301-303).
//   A tile of TP consecutive pixels needs one contiguous TP*4-byte segment of each of the C+K input
//   planes and nothing else (no halo: the neighbour contributes no probability), so each
//   plane segment is brought to shared memory by ONE 1-D TMA bulk copy (cp.async.bulk + mbarrier),
//   double buffered; every input byte is read exactly once for all K offsets, including the large
//   occlusion offsets.  Results are staged in shared memory and leave as TMA bulk stores.
//   Algorithmic bytes per pixel: 4*(C+K) read + 4*(C+2K) written (+4 for cls).
//
// Kernel 2  mn_record_init_kernel
//   reference: cc:209-231 + SortAndUpdateHash cc:49-56 + UpdateMergePriority cc:145-150.
//   Per record slot r = pixel*K + k: bounds test, endpoints (lo,hi), oml = same - diff, the initial
//   merge priority (needs the neighbour's class / class vector -> separate pass after kernel 1),
//   the (lo,hi)->record hash insert, the per-pixel live-record bit masks, and the 64-bit sort key
//   of the initial queue entry (sentinel when mp < 0 or the slot does not exist, cc:225-227).
#pragma once
#include <cuda.h>  // CUtensorMap (type only: the encoder is fetched from the driver at run time, no -lcuda)
#include <cuda_runtime.h>

#include "mn_common.h"
#include "mn_layout.h"
#include "mn_log1m_tab.h"

// ---- PTX helpers: mbarrier + 1-D TMA bulk copies --------------------------------------------
__device__ __forceinline__ uint32_t mn_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mn_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mn_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mn_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mn_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mn_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(mn_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void mn_tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes,
                                               uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          mn_smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(mn_smem_u32(bar))
      : "memory");
}
// 2-D tiled TMA load: one box (inner extent x rows) of a [rows][inner] tensor described by a tensor map; rows and
// columns beyond the tensor arrive as zeros and still count towards the barrier's transaction bytes
__device__ __forceinline__ void mn_tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c_inner, int c_row, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          mn_smem_u32(smem_dst)),
      "l"(tmap), "r"(c_inner), "r"(c_row), "r"(mn_smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mn_tma_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(mn_smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mn_tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void mn_tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void mn_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// diff = (float)log(1.0 - (double)s)   (cc:34).  1.0 - s is exact in fp64 for s >= 2^-23.
__device__ __forceinline__ float mn_log1m_exact(float s) { return (float)log(1.0 - (double)s); }

// The edge pass is bound by the SM's fp64 pipe unless these two are lean (B200: 64 fp64 lanes / SM):
//
// logf as glibc's FMA build evaluates it (the recipe of mn_logf_exact with its five multiply-adds
// fused; bit-identical on the whole clipped domain, pinned exhaustively on the device): 6 fp64 ops.
// polynomial constants live in constant memory: a DFMA takes one operand straight from the constant
// bank, whereas an immediate double costs two MOVs per use inside the (register-tight) plane loops
__constant__ double mn_kc[12] = {
    0x1.62e42fefa39efp-1,    // 0: Ln2
    0x1.5575b0be00b6ap-2,    // 1: logf A1
    -0x1.ffffef20a4123p-2,   // 2: logf A2
    -0x1.00ea348b88334p-2,   // 3: logf A0
    -1.0 / 6, 0.2, -0.25, 1.0 / 3, -0.5,  // 4..8: log1p polynomial of log1m
    -1.0, 1.0, 0.0};
// (conversions run on the quarter-rate XU pipe, so the exact widenings are done with integer ops)
__device__ __forceinline__ double mn_f32bits_to_f64(uint32_t b) {  // b = bits of a positive normal float
  return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}
__device__ __forceinline__ double mn_small_int_to_f64(int k) { return (double)k; }  // (one XU op; exact)
__device__ __forceinline__ float mn_logf_fast(float x, const MnLogfTab* tab) {
  const uint32_t ix = __float_as_uint(x);
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) return mn_logf_exact(x, tab);  // 0, negative, subnormal, inf, NaN
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (tmp >> 19) & 15;
  const int k = (int32_t)tmp >> 23;
  const double z = mn_f32bits_to_f64(ix - (tmp & 0xff800000u));
  const double2 e = *reinterpret_cast<const double2*>(&tab[i]);  // (invc, logc) in one 16-byte load
  const double r = __fma_rn(z, e.x, -1.0);
  const double y0 = __fma_rn(mn_small_int_to_f64(k), mn_kc[0], e.y);
  const double r2 = __dmul_rn(r, r);
  double y = __fma_rn(r, mn_kc[1], mn_kc[2]);
  y = __fma_rn(r2, mn_kc[3], y);
  y = __fma_rn(y, r2, __dadd_rn(y0, r));
  return (float)y;
}
// (float)log(x), x = 1 - s in fp64: a 128-bin table (tools/gen_log1m_table.py) and a degree-6 log1p
// polynomial give log(x) to ~2^-45 relative in 10 fp64 ops.  The float rounding of that value equals
// the float rounding of libm's log(x) unless the value lies within 2^-38 (relative) of a rounding
// boundary; only then (6e-5 of all inputs) the full-precision log decides.  Verified against the host
// over all 192,937,983 inputs (0 differences outside the fallback set) and pinned on the device by
// tests/test_libm_parity.py.
struct MnLog1mTab {
  double invc, logc;
};
__constant__ MnLogfTab mn_logf_table_c[16] = {MN_LOGF_TABLE};
__constant__ MnLog1mTab mn_log1m_table[128] = {MN_LOG1M_TABLE};
__device__ __forceinline__ float mn_log1m_fast(float s, const MnLog1mTab* tab) {
  // outside [2^-126, 1) (an unclipped caller; a biased probability rounded to 0 or 1): the libm expression itself,
  // with its special values (log(0) = -inf, log of a negative = NaN)
  if (__float_as_uint(s) - 0x00800000u >= 0x3f800000u - 0x00800000u) return (float)log(1.0 - (double)s);
  const double x = __dadd_rn(1.0, -mn_f32bits_to_f64(__float_as_uint(s)));  // exact
  const uint32_t hx = (uint32_t)__double2hiint(x);
  const uint32_t tmp = hx - (uint32_t)(MN_LOG1M_OFF >> 32);  // (the low word of OFF is zero)
  const int i = (int)((tmp >> 13) & 127u);
  const int k = (int32_t)tmp >> 20;
  const double z = __hiloint2double((int)(hx - (tmp & 0xfff00000u)), __double2loint(x));
  const double2 e = *reinterpret_cast<const double2*>(&tab[i]);
  const double r = __fma_rn(z, e.x, -1.0);
  const double t = __fma_rn(mn_small_int_to_f64(k), mn_kc[0], e.y);
  double q = __fma_rn(r, mn_kc[4], mn_kc[5]);
  q = __fma_rn(r, q, mn_kc[6]);
  q = __fma_rn(r, q, mn_kc[7]);
  q = __fma_rn(r, q, mn_kc[8]);
  const double r2 = __dmul_rn(r, r);
  double y = __fma_rn(r2, q, r);
  y = __dadd_rn(y, t);
  const int d = abs((int)((uint32_t)__double2loint(y) & 0x1fffffffu) - 0x10000000);
  if (d < (1 << 14)) return (float)log(x);  // too close to a float rounding boundary: decide exactly
  return (float)y;
}

// glibc-exact expf for |x| < 88 (no overflow / underflow handling: the biased logit of a clipped
// probability is far inside): the public glibc / ARM optimized-routines algorithm (N = 32 table,
// degree-3 polynomial, all in fp64).  The x86-64 libm selects its FMA build, so the three
// polynomial steps are fused exactly as that build fuses them; with separate multiply-adds 2 of the
// 1.78e9 inputs in [-80, 80] differ.  Pinned to the host libm by tests/test_libm_parity.py.
__constant__ unsigned long long mn_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
__device__ __forceinline__ float mn_expf_exact(float x) {
  if (!(fabsf(x) < 87.0f)) return expf(x);  // overflow / underflow / NaN: the special values (a saturated logit)
  const double InvLn2N = 0x1.71547652b82fep+0 * 32.0, SHIFT = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / 32768.0, C1 = 0x1.ebfce50fac4f3p-3 / 1024.0, C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
  double z = __dmul_rn(InvLn2N, (double)x);
  double kd = __dadd_rn(z, SHIFT);
  unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
  kd = __dadd_rn(kd, -SHIFT);
  double r = __dadd_rn(z, -kd);
  unsigned long long t = mn_exp2f_tab[ki & 31ull] + (ki << 47);
  double s = __longlong_as_double((long long)t);
  z = __fma_rn(C0, r, C1);
  double r2 = __dmul_rn(r, r);
  double y = __fma_rn(C2, r, 1.0);
  y = __fma_rn(z, r2, y);
  y = __dmul_rn(y, s);
  return (float)y;
}

// same_different_bias transform (cc:183-195), evaluated like the reference: logf + fp64 log,
// rounded to float, expf, then 1/(1+e) in fp64 rounded to float.
__device__ __forceinline__ float mn_bias_sameness(float s, float sdb, const MnLogfTab* tab) {
  float logit = (float)(((double)mn_logf_exact(s, tab) - log(1.0 - (double)s)) + (double)sdb);
  return (float)(1.0 / (1.0 + (double)mn_expf_exact(-logit)));
}

// F.sigmoid as torch evaluates it on the device in fp32: 1 / (1 + exp(-x)), IEEE division (utils/inference_utils.py:43-44)
__device__ __forceinline__ float mn_sigmoid_f32(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

struct MnEdgeParams {
  const float* class_pred;  // [B][C][N]
  const float* adj_pred;    // [B][K][N]
  float* adj_pred_rw;       // same buffer, written in place when sdb != 0 (cc:187-191), else null
  const MnImage* imgs;      // per-image workspace: clp [N][C], cls [N], rec_same / rec_diff [N*K]
  int B, C, K, N;
  int TP;                   // pixels per tile (multiple of 4)
  int tiles_per_image;
  int use_tma;              // N % 4 == 0 and 16-byte aligned bases
  int use_tmap;             // tile kernel: the input planes of a tile arrive as TWO 2-D tensor-map boxes (class planes x
                            // pixels, offset planes x pixels) instead of C + K 1-D bulk copies (needs use_tma, TP % 32 == 0)
  int clip;                 // apply the wrapper's clip to [2^-23, 1-2^-23] (c_segment.pyx:53-55)
  int logits;               // inputs are the network's logits: sigmoid first (inference_utils.py:43-44,95-96), then clip
  float sdb;
  // warp-pipeline kernel: image b's outputs live at ws0_* + b * ws_stride (no pointer loads per tile)
  float *ws0_clp, *ws0_same, *ws0_diff;
  int* ws0_cls;
  size_t ws_stride;  // bytes
  int stages;        // input ring depth (2..4)
  const int* only_if;  // tile kernel: when set, do nothing unless *only_if != 0 (the fix-up launch after the domain check)
};

// Raw C-ABI callers (clip = 0) may hand in values outside [2^-126, 1): the warp-pipeline kernel's log recipes assume
// that range.  One streaming pass raises a flag; the general kernel (whose logs honour libm's special values) then
// redoes the batch -- launched right behind, it returns at once when the flag is down (no host round trip).
__global__ void __launch_bounds__(256) mn_domain_check_kernel(const float* a, size_t na, const float* b, size_t nb, int* flag) {
  unsigned bad = 0;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  const uint4* a4 = reinterpret_cast<const uint4*>(a);
  const uint4* b4 = reinterpret_cast<const uint4*>(b);
  for (size_t i = tid; i < na / 4; i += nth) {
    const uint4 v = a4[i];
    bad |= (v.x - 0x00800000u >= 0x3f000000u) | (v.y - 0x00800000u >= 0x3f000000u) | (v.z - 0x00800000u >= 0x3f000000u) | (v.w - 0x00800000u >= 0x3f000000u);
  }
  for (size_t i = tid; i < nb / 4; i += nth) {
    const uint4 v = b4[i];
    bad |= (v.x - 0x00800000u >= 0x3f000000u) | (v.y - 0x00800000u >= 0x3f000000u) | (v.z - 0x00800000u >= 0x3f000000u) | (v.w - 0x00800000u >= 0x3f000000u);
  }
  if (__any_sync(0xffffffffu, bad != 0) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// dynamic smem layout: [bar0, bar1][pad to 128][in0][in1][out0: clp|same|diff][out1][logf tab][log1m tab]
// two CTAs of 512 threads per SM (tiles of 256 pixels, ~100 KB of shared memory each): while one waits
// at a barrier or on its TMA loads the other computes
#define MN_EDGE_THREADS 512
#define MN_EDGE_CTAS_PER_SM 2
__global__ void __launch_bounds__(MN_EDGE_THREADS, MN_EDGE_CTAS_PER_SM) mn_edge_pass_kernel(MnEdgeParams P, const __grid_constant__ CUtensorMap tm_class,
                                                                                            const __grid_constant__ CUtensorMap tm_adj) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (P.only_if && *P.only_if == 0) return;
  const int C = P.C, K = P.K, TP = P.TP, NPL = C + K;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* in0 = reinterpret_cast<float*>(smem_raw + 128);
  float* in1 = in0 + (size_t)NPL * TP;
  float* out_base = in1 + (size_t)NPL * TP;  // two output stages of (C + 2K) * TP floats
  const size_t out_stride = (size_t)(C + 2 * K) * TP;
  MnLogfTab* tab = reinterpret_cast<MnLogfTab*>(out_base + 2 * out_stride);
  MnLog1mTab* tab1m = reinterpret_cast<MnLog1mTab*>(tab + 16);

  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid < 16) {
    const MnLogfTab t16[16] = {MN_LOGF_TABLE};
    tab[tid] = t16[tid];
  }
  if (tid < 128) tab1m[tid] = mn_log1m_table[tid];
  if (tid == 0) {
    mn_mbar_init(&bars[0], 1);
    mn_mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int total_tiles = P.B * P.tiles_per_image;  // (checked < 2^31 by the host)
  auto issue = [&](int tile, int stage) {
    // one elected thread: arm the barrier, then one bulk copy per input plane
    int b = (int)(tile / P.tiles_per_image);
    int start = (int)(tile % P.tiles_per_image) * TP;
    int tl = min(TP, P.N - start);
    float* dst = stage ? in1 : in0;
    const float* cbase = P.class_pred + ((size_t)b * C) * P.N + start;
    const float* abase = P.adj_pred + ((size_t)b * K) * P.N + start;
    if (P.use_tmap) {  // two boxes: [C planes][TP pixels] and [K planes][TP pixels] (the tail of a last tile: zeros)
      mn_mbar_expect_tx(&bars[stage], (uint32_t)(NPL * TP * 4));
      mn_tma_load_2d(dst, &tm_class, start, b * C, &bars[stage]);
      mn_tma_load_2d(dst + (size_t)C * TP, &tm_adj, start, b * K, &bars[stage]);
      return;
    }
    mn_mbar_expect_tx(&bars[stage], (uint32_t)(NPL * tl * 4));
    for (int pl = 0; pl < C; pl++)
      mn_tma_load_1d(dst + (size_t)pl * TP, cbase + (size_t)pl * P.N, tl * 4, &bars[stage]);
    for (int pl = 0; pl < K; pl++)
      mn_tma_load_1d(dst + (size_t)(C + pl) * TP, abase + (size_t)pl * P.N, tl * 4, &bars[stage]);
  };

  uint32_t phase[2] = {0, 0};
  int tile = blockIdx.x;
  if (P.use_tma && tid == 0 && tile < total_tiles) issue(tile, 0);
  int stage = 0;
  for (; tile < total_tiles; tile += gridDim.x, stage ^= 1) {
    const int b = (int)(tile / P.tiles_per_image);
    const int start = (int)(tile % P.tiles_per_image) * TP;
    const int tl = min(TP, P.N - start);
    float* in = stage ? in1 : in0;
    float* out_clp = out_base + (stage ? out_stride : 0);
    float* out_same = out_clp + (size_t)C * TP;
    float* out_diff = out_same + (size_t)K * TP;
    float* im_clp = P.imgs[b].clp;
    int* im_cls = P.imgs[b].cls;
    float* im_same = P.imgs[b].rec_same;
    float* im_diff = P.imgs[b].rec_diff;
    if (P.use_tma) {
      mn_mbar_wait(&bars[stage], phase[stage]);
      phase[stage] ^= 1;
    } else {
      const float* cbase = P.class_pred + ((size_t)b * C) * P.N + start;
      const float* abase = P.adj_pred + ((size_t)b * K) * P.N + start;
      for (int i = tid; i < NPL * tl; i += nt) {
        int pl = i / tl, px = i - pl * tl;
        in[(size_t)pl * TP + px] =
            pl < C ? cbase[(size_t)pl * P.N + px] : abase[(size_t)(pl - C) * P.N + px];
      }
      __syncthreads();
    }
    // the bulk stores issued two tiles ago (same output stage) must have finished READING it; the
    // stores of the previous tile keep draining while this tile is computed
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();
    if (P.use_tma) {
      // prefetch the next tile into the other input stage: its readers finished before the barrier
      // above, and only the issuing thread is held up by the 19 bulk-copy instructions
      int nxt = tile + (int)gridDim.x;
      if (tid == 0 && nxt < total_tiles) {
        mn_fence_proxy_async();
        issue(nxt, stage ^ 1);
      }
    }

    // ---- compute: one item per (plane, pixel); consecutive lanes -> consecutive pixels.  In a full
    //      tile (a power of two of pixels) a thread keeps its pixel and walks the planes: no index
    //      arithmetic beyond pointer increments ----
    bool cls_done = false;
    if (tl == TP && nt == 2 * TP && P.sdb == 0.0f) {
      // two thread groups per tile: group 0 takes every class plane of its pixel (first-argmax kept in
      // registers, cc:18-20) and the last a0 offset planes, group 1 the other offset planes (an offset
      // plane costs about 2.1 class planes) -- no separate argmax pass, no barrier in between
      const int px = tid & (TP - 1), g = tid >= TP ? 1 : 0;
      int a0 = (21 * K - 10 * C) / 42;
      a0 = a0 < 0 ? 0 : (a0 > K ? K : a0);
      if (g == 0) {
        float best = 0.0f; int bc = 0;
        for (int pl = 0; pl < C; pl++) {
          float v = in[(size_t)pl * TP + px];
          if (P.logits) v = mn_sigmoid_f32(v);
          if (P.clip) v = fminf(fmaxf(v, 1.1920929e-07f), 0.99999988f);
          const float l = MN_FADD(0.0f, mn_logf_fast(v, tab));  // cc:11-16
          out_clp[px * C + pl] = l;
          if (pl == 0 || l > best) { best = l; bc = pl; }
        }
        im_cls[start + px] = bc;
      }
      const int k0 = g == 0 ? K - a0 : 0, k1 = g == 0 ? K : K - a0;
      for (int k = k0; k < k1; k++) {
        float v = in[(size_t)(C + k) * TP + px];
        if (P.logits) v = mn_sigmoid_f32(v);
          if (P.clip) v = fminf(fmaxf(v, 1.1920929e-07f), 0.99999988f);
        out_same[px * K + k] = mn_logf_fast(v, tab);     // cc:35
        out_diff[px * K + k] = mn_log1m_fast(v, tab1m);  // cc:34
      }
      cls_done = true;
    } else if (tl == TP && (TP & (TP - 1)) == 0 && nt >= TP) {
      const int sh = 31 - __clz(TP);
      const int px = tid & (TP - 1), g = tid >> sh, ng = nt >> sh;
      for (int pl = g; pl < C; pl += ng) {
        float v = in[(size_t)pl * TP + px];
        if (P.logits) v = mn_sigmoid_f32(v);
          if (P.clip) v = fminf(fmaxf(v, 1.1920929e-07f), 0.99999988f);
        out_clp[px * C + pl] = MN_FADD(0.0f, mn_logf_fast(v, tab));  // cc:11-16
      }
      for (int k = g; k < K; k += ng) {
        float v = in[(size_t)(C + k) * TP + px];
        if (P.logits) v = mn_sigmoid_f32(v);
          if (P.clip) v = fminf(fmaxf(v, 1.1920929e-07f), 0.99999988f);
        if (P.sdb != 0.0f) {  // cc:183-195, in place on the caller's buffer
          v = mn_bias_sameness(v, P.sdb, tab);
          P.adj_pred_rw[((size_t)b * K + k) * P.N + start + px] = v;
        }
        out_same[px * K + k] = mn_logf_fast(v, tab);     // cc:35
        out_diff[px * K + k] = mn_log1m_fast(v, tab1m);  // cc:34
      }
    } else {
      const int items = NPL * tl;
      for (int i = tid; i < items; i += nt) {
        const int pl = i / tl, px = i - pl * tl;
        float v = in[(size_t)pl * TP + px];
        if (P.logits) v = mn_sigmoid_f32(v);
          if (P.clip) v = fminf(fmaxf(v, 1.1920929e-07f), 0.99999988f);
        if (pl < C) {
          out_clp[px * C + pl] = MN_FADD(0.0f, mn_logf_fast(v, tab));  // cc:11-16
        } else {
          int k = pl - C;
          if (P.sdb != 0.0f) {  // cc:183-195, in place on the caller's buffer
            v = mn_bias_sameness(v, P.sdb, tab);
            P.adj_pred_rw[((size_t)b * K + k) * P.N + start + px] = v;
          }
          out_same[px * K + k] = mn_logf_fast(v, tab);     // cc:35
          out_diff[px * K + k] = mn_log1m_fast(v, tab1m);  // cc:34
        }
      }
    }
    if (!cls_done) __syncthreads();
    // ---- first-argmax class per pixel (cc:18-20) ----
    for (int px = tid; !cls_done && px < tl; px += nt) {
      const float* v = out_clp + px * C;
      float best = v[0];
      int bc = 0;
      for (int c = 1; c < C; c++) {
        float x = v[c];
        if (x > best) {
          best = x;
          bc = c;
        }
      }
      im_cls[start + px] = bc;
    }
    // ---- results leave as bulk stores (contiguous in global: pixel-major tiles) ----
    float* g_clp = im_clp + (size_t)start * C;
    float* g_same = im_same + (size_t)start * K;
    float* g_diff = im_diff + (size_t)start * K;
    if (P.use_tma) {
      mn_fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        mn_tma_store_1d(g_clp, out_clp, (uint32_t)(tl * C * 4));
        mn_tma_store_1d(g_same, out_same, (uint32_t)(tl * K * 4));
        mn_tma_store_1d(g_diff, out_diff, (uint32_t)(tl * K * 4));
        mn_tma_store_commit();
      }
    } else {
      __syncthreads();
      for (int i = tid; i < tl * C; i += nt) g_clp[i] = out_clp[i];
      for (int i = tid; i < tl * K; i += nt) {
        g_same[i] = out_same[i];
        g_diff[i] = out_diff[i];
      }
    }
    if (!P.use_tma) __syncthreads();  // (TMA path: the barrier before the stores already ordered every read
                                      //  of `in`, which is refilled only after the next iteration's barrier)
  }
  if (tid == 0) mn_tma_store_wait_read();
}

// ---------------------------------------------------------------------------------------------
// Kernel 1b  mn_edge_warp_kernel: the same pass as a barrier-free warp pipeline (the fast path for
// sdb == 0 and TMA-able inputs; mn_edge_pass_kernel above stays the general path).
//   * one PRODUCER warp per CTA: an elected lane waits on the stage's `empty` mbarrier and issues the
//     C+K 1-D TMA bulk loads of the next tile (full / empty mbarrier ring, no __syncthreads anywhere
//     in the steady state);
//   * each CONSUMER warp owns 32 consecutive pixels of the tile, one pixel per lane, all C+K planes
//     (plane-major smem reads: conflict free), keeps the first-argmax in registers, stages its
//     (C+2K)*32 results in its OWN buffer and sends them off with its own three bulk stores -- a
//     warp never waits for another warp, only for data;
//   * the kernel is bound by the shared-memory data pipe once the barriers are gone: a 16-byte table
//     lookup with 32 unrelated indices costs ~9.3 wavefronts instead of 4.  Both tables are therefore
//     kept as EIGHT bank-interleaved copies (entry i, copy j at 16-byte slot 8 i + j; lane l reads copy
//     l & 7): the 8 lanes of every quarter-warp phase hit 8 different bank groups whatever their
//     indices -- exactly 4 wavefronts per lookup.  That caps the table sizes: logf keeps glibc's own
//     16 entries (2 KB), log(1 - s) uses a 64-bin table (8 KB) with the same degree-6 polynomial
//     (tests/check_log1m64.c: 0 differences outside the fallback set over the whole clipped domain);
//   * the rare exact decision of log(1 - s) (6e-5 of the values) is deferred to after the plane loops,
//     so the loops carry no call and their constants stay in registers.
struct MnEdge2Smem {
  uint64_t full[4], empty[4];
  uint64_t pad[8];
  double2 tabf[16 * 8];    // (invc, logc) of logf, 8 interleaved copies
  double2 tab1m[64 * 8];   // (invc, logc) of the 64-bin log(1 - s), 8 interleaved copies
};
#define MN_EDGE2_MAX_STAGES 4
#define MN_EDGE2_MAX_THREADS 256  // up to 7 consumer warps + the producer warp
__constant__ MnLog1mTab mn_log1m64_table[64] = {MN_LOG1M64_TABLE};

// glibc's logf recipe (the arithmetic of mn_logf_fast) on an interleaved table; lane16 = (lane & 7) * 16.
// `lo29` = bits << 29 (the low word of the widened value: shared with the log(1 - s) of the same value)
__device__ __forceinline__ float mn_logf_il(uint32_t ix, uint32_t lo29, const double2* tabf, uint32_t lane16) {
  const uint32_t tmp = ix - 0x3f330000u;
  const uint32_t off = ((tmp >> 12) & (15u << 7)) | lane16;
  const int k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  const double z = __hiloint2double((int)((iz >> 3) + 0x38000000u), (int)lo29);
  const double2 e = *reinterpret_cast<const double2*>(reinterpret_cast<const unsigned char*>(tabf) + off);
  const double r = __fma_rn(z, e.x, -1.0);
  const double y0 = __fma_rn(mn_small_int_to_f64(k), mn_kc[0], e.y);
  const double r2 = __dmul_rn(r, r);
  double y = __fma_rn(r, mn_kc[1], mn_kc[2]);
  y = __fma_rn(r2, mn_kc[3], y);
  y = __fma_rn(y, r2, __dadd_rn(y0, r));
  return (float)y;
}
// log(1 - xd) in fp64 (64-bin interleaved table); the caller tests mn_log1m_ambiguous(y)
__device__ __forceinline__ double mn_log1m_il(double xd, const double2* tab1m, uint32_t lane16) {
  const double x = __dadd_rn(1.0, -xd);  // exact
  const uint32_t hx = (uint32_t)__double2hiint(x);
  const uint32_t tmp = hx - (uint32_t)(MN_LOG1M64_OFF >> 32);
  const uint32_t off = ((tmp >> 7) & (63u << 7)) | lane16;
  const int k = (int32_t)tmp >> 20;
  const double z = __hiloint2double((int)(hx - (tmp & 0xfff00000u)), __double2loint(x));
  const double2 e = *reinterpret_cast<const double2*>(reinterpret_cast<const unsigned char*>(tab1m) + off);
  const double r = __fma_rn(z, e.x, -1.0);
  const double t = __fma_rn(mn_small_int_to_f64(k), mn_kc[0], e.y);
  double q = __fma_rn(r, mn_kc[4], mn_kc[5]);
  q = __fma_rn(r, q, mn_kc[6]);
  q = __fma_rn(r, q, mn_kc[7]);
  q = __fma_rn(r, q, mn_kc[8]);
  const double r2 = __dmul_rn(r, r);
  const double y = __fma_rn(r2, q, r);
  return __dadd_rn(y, t);
}
// the float rounding of y is ambiguous: low 29 bits within [-2^14, 2^14) of the rounding boundary 2^28
__device__ __forceinline__ bool mn_log1m_ambiguous(double y) {
  const uint32_t c = ((uint32_t)__double2loint(y) << 3) + ((0x4000u - 0x10000000u) << 3);
  return c < (0x8000u << 3);
}
__device__ __noinline__ float mn_log1m_decide(float s) { return (float)log(1.0 - (double)s); }

// CLIP: 0 = values as given, 1 = clip, 2 = logits: sigmoid, then clip
template <int CLIP>
__device__ __forceinline__ float mn_edge2_clip(float v) {
  if (CLIP == 2) v = mn_sigmoid_f32(v);
  return CLIP ? fminf(fmaxf(v, 1.1920929e-07f), 0.99999988f) : v;
}

template <int CLIP>
__global__ void __launch_bounds__(MN_EDGE2_MAX_THREADS, 3) mn_edge_warp_kernel(MnEdgeParams P) {
  __shared__ __align__(128) MnEdge2Smem S;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int C = P.C, K = P.K, TP = P.TP, NPL = C + K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncons = (int)(blockDim.x >> 5) - 1;  // TP == 32 * ncons
  float* in_base = reinterpret_cast<float*>(smem_raw);
  const int NS = P.stages;
  float* out_base = in_base + (size_t)NS * NPL * TP;
  const int OW = C + 2 * K;

  for (int e = tid; e < 16 * 8; e += blockDim.x) {
    const MnLogfTab t = mn_logf_table_c[e >> 3];
    S.tabf[e] = make_double2(t.invc, t.logc);
  }
  for (int e = tid; e < 64 * 8; e += blockDim.x) {
    const MnLog1mTab t = mn_log1m64_table[e >> 3];
    S.tab1m[e] = make_double2(t.invc, t.logc);
  }
  if (tid == 0) {
    for (int s = 0; s < NS; s++) {
      mn_mbar_init(&S.full[s], 1);
      mn_mbar_init(&S.empty[s], (uint32_t)ncons);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // tiles blockIdx.x, blockIdx.x + gridDim.x, ...: (image, tile in image) advance without divisions
  const int tpi = P.tiles_per_image;
  const int step_b = (int)gridDim.x / tpi, step_t = (int)gridDim.x % tpi;
  int b = (int)blockIdx.x / tpi, t = (int)blockIdx.x % tpi;
  if (warp == ncons) {
    // ---- producer ----
    if (lane != 0) return;
    int stage = 0, lap = 0;  // lap = number of completed passes over the ring
    for (; b < P.B;) {
      if (lap > 0) {
        mn_mbar_wait(&S.empty[stage], (uint32_t)((lap - 1) & 1));
        mn_fence_proxy_async();
      }
      const int start = t * TP;
      const int tl = min(TP, P.N - start);
      float* dst = in_base + (size_t)stage * NPL * TP;
      const float* cbase = P.class_pred + ((size_t)b * C) * P.N + start;
      const float* abase = P.adj_pred + ((size_t)b * K) * P.N + start;
      mn_mbar_expect_tx(&S.full[stage], (uint32_t)(NPL * tl * 4));
      for (int pl = 0; pl < C; pl++)
        mn_tma_load_1d(dst + (size_t)pl * TP, cbase + (size_t)pl * P.N, tl * 4, &S.full[stage]);
      for (int pl = 0; pl < K; pl++)
        mn_tma_load_1d(dst + (size_t)(C + pl) * TP, abase + (size_t)pl * P.N, tl * 4, &S.full[stage]);
      b += step_b; t += step_t;
      if (t >= tpi) { t -= tpi; b++; }
      if (++stage == NS) { stage = 0; lap++; }
    }
    return;
  }

  // ---- consumers ----
  const double2* tabf = S.tabf;
  const double2* tab1m = S.tab1m;
  const uint32_t lane16 = (uint32_t)(lane & 7) * 16u;
  float* o_clp = out_base + (size_t)warp * 32 * OW;
  float* o_same = o_clp + 32 * C;
  float* o_diff = o_same + 32 * K;
  float* __restrict__ my_clp = o_clp + lane * C;
  float* __restrict__ my_same = o_same + lane * K;
  float* __restrict__ my_diff = o_diff + lane * K;
  const int px = warp * 32 + lane;
  int stage = 0;
  uint32_t parity = 0;
  for (; b < P.B;) {
    const int start = t * TP;
    const int nvalid = min(32, P.N - start - warp * 32);  // (<= 0: this warp's slice lies beyond the image)
    const float* __restrict__ in = in_base + (size_t)stage * NPL * TP + px;
    const size_t wsb = (size_t)b * P.ws_stride;
    float* g_clp = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(P.ws0_clp) + wsb);
    float* g_same = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(P.ws0_same) + wsb);
    float* g_diff = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(P.ws0_diff) + wsb);
    int* g_cls = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(P.ws0_cls) + wsb);
    mn_mbar_wait(&S.full[stage], parity);
    // this warp's previous bulk stores must have finished reading its staging buffer
    if (lane == 0) mn_tma_store_wait_read();
    __syncwarp();
    int bc = 0;
    if (lane < nvalid) {
      // Independent values are loaded first and evaluated side by side: a warp's issue rate is set by
      // the dependent fp64 chain of one value (few resident warps), so three / four chains interleave.
      float best = 0.0f;
      int pl = 0;
      for (; pl + 3 <= C; pl += 3) {  // Object ctor, cc:5-21  (0.0f + l == l: l is never -0)
        const float v0 = mn_edge2_clip<CLIP>(in[(size_t)pl * TP]);
        const float v1 = mn_edge2_clip<CLIP>(in[(size_t)(pl + 1) * TP]);
        const float v2 = mn_edge2_clip<CLIP>(in[(size_t)(pl + 2) * TP]);
        const uint32_t i0 = __float_as_uint(v0), i1 = __float_as_uint(v1), i2 = __float_as_uint(v2);
        const float l0 = mn_logf_il(i0, i0 << 29, tabf, lane16);
        const float l1 = mn_logf_il(i1, i1 << 29, tabf, lane16);
        const float l2 = mn_logf_il(i2, i2 << 29, tabf, lane16);
        my_clp[pl] = l0; my_clp[pl + 1] = l1; my_clp[pl + 2] = l2;
        if (pl == 0 || l0 > best) { best = l0; bc = pl; }
        if (l1 > best) { best = l1; bc = pl + 1; }
        if (l2 > best) { best = l2; bc = pl + 2; }
      }
      for (; pl < C; pl++) {
        const uint32_t ix = __float_as_uint(mn_edge2_clip<CLIP>(in[(size_t)pl * TP]));
        const float l = mn_logf_il(ix, ix << 29, tabf, lane16);
        my_clp[pl] = l;
        if (pl == 0 || l > best) { best = l; bc = pl; }
      }
      const float* __restrict__ ina = in + (size_t)C * TP;
      uint32_t amb = 0;  // offsets whose log(1 - s) needs the exact decision
      int k = 0;
      if ((K & 1) == 0) {
        for (; k + 4 <= K; k += 4) {  // AdjacencyRecord ctor, cc:24-36; four offsets per step
          uint32_t ix[4]; double xd[4], y[4]; float sm[4];
#pragma unroll
          for (int u = 0; u < 4; u++) ix[u] = __float_as_uint(mn_edge2_clip<CLIP>(ina[(size_t)(k + u) * TP]));
#pragma unroll
          for (int u = 0; u < 4; u++) xd[u] = mn_f32bits_to_f64(ix[u]);
#pragma unroll
          for (int u = 0; u < 4; u++) sm[u] = mn_logf_il(ix[u], (uint32_t)__double2loint(xd[u]), tabf, lane16);
#pragma unroll
          for (int u = 0; u < 4; u++) y[u] = mn_log1m_il(xd[u], tab1m, lane16);
#pragma unroll
          for (int u = 0; u < 4; u++) amb |= (mn_log1m_ambiguous(y[u]) ? 1u : 0u) << (k + u);
          // 8-byte stores at a stride of K words: conflict free
          *reinterpret_cast<float2*>(my_same + k) = make_float2(sm[0], sm[1]);
          *reinterpret_cast<float2*>(my_same + k + 2) = make_float2(sm[2], sm[3]);
          *reinterpret_cast<float2*>(my_diff + k) = make_float2((float)y[0], (float)y[1]);
          *reinterpret_cast<float2*>(my_diff + k + 2) = make_float2((float)y[2], (float)y[3]);
        }
        for (; k < K; k += 2) {
          const uint32_t ix0 = __float_as_uint(mn_edge2_clip<CLIP>(ina[(size_t)k * TP]));
          const uint32_t ix1 = __float_as_uint(mn_edge2_clip<CLIP>(ina[(size_t)(k + 1) * TP]));
          const double xd0 = mn_f32bits_to_f64(ix0), xd1 = mn_f32bits_to_f64(ix1);
          const float s0 = mn_logf_il(ix0, (uint32_t)__double2loint(xd0), tabf, lane16);
          const float s1 = mn_logf_il(ix1, (uint32_t)__double2loint(xd1), tabf, lane16);
          const double y0 = mn_log1m_il(xd0, tab1m, lane16), y1 = mn_log1m_il(xd1, tab1m, lane16);
          amb |= ((mn_log1m_ambiguous(y0) ? 1u : 0u) | (mn_log1m_ambiguous(y1) ? 2u : 0u)) << k;
          *reinterpret_cast<float2*>(my_same + k) = make_float2(s0, s1);
          *reinterpret_cast<float2*>(my_diff + k) = make_float2((float)y0, (float)y1);
        }
      } else {
        for (; k < K; k++) {
          const uint32_t ix = __float_as_uint(mn_edge2_clip<CLIP>(ina[(size_t)k * TP]));
          const double xd = mn_f32bits_to_f64(ix);
          my_same[k] = mn_logf_il(ix, (uint32_t)__double2loint(xd), tabf, lane16);
          const double y = mn_log1m_il(xd, tab1m, lane16);
          my_diff[k] = (float)y;
          amb |= (mn_log1m_ambiguous(y) ? 1u : 0u) << k;
        }
      }
      while (amb) {  // rare: the full-precision log decides (cc:34)
        const int ka = __ffs((int)amb) - 1;
        amb &= amb - 1;
        my_diff[ka] = mn_log1m_decide(mn_edge2_clip<CLIP>(ina[(size_t)ka * TP]));
      }
    }
    mn_fence_proxy_async();  // staged results -> visible to the bulk-store engine
    __syncwarp();
    if (lane == 0) {
      // hand the input stage back to the producer, then send this warp's slice off
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mn_smem_u32(&S.empty[stage])) : "memory");
      if (nvalid > 0) {
        const size_t p0 = (size_t)start + (size_t)warp * 32;
        mn_tma_store_1d(g_clp + p0 * C, o_clp, (uint32_t)(nvalid * C * 4));
        mn_tma_store_1d(g_same + p0 * K, o_same, (uint32_t)(nvalid * K * 4));
        mn_tma_store_1d(g_diff + p0 * K, o_diff, (uint32_t)(nvalid * K * 4));
        mn_tma_store_commit();
      }
    }
    if (lane < nvalid) g_cls[start + px] = bc;
    b += step_b; t += step_t;
    if (t >= tpi) { t -= tpi; b++; }
    if (++stage == NS) { stage = 0; parity ^= 1; }
  }
  if (lane == 0) mn_tma_store_wait_read();
}

// ---------------------------------------------------------------------------------------------
struct MnRecInitParams {
  MnImage im;  // one image per launch (its sort keys go to a scratch buffer shared by the batch)
  uint64_t* keys_out;
  int H, W, C, K, N;
  int off_r[MN_MAX_K], off_c[MN_MAX_K];
  int rank_of_k[MN_MAX_K];  // rank of |linear delta| among the offsets (tie-break ordinal)
  float omf, mlb;
};

__global__ void __launch_bounds__(256) mn_record_init_kernel(MnRecInitParams P) {
  const long long E = (long long)P.N * P.K;
  const MnImage& im = P.im;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < E;
       g += (long long)gridDim.x * blockDim.x) {
    const int r = (int)g;
    const int p = r / P.K, k = r - p * P.K;
    const int row = p / P.W, col = p - row * P.W;
    if (k == 0) {
      // live-record masks: bit k = record (p,k) exists; bit 16+k = record (p - o_k, k) exists
      uint32_t m = 0;
      for (int kk = 0; kk < P.K; kk++) {
        int r2 = row + P.off_r[kk], c2 = col + P.off_c[kk];
        if (r2 >= 0 && r2 < P.H && c2 >= 0 && c2 < P.W) m |= 1u << kk;
        r2 = row - P.off_r[kk];
        c2 = col - P.off_c[kk];
        if (r2 >= 0 && r2 < P.H && c2 >= 0 && c2 < P.W) m |= 1u << (16 + kk);
      }
      im.parent[p] = p;
      im.obj[p] = make_uint4(mn_pack_nc(1, im.cls[p]), 0u /* sameness sum 0.0f */, 0xffffffffu /* no pixel array */, m);
    }
    const int r2 = row + P.off_r[k], c2 = col + P.off_c[k];
    uint64_t key = ~0ull;
    if (r2 >= 0 && r2 < P.H && c2 >= 0 && c2 < P.W) {
      const int q = r2 * P.W + c2;
      const int lo = p < q ? p : q, hi = p < q ? q : p;  // cc:49-56
      const float oml = MN_FSUB(im.rec_same[r], im.rec_diff[r]);  // cc:36
      const int cl = im.cls[lo], ch = im.cls[hi];
      const float mp = mn_priority(oml, P.omf, P.mlb, P.C, 1, cl, im.clp + (size_t)lo * P.C, 1, ch,
                                   im.clp + (size_t)hi * P.C, nullptr);  // cc:45
      const MnHashPos hp = mn_hash_pos(im.hash_nbuckets, lo, hi);
      const uint32_t g = mp >= 0.0f ? MN_G_EXACT : MN_G_NONE;  // the initial entry is queued iff mp >= 0 (cc:225-227)
      mn_store_rec(im, r, make_uint4(mn_rec_pack_x(lo, MN_HS_NONE, g), (uint32_t)hi, mn_f2u(oml), mn_f2u(mp)));  // (the hash verifies keys through the record)
      const int hslot = mn_hash_insert(im, lo, hi, r);
      MN_REC(im, r).x = mn_rec_pack_x(lo, mn_hs_of_slot(hp, hslot), g);
      if (mp >= 0.0f) {  // cc:225-227
        uint32_t ord = (mn_tie_u(lo, hi) << 4) | (uint32_t)P.rank_of_k[k];
        key = ((uint64_t)(~mn_f2u(mp == 0.0f ? 0.0f : mp)) << MN_ORD_BITS) | ord;
      }
    } else {
      mn_store_rec(im, r, make_uint4(MN_REC_DEAD, 0u, 0u, mn_f2u(-1.0f)));
    }
    P.keys_out[r] = key;
  }
}
