// mn_api.cu -- host driver and C ABI of libmergenet_b200.so (see include/mergenet_b200.h).
//
// Pipeline for a batch of B images of one shape, all on one CUDA stream:
//   1. mn_edge_warp_kernel (fast path) / mn_edge_pass_kernel (general): whole GPU, HBM-bound, TMA-staged (mn_edge.cuh)
//   2. per image: mn_record_init_kernel -> radix sort of the initial queue keys (cub, library call)
//   3. mn_merge_kernel          persistent, one CTA per image, order-exact scheduler (mn_merge.cuh)
//   4. labels: root flags -> exclusive scan (cub) -> mask / object_class     (this file)
//   5. mn_partition_logprob_kernel: total log-prob of the final partition, HBM-bound streaming pass (this file)
// There is no CPU implementation behind this API: without a CUDA device every call fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cub/cub.cuh>
#include <mutex>
#include <vector>

#include "../../include/mergenet_b200.h"
#include "mn_common.h"
#include "mn_edge.cuh"
#include "mn_post.cuh"
#include "mn_layout.h"
#include "mn_merge.cuh"
#include "mn_modeb.cuh"
#include "mn_exact.cuh"
#include "mn_stl_primes.h"

// ------------------------------------------------------------------------------------------------
static thread_local int g_last_error = MN_STATUS_OK;

#define MN_CUDA_OK(expr)                                                                \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      fprintf(stderr, "mergenet_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e__), \
              __FILE__, __LINE__);                                                      \
      g_last_error = MN_STATUS_CUDA;                                                    \
      return MN_STATUS_CUDA;                                                            \
    }                                                                                   \
  } while (0)

extern "C" int mn_last_error(void) { return g_last_error; }
extern "C" const char* mn_status_string(int s) {
  switch (s) {
    case MN_STATUS_OK: return "ok";
    case MN_STATUS_BAD_ARG: return "bad argument";
    case MN_STATUS_PL_POOL: return "pixel-list chunk pool exhausted";
    case MN_STATUS_Q_POOL: return "queue entry pool exhausted";
    case MN_STATUS_TREE_POOL: return "queue tree node pool exhausted";
    case MN_STATUS_HASH_FULL: return "record hash table full";
    case MN_STATUS_INTERNAL: return "internal invariant failed";
    case MN_STATUS_CUDA: return "CUDA error / no usable device";
    case MN_STATUS_LIMIT: return "iteration guard tripped";
    case MN_STATUS_NO_BACKGROUND: return "Mode B prune: no class-0 object to fold into (the reference raises UnboundLocalError)";
  }
  return "unknown";
}
extern "C" int mn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

// ------------------------------------------------------------------------------------------------
// kernels of this file
#ifndef MN_MERGE_THREADS
#define MN_MERGE_THREADS 512
#endif
__global__ void __launch_bounds__(MN_MERGE_THREADS, 1) mn_merge_kernel(const MnImage* imgs, int nimg, MnMergeArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  MnSm& sm = *reinterpret_cast<MnSm*>(smem_raw);
  float* c_clp = reinterpret_cast<float*>(smem_raw + ((sizeof(MnSm) + 15) / 16) * 16);
  for (int b = blockIdx.x; b < nimg; b += gridDim.x) {
    const MnImage im = imgs[b];
    // every array of the workspace is global memory: lets the compiler emit LDG / STG instead of
    // generic loads and stores for the scheduler's scattered accesses
    __builtin_assume(__isGlobal(im.clp)); __builtin_assume(__isGlobal(im.cls)); __builtin_assume(__isGlobal(im.obj));
    __builtin_assume(__isGlobal(im.parent)); __builtin_assume(__isGlobal(im.pix_pool)); __builtin_assume(__isGlobal(im.rec));
    __builtin_assume(__isGlobal(im.hash)); __builtin_assume(__isGlobal(im.hash_ovf)); __builtin_assume(__isGlobal(im.init_keys));
    __builtin_assume(__isGlobal(im.q_ent)); __builtin_assume(__isGlobal(im.qc_next)); __builtin_assume(__isGlobal(im.qc_free));
    __builtin_assume(__isGlobal(im.tn)); __builtin_assume(__isGlobal(im.tn_dir)); __builtin_assume(__isGlobal(im.ctl));
    long long t0 = clock64();
    mn_merge_image(im, sm, A, c_clp);
    if (threadIdx.x == 0) im.ctl->cycles_total = clock64() - t0;
    __syncthreads();
  }
}

// control blocks, hash tables, tree nodes: reset before every run
__global__ void mn_reset_kernel(const MnImage* imgs, int nimg) {
  for (int b = blockIdx.y; b < nimg; b += gridDim.y) {
    const MnImage im = imgs[b];
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    const long long nh = (long long)im.hash_nbuckets * 8;
    for (long long i = tid; i < nh; i += nth) im.hash[i] = 0;
    for (long long i = tid; i < im.tn_cap; i += nth) im.tn[i] = make_int4(-1, -1, 0, -1);
    if (tid == 0) {
      MnCtl z;
      memset(&z, 0, sizeof(z));
      z.tn_bump = MN_NROOTS;
      *im.ctl = z;
    }
  }
}

// root flags: 1 for surviving objects with class != 0 (cc:503-508); written over the dead cls array
__global__ void mn_label_flags_kernel(const MnImage* imgs, int nimg, int N) {
  for (int b = blockIdx.y; b < nimg; b += gridDim.y) {
    const MnImage im = imgs[b];
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x)
      im.cls[p] = (im.parent[p] == p && mn_nc_cls(im.obj[p].x) != 0) ? 1 : 0;
  }
}
// mask / object_class (cc:491-517); labels ascend with the surviving object's id
__global__ void mn_label_write_kernel(const MnImage* imgs, int nimg, int N, int* d_mask,
                                      int* d_object_class, int* d_ninst) {
  for (int b = blockIdx.y; b < nimg; b += gridDim.y) {
    const MnImage im = imgs[b];
    const int* flags = im.cls;
    const int* excl = im.pix_pool;  // exclusive scan of flags (the pixel arrays are dead by now)
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
      int r = p;
      for (int g = 0; g < (1 << 26); g++) {
        int q = im.parent[r];
        if (q == r) break;
        r = q;
      }
      int lab = flags[r] ? excl[r] + 1 : 0;
      d_mask[(size_t)b * N + p] = lab;
      if (r == p && flags[p]) d_object_class[(size_t)b * N + excl[p]] = mn_nc_cls(im.obj[p].x);
      if (p == N - 1) {
        int n = excl[N - 1] + flags[N - 1];
        d_ninst[b] = n;
        im.ctl->n_instances = n;
      }
    }
  }
}

// Partition statistics pass (cc:314-350, ComputeTotalLogprobFromScratch; SURVEY 8 a12 names this definition the
// parity number): the total log-probability of the FINAL partition, evaluated from the maps and the label mask in
// float64 -- class term over every pixel (the class of its instance, 0 for background), sameness term over the
// in-image (pixel, offset) pairs that ended inside one instance, differentness term over those that ended across
// two.  A pure HBM-bound streaming pass: K sameness planes + the label mask + one class value per pixel read once
// (neighbour labels come from L1/L2), per-thread fp64 partial sums folded by warp shuffles, one partial triple per
// block, folded in a fixed order by mn_partition_logprob_fold_kernel (deterministic, no global atomics).
struct MnLogprobParams {
  const float* d_class; const float* d_adj; const int* d_mask; const int* d_object_class;
  double* partial;  // [nimg][gridDim.x][3]
  int nimg, H, W, C, K, N;
  int maxr, maxc;   // largest |row offset| / |column offset|: pixels farther than that from every border take no bounds test
  int off_r[MN_MAX_K], off_c[MN_MAX_K], delta[MN_MAX_K];
};
// MODE: the MN_INPUT_* transform the run applied while reading the maps (0: values as given)
template <int MODE>
__device__ __forceinline__ float mn_logprob_input(float v) {
  if (MODE & MN_INPUT_LOGITS) v = mn_sigmoid_f32(v);
  if (MODE & (MN_INPUT_LOGITS | MN_INPUT_CLIP)) v = fminf(fmaxf(v, 1.1920929e-07f), 0.99999988f);
  return v;
}
// log(a) + log(b) = log(a * b): a thread keeps its sums as PRODUCTS in float64 -- mantissa in the double, the exponent
// peeled off into an integer every few pixels (a factor is >= 2^-24, so 34 of them cannot leave the double range) --
// and takes one log per sum at the end: the pass is bound by its loads and instruction issue, not by the fp64 pipe.
// Equals the sum of the logs to float64 rounding (the widening of s and 1.0 - (double)s are exact).  Factors outside
// [2^-24, 1) -- an unclipped caller -- take libm's log directly, with its special values.
struct MnProdAcc {
  double m; int e;
  __device__ __forceinline__ void init() { m = 1.0; e = 0; }
  __device__ __forceinline__ void renorm() {  // mantissa back into [1, 2)
    const int hi = __double2hiint(m);
    e += (hi >> 20) - 1023;
    m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(m));
  }
  __device__ __forceinline__ double total() const { return log(m) + (double)e * 0x1.62e42fefa39efp-1; }
};
#ifndef MN_LP_BLOCKS
#define MN_LP_BLOCKS 4  // resident blocks per SM the register budget is set for (measured: 0.50 ms vs 0.57 ms at 3, 16 images; fetching label and class one pixel ahead and dropping the division: 0.64 ms, not kept)
#endif
#define MN_LP_FAST(bits) ((bits) - 0x33800000u < 0x3f800000u - 0x33800000u)  // a float in [2^-24, 1)
// KT: the number of offsets when it is one of the usual ones (fully unrolled, no predicates), 0 = any K <= MN_MAX_K
template <int MODE, int KT>
__global__ void __launch_bounds__(256, MN_LP_BLOCKS) mn_partition_logprob_kernel(MnLogprobParams P) {
  __shared__ double red[3][8];
  const int N = P.N, W = P.W, H = P.H, K = KT ? KT : P.K;
  constexpr int KU = KT ? KT : MN_MAX_K;
  for (int b = blockIdx.y; b < P.nimg; b += gridDim.y) {
    const int* mask = P.d_mask + (size_t)b * N;
    const int* ocls = P.d_object_class + (size_t)b * N;
    const float* cp = P.d_class + (size_t)b * P.C * N;
    const float* ap = P.d_adj + (size_t)b * K * N;
    MnProdAcc ac, as, ad;  // class factors; s of the pairs inside an instance; 1 - s of the pairs across two
    ac.init(); as.init(); ad.init();
    double slow_c = 0.0, slow_s = 0.0, slow_d = 0.0;  // logs taken one by one (out-of-domain factors)
    int it = 0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x, it++) {
      const int row = p / W, col = p - row * W;
      const int lab = mask[p];
      int cls = lab > 0 ? ocls[lab - 1] : 0;
      cls = cls < 0 ? 0 : (cls >= P.C ? P.C - 1 : cls);  // (only a failed image can hold anything else)
      const float cv = mn_logprob_input<MODE>(cp[(size_t)cls * N + p]);
      float sv[KU]; int lq[KU];
      const bool interior = row >= P.maxr && row < H - P.maxr && col >= P.maxc && col < W - P.maxc;
      uint32_t missing = 0;  // bit k: the pair (p, k) leaves the image
      if (interior) {   // every load of the pixel issued before the first use
#pragma unroll
        for (int k = 0; k < KU; k++)
          if (KT || k < K) { sv[k] = ap[(size_t)k * N + p]; lq[k] = mask[p + P.delta[k]]; }
      } else {          // border band: pairs that leave the image do not exist
#pragma unroll
        for (int k = 0; k < KU; k++) {
          if (KT || k < K) {
            const int r2 = row + P.off_r[k], c2 = col + P.off_c[k];
            const bool in = r2 >= 0 && r2 < H && c2 >= 0 && c2 < W;
            sv[k] = in ? ap[(size_t)k * N + p] : 0.5f;
            lq[k] = in ? mask[r2 * W + c2] : lab;
            missing |= in ? 0u : (1u << k);
          }
        }
      }
      bool fast = true;
#pragma unroll
      for (int k = 0; k < KU; k++)
        if (KT || k < K) {
          sv[k] = mn_logprob_input<MODE>(sv[k]);
          if (!interior && ((missing >> k) & 1u)) sv[k] = 1.0f;  // no pair: the factor 1 on the "inside" product (lq = lab)
          else if (MODE == 0) fast = fast && MN_LP_FAST(__float_as_uint(sv[k]));  // (clipped values are in the domain)
        }
      if (fast) {
#pragma unroll
        for (int k = 0; k < KU; k++) {
          if (KT || k < K) {
            const double sd = mn_f32bits_to_f64(__float_as_uint(sv[k]));
            if (lq[k] == lab) as.m = __dmul_rn(as.m, sd);
            else ad.m = __dmul_rn(ad.m, __dadd_rn(1.0, -sd));
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < KU; k++) {  // (static indices: sv / lq stay in registers)
          if ((KT || k < K) && !((missing >> k) & 1u)) {
            const bool same = lq[k] == lab;
            const double l = log(same ? (double)sv[k] : 1.0 - (double)sv[k]);
            slow_s += same ? l : 0.0; slow_d += same ? 0.0 : l;
          }
        }
      }
      {  // the class factor last: its load hangs on two earlier ones (label -> class of the label -> that plane)
        const uint32_t vb = __float_as_uint(cv);
        if (vb - 0x00800000u < 0x7f000000u) ac.m = __dmul_rn(ac.m, mn_f32bits_to_f64(vb));  // positive normal float
        else slow_c += log((double)cv);
      }
      if (it & 1) { as.renorm(); ad.renorm(); }   // (at most 2 * 16 factors >= 2^-24 since the last one)
      if ((it & 7) == 7) ac.renorm();              // (8 class factors >= 2^-126)
    }
    double tc = ac.total() + slow_c, ts = as.total() + slow_s, td = ad.total() + slow_d;
    for (int o = 16; o > 0; o >>= 1) {
      tc += __shfl_xor_sync(0xffffffffu, tc, o); ts += __shfl_xor_sync(0xffffffffu, ts, o); td += __shfl_xor_sync(0xffffffffu, td, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = tc; red[1][threadIdx.x >> 5] = ts; red[2][threadIdx.x >> 5] = td; }
    __syncthreads();
    if (threadIdx.x < 3) {
      double t = 0.0;
      for (int w = 0; w < 8; w++) t += red[threadIdx.x][w];
      P.partial[((size_t)b * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = t;
    }
    __syncthreads();
  }
}
template <int MODE>
static void mn_launch_partition_logprob(const MnLogprobParams& L, dim3 g, cudaStream_t s) {
  if (L.K == 10) mn_partition_logprob_kernel<MODE, 10><<<g, 256, 0, s>>>(L);
  else if (L.K == 16) mn_partition_logprob_kernel<MODE, 16><<<g, 256, 0, s>>>(L);
  else mn_partition_logprob_kernel<MODE, 0><<<g, 256, 0, s>>>(L);
}
// one warp per image and term: the block partials in index order (fixed summation tree)
__global__ void mn_partition_logprob_fold_kernel(const double* partial, int nimg, int nblk, double* out /* [nimg][4] */) {
  const int b = blockIdx.x, term = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (b >= nimg || term >= 3) return;
  double t = 0.0;
  for (int i = lane; i < nblk; i += 32) t += partial[((size_t)b * nblk + i) * 3 + term];
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) out[(size_t)b * 4 + term] = t;
}

__global__ void mn_libm_kernel(int which, uint32_t first_bits, uint32_t n, float bias, float* out) {
  __shared__ MnLogfTab tab[16];
  __shared__ MnLog1mTab tab1m[128];
  __shared__ double2 tabf_il[16 * 8], tab1m_il[64 * 8];  // the interleaved tables of mn_edge_warp_kernel
  for (int e = threadIdx.x; e < 16 * 8; e += blockDim.x) tabf_il[e] = make_double2(mn_logf_table_c[e >> 3].invc, mn_logf_table_c[e >> 3].logc);
  for (int e = threadIdx.x; e < 64 * 8; e += blockDim.x) tab1m_il[e] = make_double2(mn_log1m64_table[e >> 3].invc, mn_log1m64_table[e >> 3].logc);
  const uint32_t lane16 = (threadIdx.x & 7) * 16u;
  if (threadIdx.x < 16) {
    const MnLogfTab t16[16] = {MN_LOGF_TABLE};
    tab[threadIdx.x] = t16[threadIdx.x];
  }
  if (threadIdx.x < 128) tab1m[threadIdx.x] = mn_log1m_table[threadIdx.x];
  __syncthreads();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float x = __uint_as_float(first_bits + i);
    float r;
    if (which == 0) r = mn_logf_fast(x, tab);            // what the edge pass evaluates
    else if (which == 1) r = mn_log1m_fast(x, tab1m);
    else if (which == 3) r = mn_logf_exact(x, tab);      // the unfused recipe (bias path)
    else if (which == 4) r = mn_log1m_exact(x);
    else if (which == 5) r = mn_logf_il(__float_as_uint(x), __float_as_uint(x) << 29, tabf_il, lane16);  // warp-pipeline kernel
    else if (which == 6) {
      const double y = mn_log1m_il(mn_f32bits_to_f64(__float_as_uint(x)), tab1m_il, lane16);
      r = mn_log1m_ambiguous(y) ? mn_log1m_decide(x) : (float)y;
    }
    else r = mn_bias_sameness(x, bias, tab);
    out[i] = r;
  }
}

// ------------------------------------------------------------------------------------------------
#define MN_COPY_EVENTS 64
#define MN_LOGPROB_BLOCKS 512  // blocks per image of mn_partition_logprob_kernel (upper bound)
#define MN_INPUT_CHECK_DOMAIN 16  // internal input flag of the drop-in entry: maps may lie outside [2^-126, 1)
struct mn_plan {
  int device, max_batch, H, W, C, K, N;
  long long E;
  int offsets[2 * MN_MAX_K];
  int rank_of_k[MN_MAX_K];
  MnOffsets off;
  size_t per_image_bytes;
  unsigned char* d_ws;        // max_batch * per_image_bytes
  MnImage* d_imgs;            // device array of per-image pointer sets
  std::vector<MnImage> h_imgs;
  uint64_t* d_keys_scratch;   // E keys (record init writes, the sort reads)
  void* d_cub_temp;
  size_t cub_temp_bytes;
  // staging for the host-buffer entry point
  float* d_in_class; float* d_in_adj; int* d_out_mask; int* d_out_cls; int* d_out_ninst;
  size_t staging_batch;
  cudaStream_t stream;
  cudaStream_t copy_stream;  // uploads of the host-buffer entry point
  cudaEvent_t copy_ev[MN_COPY_EVENTS];
  cudaEvent_t ev[9];
  int num_sms;
  int edge_tp, edge_smem, merge_smem, merge_H;
  int edge2_ncons, edge2_ctas, edge2_smem, edge2_stages;  // warp-pipeline edge kernel: consumer warps per CTA (0: not usable), CTAs per SM
  std::vector<MnCtl> h_ctl;
  int* d_domain_flag;           // raised by mn_domain_check_kernel (drop-in entry: unclipped maps)
  double* d_logprob;            // [max_batch][4] class / sameness / differentness terms
  double* d_logprob_partial;    // [max_batch][MN_LOGPROB_BLOCKS][3] block partials of the partition statistics pass
  std::vector<double> h_logprob;
  float last_omf;
  mn_timings timings;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct WsLayout {
  size_t clp, cls, obj, parent, pix_pool, rec, arena, hash, hash_ovf,
      qc_next, qc_free, tn, tn_dir, ctl, total;
  int pix_cap, qc_cap, qc_low_n, tn_cap;
  uint32_t hash_nbuckets, hash_ovf_cap;
};
static WsLayout ws_layout(int H, int W, int C, int K) {
  WsLayout L;
  const size_t N = (size_t)H * W, E = N * K;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
  const MnCaps caps = mn_workspace_caps(N, E);  // (mn_layout.h: shared with the scheduler's host build)
  L.pix_cap = caps.pix_cap; L.qc_low_n = caps.qc_low_n; L.qc_cap = caps.qc_cap; L.tn_cap = caps.tn_cap;
  L.hash_nbuckets = caps.hash_nbuckets; L.hash_ovf_cap = caps.hash_ovf_cap;
  L.clp = take(N * C * 4);
  L.cls = take(N * 4);
  L.obj = take(N * 16);
  L.parent = take(N * 4);
  L.pix_pool = take((size_t)L.pix_cap * 4);
  L.rec = take(E * 16);
  L.arena = take(std::max(E * 8, (size_t)L.qc_cap * MN_QCH * 16));
  L.hash = take((size_t)L.hash_nbuckets * 8 * 4);
  L.hash_ovf = take((size_t)L.hash_ovf_cap * 4);
  L.qc_next = take((size_t)L.qc_cap * 4);
  L.qc_free = take((size_t)L.qc_cap * 4);
  L.tn = take((size_t)L.tn_cap * 16);
  L.tn_dir = take((size_t)L.tn_cap * 8 * 4);
  L.ctl = take(sizeof(MnCtl));
  L.total = o;
  return L;
}

extern "C" size_t mn_workspace_bytes_per_image(int H, int W, int C, int K) {
  if (H <= 0 || W <= 0 || C <= 0 || K <= 0) return 0;
  return ws_layout(H, W, C, K).total;
}

static int choose_edge_tile(int C, int K, int* smem_bytes) {
  // per pixel: double-buffered inputs 2*(C+K) floats + double-buffered staged outputs 2*(C + 2K) floats
  const size_t per_px = 4 * (size_t)(2 * (C + K) + 2 * (C + 2 * K));
  size_t budget = (size_t)(226 * 1024) / MN_EDGE_CTAS_PER_SM - 4096;
  int tp = (int)(budget / per_px);
  tp = std::min(tp, MN_EDGE_THREADS / 2);
  tp = tp >= 32 ? tp / 32 * 32 : tp / 4 * 4;  // (multiples of 32 pixels: the 2-D tensor-map boxes land 128-byte aligned)
  if (tp >= 128) tp = tp / 128 * 128;
  if (tp < 4) tp = 4;
  *smem_bytes = (int)(128 + per_px * tp + 16 * sizeof(MnLogfTab) + 128 * sizeof(MnLog1mTab) + 64);
  return tp;
}

// warp-pipeline edge kernel: consumer warps per CTA and CTAs per SM that maximise the resident consumer
// warps; 0 when the shape leaves too few (many classes: the tile kernel's smaller tiles win)
static int choose_edge2(int C, int K, int* ctas_out, int* smem_bytes, int* stages_out) {
  int stages = 2;
  if (const char* e = getenv("MN_EDGE2_STAGES")) { int v = atoi(e); if (v >= 2 && v <= MN_EDGE2_MAX_STAGES) stages = v; }
  *stages_out = stages;
  const size_t per_warp = 128 * ((size_t)stages * (C + K) + (size_t)(C + 2 * K));
  const size_t fixed = sizeof(MnEdge2Smem) + 1024 + 256;  // static tables + per-CTA reservation
  int best = 0, best_score = 0, best_ctas = 0;
  for (int nc = MN_EDGE2_MAX_THREADS / 32 - 1; nc >= 1; nc--) {
    size_t per_cta = per_warp * nc + fixed;
    int ctas = (int)std::min<size_t>((size_t)(227 * 1024) / per_cta, 2048 / (32 * (nc + 1)));
    ctas = std::min(ctas, 65536 / (80 * 32 * (nc + 1)));  // register file at <= 80 registers per thread
    if (ctas * nc > best_score) { best_score = ctas * nc; best = nc; best_ctas = ctas; }
  }
  if (const char* e = getenv("MN_EDGE2_NCONS")) {  // tuning hook: "ncons,ctas"; "0" disables the kernel
    int nc = 0, ct = 0;
    int got = sscanf(e, "%d,%d", &nc, &ct);
    if (got >= 1 && nc == 0) return 0;
    if (got == 2 && nc >= 1 && nc <= MN_EDGE2_MAX_THREADS / 32 - 1 && ct >= 1) { best = nc; best_ctas = ct; best_score = 99; }
  }
  if (best_score < 12) return 0;
  *ctas_out = best_ctas;
  *smem_bytes = (int)(per_warp * best);
  return best;
}

extern "C" void mn_plan_destroy(mn_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  if (p->copy_stream) {
    cudaStreamSynchronize(p->copy_stream);
    for (int i = 0; i < MN_COPY_EVENTS; i++) cudaEventDestroy(p->copy_ev[i]);
    cudaStreamDestroy(p->copy_stream);
  }
  cudaFree(p->d_ws); cudaFree(p->d_imgs); cudaFree(p->d_keys_scratch); cudaFree(p->d_cub_temp); cudaFree(p->d_logprob); cudaFree(p->d_logprob_partial); cudaFree(p->d_domain_flag);
  cudaFree(p->d_in_class); cudaFree(p->d_in_adj); cudaFree(p->d_out_mask); cudaFree(p->d_out_cls);
  cudaFree(p->d_out_ninst);
  for (int i = 0; i < 9; i++) if (p->ev[i]) cudaEventDestroy(p->ev[i]);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
}

extern "C" int mn_plan_create(mn_plan** out, int max_batch, int H, int W, int C, int K,
                              const int* offset_list, int device) {
  g_last_error = MN_STATUS_OK;
  if (!out || max_batch <= 0 || H <= 0 || W <= 0 || C <= 0 || C >= MN_MAX_C || K <= 0 || K > MN_MAX_K ||
      !offset_list || (long long)H * W * K > (1ll << 25) || (long long)H * W >= (1 << 24)) {
    g_last_error = MN_STATUS_BAD_ARG;
    return MN_STATUS_BAD_ARG;
  }
  // the reference's config contract (core_config.py:66-73): no (0,0), no duplicates, no negated pairs
  for (int a = 0; a < K; a++) {
    int ar = offset_list[2 * a], ac = offset_list[2 * a + 1];
    bool bad = (ar == 0 && ac == 0);
    for (int b = 0; b < a && !bad; b++) {
      int br = offset_list[2 * b], bc = offset_list[2 * b + 1];
      bad = (ar == br && ac == bc) || (ar == -br && ac == -bc);
    }
    if (bad) { g_last_error = MN_STATUS_BAD_ARG; return MN_STATUS_BAD_ARG; }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    g_last_error = MN_STATUS_CUDA;
    return MN_STATUS_CUDA;
  }
  MN_CUDA_OK(cudaSetDevice(device));
  mn_plan* p = new mn_plan();
  memset(&p->timings, 0, sizeof(p->timings));
  p->device = device; p->max_batch = max_batch; p->H = H; p->W = W; p->C = C; p->K = K; p->N = H * W;
  p->E = (long long)p->N * K;
  p->d_ws = nullptr; p->d_imgs = nullptr; p->d_keys_scratch = nullptr; p->d_cub_temp = nullptr; p->d_logprob = nullptr; p->d_logprob_partial = nullptr; p->d_domain_flag = nullptr; p->last_omf = 1.0f;
  p->d_in_class = nullptr; p->d_in_adj = nullptr; p->d_out_mask = nullptr; p->d_out_cls = nullptr; p->d_out_ninst = nullptr;
  p->staging_batch = 0; p->stream = nullptr;
  for (int i = 0; i < 9; i++) p->ev[i] = nullptr;
  memcpy(p->offsets, offset_list, sizeof(int) * 2 * K);
  std::vector<std::pair<int, int>> mag;
  p->off.K = K;
  for (int k = 0; k < K; k++) {
    p->off.delta[k] = offset_list[2 * k] * W + offset_list[2 * k + 1];
    mag.push_back(std::make_pair(abs(p->off.delta[k]), k));
  }
  std::sort(mag.begin(), mag.end());
  for (int r = 0; r < K; r++) { p->off.k_of_rank[r] = mag[r].second; p->rank_of_k[mag[r].second] = r; }

  cudaDeviceProp prop;
  MN_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  p->num_sms = prop.multiProcessorCount;
  WsLayout L = ws_layout(H, W, C, K);
  p->per_image_bytes = L.total;
  auto fail = [&](int code) { mn_plan_destroy(p); g_last_error = code; return code; };
  if (cudaMalloc(&p->d_ws, L.total * (size_t)max_batch) != cudaSuccess) return fail(MN_STATUS_CUDA);
  if (cudaMalloc(&p->d_imgs, sizeof(MnImage) * max_batch) != cudaSuccess) return fail(MN_STATUS_CUDA);
  if (cudaMalloc(&p->d_keys_scratch, (size_t)p->E * 8) != cudaSuccess) return fail(MN_STATUS_CUDA);
  if (cudaMalloc(&p->d_logprob, sizeof(double) * 4 * max_batch) != cudaSuccess) return fail(MN_STATUS_CUDA);
  if (cudaMalloc(&p->d_logprob_partial, sizeof(double) * 3 * MN_LOGPROB_BLOCKS * max_batch) != cudaSuccess) return fail(MN_STATUS_CUDA);
  if (cudaMalloc(&p->d_domain_flag, sizeof(int)) != cudaSuccess) return fail(MN_STATUS_CUDA);
  p->h_logprob.assign((size_t)4 * max_batch, 0.0);
  p->h_imgs.resize(max_batch);
  for (int b = 0; b < max_batch; b++) {
    unsigned char* base = p->d_ws + (size_t)b * L.total;
    MnImage im;
    memset(&im, 0, sizeof(im));
    im.clp = (float*)(base + L.clp); im.cls = (int*)(base + L.cls);
    im.obj = (uint4*)(base + L.obj);
    im.parent = (int*)(base + L.parent);
    im.pix_pool = (int*)(base + L.pix_pool);
    im.rec = (uint4*)(base + L.rec);
    im.rec_same = (float*)(base + L.arena); im.rec_diff = im.rec_same + p->E;
    im.hash = (uint32_t*)(base + L.hash); im.hash_ovf = (uint32_t*)(base + L.hash_ovf);
    im.hash_nbuckets = L.hash_nbuckets; im.hash_ovf_cap = L.hash_ovf_cap;
    im.init_keys = (uint64_t*)(base + L.arena);
    im.q_ent = (uint4*)(base + L.arena);
    im.qc_next = (int*)(base + L.qc_next); im.qc_free = (int*)(base + L.qc_free);
    im.tn = (int4*)(base + L.tn); im.tn_dir = (int*)(base + L.tn_dir);
    im.pix_cap = L.pix_cap; im.qc_cap = L.qc_cap; im.qc_low_n = L.qc_low_n; im.tn_cap = L.tn_cap;
    im.out_mask = nullptr; im.out_cls = nullptr;
    im.ctl = (MnCtl*)(base + L.ctl);
    p->h_imgs[b] = im;
  }
  if (cudaMemcpy(p->d_imgs, p->h_imgs.data(), sizeof(MnImage) * max_batch, cudaMemcpyHostToDevice) != cudaSuccess)
    return fail(MN_STATUS_CUDA);
  // cub temp storage: the larger of the key sort and the label scan
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, t1, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)p->E, 0, 32 + MN_ORD_BITS);
  cub::DeviceScan::ExclusiveSum(nullptr, t2, (const int*)nullptr, (int*)nullptr, p->N);
  p->cub_temp_bytes = std::max(t1, t2);
  if (cudaMalloc(&p->d_cub_temp, p->cub_temp_bytes) != cudaSuccess) return fail(MN_STATUS_CUDA);
  if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(MN_STATUS_CUDA);
  for (int i = 0; i < 9; i++)
    if (cudaEventCreate(&p->ev[i]) != cudaSuccess) return fail(MN_STATUS_CUDA);
  p->edge_tp = choose_edge_tile(C, K, &p->edge_smem);
  {
    // window: as many candidates as the staged class vectors (3 per candidate) leave room for
    const size_t base = ((sizeof(MnSm) + 15) / 16) * 16;
    const size_t limit = (size_t)prop.sharedMemPerBlockOptin;
    int Hwin = MN_H;
    while (Hwin > 1 && base + (size_t)Hwin * 3 * C * 4 > limit) Hwin--;
    p->merge_H = Hwin;
    p->merge_smem = (int)(base + (size_t)Hwin * 3 * C * 4);
    if ((size_t)p->merge_smem > limit) return fail(MN_STATUS_BAD_ARG);
  }
  p->edge2_ncons = choose_edge2(C, K, &p->edge2_ctas, &p->edge2_smem, &p->edge2_stages);
  if (p->edge2_ncons &&
      (cudaFuncSetAttribute(mn_edge_warp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->edge2_smem) != cudaSuccess ||
       cudaFuncSetAttribute(mn_edge_warp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->edge2_smem) != cudaSuccess ||
       cudaFuncSetAttribute(mn_edge_warp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->edge2_smem) != cudaSuccess))
    p->edge2_ncons = 0;
  if (cudaFuncSetAttribute(mn_edge_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p->edge_smem) != cudaSuccess ||
      cudaFuncSetAttribute(mn_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p->merge_smem) != cudaSuccess)
    return fail(MN_STATUS_CUDA);
  p->h_ctl.resize(max_batch);
  *out = p;
  return MN_STATUS_OK;
}

// 2-D tensor map over a [planes][N] fp32 array, box = `box_rows` planes x `tp` pixels (dense rows in shared memory);
// the encoder comes from the driver at run time (cudaGetDriverEntryPoint: the library does not link libcuda)
typedef CUresult (*mn_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static mn_encode_tiled_fn mn_tensor_map_encoder() {
  static mn_encode_tiled_fn fn = []() -> mn_encode_tiled_fn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return (mn_encode_tiled_fn)f;
  }();
  return fn;
}
static bool make_plane_map(CUtensorMap* m, const float* base, size_t planes, size_t n, int box_rows, int tp) {
  mn_encode_tiled_fn enc = mn_tensor_map_encoder();
  if (!enc || box_rows > 256 || tp > 256 || planes >= (1ull << 32)) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)planes};
  const cuuint64_t strides[1] = {(cuuint64_t)n * 4};
  const cuuint32_t box[2] = {(cuuint32_t)tp, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// the edge pass for images [0, B): the warp-pipeline kernel when it applies, else the tile kernel
static int launch_edge(mn_plan* p, int b0, int B, const float* d_class, float* d_adj, int clip, float sdb, cudaStream_t s) {
  const int N = p->N, C = p->C, K = p->K;
  MnEdgeParams P;
  P.class_pred = d_class; P.adj_pred = d_adj; P.adj_pred_rw = sdb != 0.0f ? d_adj : nullptr;
  P.imgs = p->d_imgs + b0;
  P.B = B; P.C = C; P.K = K; P.N = N; P.TP = p->edge_tp;
  P.tiles_per_image = (N + P.TP - 1) / P.TP;
  P.use_tma = (N % 4 == 0) && (((uintptr_t)d_class & 15) == 0) && (((uintptr_t)d_adj & 15) == 0);
  P.logits = (clip & MN_INPUT_LOGITS) ? 1 : 0;
  P.clip = ((clip & MN_INPUT_CLIP) || P.logits) ? 1 : 0;  // (sigmoid saturates to 0 / 1 in fp32: logits are always clipped)
  P.sdb = sdb;
  P.only_if = nullptr;
  P.use_tmap = 0;
  CUtensorMap tmc, tma;
  memset(&tmc, 0, sizeof(tmc)); memset(&tma, 0, sizeof(tma));
  const bool no_tmap = getenv("MN_EDGE_NO_TMAP") != nullptr;  // (A/B hook: the 1-D bulk-copy producer)
  if (P.use_tma && !no_tmap && p->edge_tp % 32 == 0 && make_plane_map(&tmc, d_class, (size_t)B * C, N, C, p->edge_tp) &&
      make_plane_map(&tma, d_adj, (size_t)B * K, N, K, p->edge_tp))
    P.use_tmap = 1;
  const bool warp_pipeline = p->edge2_ncons > 0 && P.use_tma && sdb == 0.0f;
  // the drop-in symbol takes whatever floats the caller hands in (the reference computes libm's special values
  // for them): check the domain of the fast kernel's log recipes and let the general kernel redo the batch if needed
  const bool check_domain = warp_pipeline && !P.clip && (clip & MN_INPUT_CHECK_DOMAIN);
  if (check_domain) {
    MN_CUDA_OK(cudaMemsetAsync(p->d_domain_flag, 0, sizeof(int), s));
    mn_domain_check_kernel<<<p->num_sms * 8, 256, 0, s>>>(d_class, (size_t)B * C * N, d_adj, (size_t)B * K * N, p->d_domain_flag);
    p->timings.other_launches++;
  }
  if (warp_pipeline) {
    P.TP = 32 * p->edge2_ncons;
    P.stages = p->edge2_stages;
    P.ws0_clp = p->h_imgs[b0].clp; P.ws0_same = p->h_imgs[b0].rec_same; P.ws0_diff = p->h_imgs[b0].rec_diff; P.ws0_cls = p->h_imgs[b0].cls;
    P.ws_stride = p->per_image_bytes;
    P.tiles_per_image = (N + P.TP - 1) / P.TP;
  }
  long long tiles = (long long)B * P.tiles_per_image;
  if (tiles >= (1ll << 31)) { g_last_error = MN_STATUS_BAD_ARG; return MN_STATUS_BAD_ARG; }
  if (warp_pipeline) {
    int grid = (int)std::min<long long>(tiles, (long long)p->num_sms * p->edge2_ctas);
    int threads = 32 * (p->edge2_ncons + 1);
    if (P.logits) mn_edge_warp_kernel<2><<<grid, threads, p->edge2_smem, s>>>(P);
    else if (P.clip) mn_edge_warp_kernel<1><<<grid, threads, p->edge2_smem, s>>>(P);
    else mn_edge_warp_kernel<0><<<grid, threads, p->edge2_smem, s>>>(P);
    if (check_domain) {  // fix-up: the general kernel, which returns at once unless the flag was raised
      MnEdgeParams Q = P;
      Q.TP = p->edge_tp;
      Q.tiles_per_image = (N + Q.TP - 1) / Q.TP;
      Q.only_if = p->d_domain_flag;
      const long long qt = (long long)B * Q.tiles_per_image;
      int qgrid = (int)std::min<long long>(qt, (long long)p->num_sms * MN_EDGE_CTAS_PER_SM);
      mn_edge_pass_kernel<<<qgrid, MN_EDGE_THREADS, p->edge_smem, s>>>(Q, tmc, tma);
      p->timings.edge_launches++;
    }
  } else {
    int grid = (int)std::min<long long>(tiles, (long long)p->num_sms * MN_EDGE_CTAS_PER_SM);
    mn_edge_pass_kernel<<<grid, MN_EDGE_THREADS, p->edge_smem, s>>>(P, tmc, tma);
  }
  p->timings.edge_launches++;
  return MN_STATUS_OK;
}

// edge pass + record init + sort for images [b0, b0 + B); d_class / d_adj point at image b0's maps
static int run_front(mn_plan* p, int b0, int B, const float* d_class, float* d_adj, int clip, float sdb, float omf,
                     float mlb, cudaStream_t s, bool sort_keys, bool record_events = true) {
  const int N = p->N, C = p->C, K = p->K;
  {
    dim3 g((unsigned)std::min<long long>(4096, (p->h_imgs[0].hash_nbuckets * 8ll + 255) / 256), (unsigned)std::min(B, 65535));
    mn_reset_kernel<<<g, 256, 0, s>>>(p->d_imgs + b0, B);
    p->timings.other_launches++;
  }
  if (record_events) MN_CUDA_OK(cudaEventRecord(p->ev[1], s));
  if (int rc = launch_edge(p, b0, B, d_class, d_adj, clip, sdb, s)) return rc;
  if (record_events) MN_CUDA_OK(cudaEventRecord(p->ev[2], s));
  for (int b = b0; b < b0 + B; b++) {
    MnRecInitParams R;
    R.im = p->h_imgs[b];
    R.keys_out = p->d_keys_scratch;  // (never init_keys: it shares the arena with the inputs rec_same | rec_diff)
    R.H = p->H; R.W = p->W; R.C = C; R.K = K; R.N = N; R.omf = omf; R.mlb = mlb;
    for (int k = 0; k < K; k++) { R.off_r[k] = p->offsets[2 * k]; R.off_c[k] = p->offsets[2 * k + 1]; R.rank_of_k[k] = p->rank_of_k[k]; }
    int g = (int)std::min<long long>((p->E + 255) / 256, (long long)p->num_sms * 16);
    mn_record_init_kernel<<<g, 256, 0, s>>>(R);
    p->timings.other_launches++;
    if (sort_keys) {
      size_t tb = p->cub_temp_bytes;
      MN_CUDA_OK(cub::DeviceRadixSort::SortKeys(p->d_cub_temp, tb, (const uint64_t*)p->d_keys_scratch,
                                                p->h_imgs[b].init_keys, (int)p->E, 0, 32 + MN_ORD_BITS, s));
      // (library kernels of one 60-bit sort: histogram + exclusive sum + 8 onesweep passes, profiles/r01_launches_*.csv)
      p->timings.other_launches += 2 + (32 + MN_ORD_BITS + 7) / 8;
    }
  }
  if (record_events) MN_CUDA_OK(cudaEventRecord(p->ev[3], s));
  MN_CUDA_OK(cudaGetLastError());
  return MN_STATUS_OK;
}

static int run_back(mn_plan* p, int B, const float* d_class, const float* d_adj, int clip, int* d_mask, int* d_object_class, int* d_ninst,
                    float omf, float mlb, cudaStream_t s);

extern "C" int mn_segment_batch_device(mn_plan* p, int B, const float* d_class, float* d_adj, int* d_mask,
                                       int* d_object_class, int* d_ninst, int clip, float sdb, float omf,
                                       float mlb, void* stream) {
  g_last_error = MN_STATUS_OK;
  if (!p || B <= 0 || B > p->max_batch || !d_class || !d_adj || !d_mask || !d_object_class || !d_ninst) {
    g_last_error = MN_STATUS_BAD_ARG;
    return MN_STATUS_BAD_ARG;
  }
  MN_CUDA_OK(cudaSetDevice(p->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : p->stream;
  p->timings.edge_launches = 0; p->timings.other_launches = 0;
  int rc = run_front(p, 0, B, d_class, d_adj, clip, sdb, omf, mlb, s, true);
  if (rc) return rc;
  return run_back(p, B, d_class, d_adj, clip, d_mask, d_object_class, d_ninst, omf, mlb, s);
}

// merge scheduler + labels + partition statistics for images [0, B), whose front end has been enqueued on s
static int run_back(mn_plan* p, int B, const float* d_class, const float* d_adj, int clip, int* d_mask, int* d_object_class, int* d_ninst,
                    float omf, float mlb, cudaStream_t s) {
  const int N = p->N;
  MnMergeArgs A;
  memset(&A, 0, sizeof(A));
  A.C = p->C; A.K = p->K; A.N = N; A.W = p->W; A.omf = omf; A.mlb = mlb; A.off = p->off; A.H = p->merge_H; A.max_rounds = 40ll * N + 100000;  // guard: rounds <= events, a few per pixel
  int grid = std::min(B, p->num_sms);
  mn_merge_kernel<<<grid, MN_MERGE_THREADS, p->merge_smem, s>>>(p->d_imgs, B, A);
  p->timings.other_launches++;
  MN_CUDA_OK(cudaEventRecord(p->ev[4], s));
  // labels
  {
    dim3 g((unsigned)std::min(1024, (N + 255) / 256), (unsigned)std::min(B, 65535));
    mn_label_flags_kernel<<<g, 256, 0, s>>>(p->d_imgs, B, N);
    for (int b = 0; b < B; b++) {
      size_t tb = p->cub_temp_bytes;
      MN_CUDA_OK(cub::DeviceScan::ExclusiveSum(p->d_cub_temp, tb, (const int*)p->h_imgs[b].cls, p->h_imgs[b].pix_pool, N, s));
      p->timings.other_launches += 2;  // (scan init + scan)
    }
    MN_CUDA_OK(cudaMemsetAsync(d_object_class, 0xFF, (size_t)B * N * 4, s));
    mn_label_write_kernel<<<g, 256, 0, s>>>(p->d_imgs, B, N, d_mask, d_object_class, d_ninst);
    p->timings.other_launches += 2;
  }
  MN_CUDA_OK(cudaEventRecord(p->ev[8], s));
  // partition statistics pass: total log-prob of the final partition, from the maps and the label mask
  {
    MnLogprobParams L;
    L.d_class = d_class; L.d_adj = d_adj; L.d_mask = d_mask; L.d_object_class = d_object_class; L.partial = p->d_logprob_partial;
    L.nimg = B; L.H = p->H; L.W = p->W; L.C = p->C; L.K = p->K; L.N = N; L.maxr = 0; L.maxc = 0;
    for (int k = 0; k < p->K; k++) {
      L.off_r[k] = p->offsets[2 * k]; L.off_c[k] = p->offsets[2 * k + 1]; L.delta[k] = L.off_r[k] * p->W + L.off_c[k];
      L.maxr = std::max(L.maxr, std::abs(L.off_r[k])); L.maxc = std::max(L.maxc, std::abs(L.off_c[k]));
    }
    const int gx = std::max(1, std::min(MN_LOGPROB_BLOCKS, (N + 1023) / 1024));
    dim3 g((unsigned)gx, (unsigned)std::min(B, 65535));
    const int mode = clip & (MN_INPUT_CLIP | MN_INPUT_LOGITS);
    if (mode & MN_INPUT_LOGITS) mn_launch_partition_logprob<MN_INPUT_LOGITS>(L, g, s);
    else if (mode & MN_INPUT_CLIP) mn_launch_partition_logprob<MN_INPUT_CLIP>(L, g, s);
    else mn_launch_partition_logprob<0>(L, g, s);
    mn_partition_logprob_fold_kernel<<<B, 96, 0, s>>>(p->d_logprob_partial, B, gx, p->d_logprob);
    p->timings.other_launches += 2;
    p->last_omf = omf;
  }
  MN_CUDA_OK(cudaEventRecord(p->ev[5], s));
  MN_CUDA_OK(cudaGetLastError());
  // per-image status / statistics
  for (int b = 0; b < B; b++)
    MN_CUDA_OK(cudaMemcpyAsync(&p->h_ctl[b], p->h_imgs[b].ctl, sizeof(MnCtl), cudaMemcpyDeviceToHost, s));
  MN_CUDA_OK(cudaMemcpyAsync(p->h_logprob.data(), p->d_logprob, sizeof(double) * 4 * B, cudaMemcpyDeviceToHost, s));
  MN_CUDA_OK(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, p->ev[1], p->ev[2]); p->timings.edge_ms = ms;
  cudaEventElapsedTime(&ms, p->ev[2], p->ev[3]); p->timings.record_init_sort_ms = ms;
  cudaEventElapsedTime(&ms, p->ev[3], p->ev[4]); p->timings.merge_ms = ms;
  cudaEventElapsedTime(&ms, p->ev[4], p->ev[8]); p->timings.label_ms = ms;
  cudaEventElapsedTime(&ms, p->ev[8], p->ev[5]); p->timings.aggregate_ms = ms;
  cudaEventElapsedTime(&ms, p->ev[1], p->ev[5]); p->timings.total_ms = ms;
  int worst = MN_STATUS_OK;
  for (int b = 0; b < B; b++)
    if (p->h_ctl[b].status != MN_OK && worst == MN_STATUS_OK) worst = p->h_ctl[b].status;
  g_last_error = worst;
  return worst;
}

static int ensure_staging(mn_plan* p, int B) {
  if ((size_t)B <= p->staging_batch) return MN_STATUS_OK;
  cudaFree(p->d_in_class); cudaFree(p->d_in_adj); cudaFree(p->d_out_mask); cudaFree(p->d_out_cls); cudaFree(p->d_out_ninst);
  p->d_in_class = nullptr; p->d_in_adj = nullptr; p->d_out_mask = nullptr; p->d_out_cls = nullptr; p->d_out_ninst = nullptr;
  p->staging_batch = 0;
  const size_t N = p->N;
  MN_CUDA_OK(cudaMalloc(&p->d_in_class, (size_t)B * p->C * N * 4));
  MN_CUDA_OK(cudaMalloc(&p->d_in_adj, (size_t)B * p->K * N * 4));
  MN_CUDA_OK(cudaMalloc(&p->d_out_mask, (size_t)B * N * 4));
  MN_CUDA_OK(cudaMalloc(&p->d_out_cls, (size_t)B * N * 4));
  MN_CUDA_OK(cudaMalloc(&p->d_out_ninst, (size_t)B * 4));
  p->staging_batch = B;
  return MN_STATUS_OK;
}

extern "C" int mn_segment_batch_host(mn_plan* p, int B, const float* h_class, float* h_adj, int* h_mask,
                                     int* h_object_class, int* h_ninst, int clip, float sdb, float omf,
                                     float mlb) {
  g_last_error = MN_STATUS_OK;
  if (!p || B <= 0 || B > p->max_batch || !h_class || !h_adj || !h_mask || !h_object_class) {
    g_last_error = MN_STATUS_BAD_ARG;
    return MN_STATUS_BAD_ARG;
  }
  MN_CUDA_OK(cudaSetDevice(p->device));
  int rc = ensure_staging(p, B);
  if (rc) return rc;
  const size_t N = p->N;
  cudaStream_t s = p->stream;
  // The uploads run on a second stream, in chunks of a few images, while the front end (edge pass,
  // record init, sort) of the previous chunk runs on the plan's stream: only the first chunk's copy is exposed.
  if (!p->copy_stream) {
    MN_CUDA_OK(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < MN_COPY_EVENTS; i++) MN_CUDA_OK(cudaEventCreateWithFlags(&p->copy_ev[i], cudaEventDisableTiming));
  }
  p->timings.edge_launches = 0; p->timings.other_launches = 0;
  MN_CUDA_OK(cudaEventRecord(p->ev[0], s));
  MN_CUDA_OK(cudaStreamWaitEvent(p->copy_stream, p->ev[0], 0));  // (the staging buffers are free: earlier work on s is done)
  const int chunk = std::max(8, (B + MN_COPY_EVENTS - 1) / MN_COPY_EVENTS);  // 8 images (1.3 GB at 1024x2048), at most MN_COPY_EVENTS chunks
  int ci = 0;
  for (int b0 = 0; b0 < B; b0 += chunk, ci++) {
    const int nb = std::min(chunk, B - b0);
    const size_t oc = (size_t)b0 * p->C * N, oa = (size_t)b0 * p->K * N;
    MN_CUDA_OK(cudaMemcpyAsync(p->d_in_class + oc, h_class + oc, (size_t)nb * p->C * N * 4, cudaMemcpyHostToDevice, p->copy_stream));
    MN_CUDA_OK(cudaMemcpyAsync(p->d_in_adj + oa, h_adj + oa, (size_t)nb * p->K * N * 4, cudaMemcpyHostToDevice, p->copy_stream));
    MN_CUDA_OK(cudaEventRecord(p->copy_ev[ci], p->copy_stream));
    MN_CUDA_OK(cudaStreamWaitEvent(s, p->copy_ev[ci], 0));
    if (b0 == 0) MN_CUDA_OK(cudaEventRecord(p->ev[1], s));  // h2d_ms = the exposed part of the upload
    rc = run_front(p, b0, nb, p->d_in_class + oc, p->d_in_adj + oa, clip, sdb, omf, mlb, s, true, false);
    if (rc) return rc;
  }
  MN_CUDA_OK(cudaEventRecord(p->ev[2], s));  // (chunked: edge_ms covers the whole front end, record_init_sort_ms is 0)
  MN_CUDA_OK(cudaEventRecord(p->ev[3], s));
  rc = run_back(p, B, p->d_in_class, p->d_in_adj, clip, p->d_out_mask, p->d_out_cls, p->d_out_ninst, omf, mlb, s);
  MN_CUDA_OK(cudaEventRecord(p->ev[6], s));
  MN_CUDA_OK(cudaMemcpyAsync(h_mask, p->d_out_mask, (size_t)B * N * 4, cudaMemcpyDeviceToHost, s));
  MN_CUDA_OK(cudaMemcpyAsync(h_object_class, p->d_out_cls, (size_t)B * N * 4, cudaMemcpyDeviceToHost, s));
  if (h_ninst) MN_CUDA_OK(cudaMemcpyAsync(h_ninst, p->d_out_ninst, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
  if (sdb != 0.0f)  // the reference rewrites the caller's buffer (segment.cc:187-191)
    MN_CUDA_OK(cudaMemcpyAsync(h_adj, p->d_in_adj, (size_t)B * p->K * N * 4, cudaMemcpyDeviceToHost, s));
  MN_CUDA_OK(cudaEventRecord(p->ev[7], s));
  MN_CUDA_OK(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, p->ev[0], p->ev[1]); p->timings.h2d_ms = ms;
  cudaEventElapsedTime(&ms, p->ev[6], p->ev[7]); p->timings.d2h_ms = ms;
  g_last_error = rc;
  return rc;
}

extern "C" int mn_plan_image_stats(mn_plan* p, int image, mn_image_stats* o) {
  if (!p || !o || image < 0 || image >= p->max_batch) return MN_STATUS_BAD_ARG;
  const MnCtl& c = p->h_ctl[image];
  o->status = c.status; o->fail_line = c.fail_line; o->n_instances = c.n_instances; o->n_init_entries = c.n_init;
  o->rounds = c.rounds; o->events = c.events; o->merges = c.merges; o->restores = c.restores;
  o->invalid_pops = c.invalid_pops; o->solo_events = c.solo_events; o->refills = c.refills;
  o->flushes = c.flushes; o->splits = c.splits; o->pairs = c.pairs; o->cuts_conflict = c.cuts_conflict;
  o->cuts_cascade = c.cuts_cascade; o->cuts_capacity = c.cuts_capacity;
  for (int i = 0; i < 16; i++) o->cycles[i] = c.cyc[i];
  o->cycles_total = c.cycles_total;
  o->queue_chunks_used = c.qc_bump; o->pixel_pool_used = c.pix_bump; o->tree_nodes_used = c.tn_bump;
  o->requeues = c.requeues; o->hash_overflow = c.hash_ovf_n; o->pixel_pool_collections = c.pix_gcs;
  return MN_STATUS_OK;
}
extern "C" int mn_plan_image_logprob(mn_plan* p, int image, double* out4) {
  if (!p || !out4 || image < 0 || image >= p->max_batch) return MN_STATUS_BAD_ARG;
  const double* t = &p->h_logprob[(size_t)4 * image];
  out4[0] = t[0]; out4[1] = t[1]; out4[2] = t[2];
  out4[3] = t[0] + (double)p->last_omf * (t[2] + t[1]);  // cc:314-350
  return MN_STATUS_OK;
}
extern "C" int mn_plan_timings(mn_plan* p, mn_timings* o) {
  if (!p || !o) return MN_STATUS_BAD_ARG;
  *o = p->timings;
  return MN_STATUS_OK;
}

// ------------------------------------------------------------------------------------------------
// drop-in symbol (segment.cc:742-752): one cached plan per thread AND device, re-created when the shape changes
struct PlanCache {
  mn_plan* plan = nullptr;
  int H = 0, W = 0, C = 0, K = 0, device = -1;
  int offsets[2 * MN_MAX_K];
  void release() {
    if (!plan) return;
    // at process teardown the CUDA runtime may already be unloading: its calls then fail with
    // cudaErrorCudartUnloading and there is nothing left to free; in a live process (a worker thread that
    // exits, mn_shutdown()) the workspace goes back to the device
    int d = 0;
    if (cudaGetDevice(&d) == cudaSuccess) mn_plan_destroy(plan);
    plan = nullptr;
  }
  ~PlanCache() { release(); }
};
static thread_local PlanCache g_cache;

// scratch of the host-buffer post-pass entries (mask resize, COCO RLE): per thread, grow-only, own stream
struct PostScratch {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  void* buf[3] = {nullptr, nullptr, nullptr};
  size_t cap[3] = {0, 0, 0};
  void release() {
    int d = 0;
    if (cudaGetDevice(&d) == cudaSuccess && device >= 0) {
      for (int i = 0; i < 3; i++) cudaFree(buf[i]);
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
      if (stream) cudaStreamDestroy(stream);
    }
    for (int i = 0; i < 3; i++) { buf[i] = nullptr; cap[i] = 0; }
    stream = nullptr; e0 = e1 = nullptr; device = -1;
  }
  // buffers of at least the given sizes on the caller's current device
  bool ensure(size_t b0, size_t b1, size_t b2) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) return false;
    if (device != d) {
      release();
      if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) { stream = nullptr; return false; }
      if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { release(); return false; }
      device = d;
    }
    const size_t need[3] = {b0, b1, b2};
    for (int i = 0; i < 3; i++) {
      if (need[i] <= cap[i]) continue;
      cudaFree(buf[i]); buf[i] = nullptr; cap[i] = 0;
      if (cudaMalloc(&buf[i], need[i]) != cudaSuccess) return false;
      cap[i] = need[i];
    }
    return true;
  }
  ~PostScratch() { release(); }
};
static thread_local PostScratch g_post;

extern "C" void mn_shutdown(void) {
  g_cache.release();
  g_post.release();
}

extern "C" void c_run_segmentation(float* class_pred, int class_dim, float* adj_pred, int offset_dim,
                                   int img_width, int img_height, int num_classes, int* offset_list,
                                   int* output, int* object_class, float sdb, float omf, float mlb) {
  g_last_error = MN_STATUS_OK;
  const long long N = (long long)img_width * img_height;
  if (output && N > 0) memset(output, 0, sizeof(int) * (size_t)N);
  if (object_class && N > 0) memset(object_class, 0xFF, sizeof(int) * (size_t)N);
  auto fail = [&](int code) {
    // The symbol returns void, like the reference's (which exit(1)s on its internal errors, cc:40-43,666-673):
    // a caller that links it directly cannot see a status, so the failure is also reported on stderr.
    g_last_error = code;
    fprintf(stderr, "mergenet_b200: c_run_segmentation failed: %s (status %d); outputs left empty (mask 0, classes -1)\n",
            mn_status_string(code), code);
  };
  if (!class_pred || !adj_pred || !offset_list || !output || !object_class || class_dim != num_classes ||
      offset_dim <= 0 || offset_dim > MN_MAX_K) {
    fail(MN_STATUS_BAD_ARG);
    return;
  }
  // MN_TIE_ORDER=reference in the environment: the tie-exact replay (mn_exact_segment_host) instead of the hot path, for
  // callers that link this symbol directly (the reference's own c_segment.pyx) and need the reference's RAW arrays also on
  // inputs whose partition depends on the order among exactly equal priorities.  Sequential; anything else: the hot path.
  if (const char* tie = getenv("MN_TIE_ORDER")) {
    if (strcmp(tie, "reference") == 0) {
      int ninst_exact = 0;
      const int rc = mn_exact_segment_host(class_pred, num_classes, adj_pred, offset_dim, img_height, img_width, offset_list,
                                           0, sdb, omf, mlb, output, object_class, &ninst_exact, nullptr);
      if (rc != MN_STATUS_OK) {
        if (N > 0) {
          memset(output, 0, sizeof(int) * (size_t)N);
          memset(object_class, 0xFF, sizeof(int) * (size_t)N);
        }
        fail(rc);
      }
      return;
    }
  }
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess) { fail(MN_STATUS_CUDA); return; }  // the caller's current device
  PlanCache& c = g_cache;
  bool same = c.plan && c.device == device && c.H == img_height && c.W == img_width && c.C == num_classes && c.K == offset_dim &&
              memcmp(c.offsets, offset_list, sizeof(int) * 2 * offset_dim) == 0;
  if (!same) {
    c.release();
    int rc = mn_plan_create(&c.plan, 1, img_height, img_width, num_classes, offset_dim, offset_list, device);
    if (rc) { c.plan = nullptr; fail(rc); return; }
    c.H = img_height; c.W = img_width; c.C = num_classes; c.K = offset_dim; c.device = device;
    memcpy(c.offsets, offset_list, sizeof(int) * 2 * offset_dim);
  }
  int ninst = 0;
  mn_segment_batch_host(c.plan, 1, class_pred, adj_pred, output, object_class, &ninst, MN_INPUT_CHECK_DOMAIN, sdb, omf, mlb);
  cudaSetDevice(device);  // (the plan's device is the caller's; restated for symmetry with the batch entries)
  if (g_last_error != MN_STATUS_OK) {
    memset(output, 0, sizeof(int) * (size_t)N);
    memset(object_class, 0xFF, sizeof(int) * (size_t)N);
    fail(g_last_error);
  }
}

// ------------------------------------------------------------------------------------------------
// the step after the path: masks at the image size, COCO RLE (mn_post.cuh)
static thread_local float g_post_ms = 0.f;
extern "C" float mn_post_last_ms(void) { return g_post_ms; }

extern "C" int mn_resize_masks_nearest_device(const int* d_in, int B, int H, int W, int* d_out, int OH, int OW, void* stream) {
  g_last_error = MN_STATUS_OK;
  if (!d_in || !d_out || B <= 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0) { g_last_error = MN_STATUS_BAD_ARG; return MN_STATUS_BAD_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  MN_CUDA_OK(mn_resize_nearest_launch(d_in, B, H, W, d_out, OH, OW, (cudaStream_t)stream));
  return MN_STATUS_OK;
}

extern "C" int mn_resize_masks_nearest_host(const int* h_in, int B, int H, int W, int* h_out, int OH, int OW) {
  g_last_error = MN_STATUS_OK;
  if (!h_in || !h_out || B <= 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0) { g_last_error = MN_STATUS_BAD_ARG; return MN_STATUS_BAD_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  const size_t nin = (size_t)B * H * W * 4, nout = (size_t)B * OH * OW * 4;
  PostScratch& ps = g_post;  // scratch and stream are kept between calls (per thread, current device)
  if (!ps.ensure(nin, nout, 0)) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  int* di = (int*)ps.buf[0]; int* dout = (int*)ps.buf[1];
  cudaStream_t st = ps.stream;
  cudaError_t e = cudaMemcpyAsync(di, h_in, nin, cudaMemcpyHostToDevice, st);
  cudaEventRecord(ps.e0, st);
  if (e == cudaSuccess) e = mn_resize_nearest_launch(di, B, H, W, dout, OH, OW, st);
  cudaEventRecord(ps.e1, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, dout, nout, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) cudaEventElapsedTime(&g_post_ms, ps.e0, ps.e1);
  if (e != cudaSuccess) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  return MN_STATUS_OK;
}

extern "C" int mn_resize_maps_bilinear_device(const float* d_in, long long planes, int H, int W, float* d_out, int OH, int OW, void* stream) {
  g_last_error = MN_STATUS_OK;
  if (!d_in || !d_out || planes <= 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0) { g_last_error = MN_STATUS_BAD_ARG; return MN_STATUS_BAD_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  MN_CUDA_OK(mn_resize_bilinear_launch(d_in, planes, H, W, d_out, OH, OW, (cudaStream_t)stream));
  return MN_STATUS_OK;
}
extern "C" int mn_resize_maps_bilinear_host(const float* h_in, long long planes, int H, int W, float* h_out, int OH, int OW) {
  g_last_error = MN_STATUS_OK;
  if (!h_in || !h_out || planes <= 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0) { g_last_error = MN_STATUS_BAD_ARG; return MN_STATUS_BAD_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  const size_t nin = (size_t)planes * H * W * 4, nout = (size_t)planes * OH * OW * 4;
  PostScratch& ps = g_post;
  if (!ps.ensure(nin, nout, 0)) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  float* di = (float*)ps.buf[0]; float* dout = (float*)ps.buf[1];
  cudaStream_t st = ps.stream;
  cudaError_t e = cudaMemcpyAsync(di, h_in, nin, cudaMemcpyHostToDevice, st);
  cudaEventRecord(ps.e0, st);
  if (e == cudaSuccess) e = mn_resize_bilinear_launch(di, planes, H, W, dout, OH, OW, st);
  cudaEventRecord(ps.e1, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, dout, nout, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) cudaEventElapsedTime(&g_post_ms, ps.e0, ps.e1);
  if (e != cudaSuccess) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  return MN_STATUS_OK;
}

extern "C" int mn_mask_to_coco_rle_host(const int* h_mask, int H, int W, int n, unsigned char* counts, long long cap,
                                        long long* offsets) {
  g_last_error = MN_STATUS_OK;
  if (!h_mask || !offsets || H <= 0 || W <= 0 || n < 0 || cap < 0 || (cap > 0 && !counts) || (long long)H * W >= (1ll << 31)) {
    g_last_error = MN_STATUS_BAD_ARG;
    return MN_STATUS_BAD_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  const size_t a = (size_t)H * W;
  PostScratch& ps = g_post;
  if (!ps.ensure(a * 4, (size_t)(cap > 0 ? cap : 1), ((size_t)n + 1) * 8)) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  int* dm = (int*)ps.buf[0]; unsigned char* dc = (unsigned char*)ps.buf[1]; long long* dofs = (long long*)ps.buf[2];
  cudaStream_t st = ps.stream;
  long long total = 0;
  cudaError_t e = cudaMemcpyAsync(dm, h_mask, a * 4, cudaMemcpyHostToDevice, st);
  cudaEventRecord(ps.e0, st);
  if (e == cudaSuccess) e = mn_coco_rle_device(dm, H, W, n, dc, cap, dofs, &total, st);
  cudaEventRecord(ps.e1, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(offsets, dofs, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && total <= cap && total > 0) e = cudaMemcpyAsync(counts, dc, (size_t)total, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) cudaEventElapsedTime(&g_post_ms, ps.e0, ps.e1);
  if (e != cudaSuccess) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  if (total > cap) { offsets[n] = total; g_last_error = MN_STATUS_BAD_ARG; return MN_STATUS_BAD_ARG; }
  return MN_STATUS_OK;
}

// ------------------------------------------------------------------------------------------------
// Mode B: the reference's pure-Python segmenter semantics (utils/segmenter.py), strictly sequential (mn_modeb.cuh)
__global__ void mn_modeb_kernel(MnModeB m) {
  if (threadIdx.x == 0 && blockIdx.x == 0) mnb_run(m);
}

extern "C" int mn_modeb_segment_host(const float* h_logc, const float* h_lsame, const float* h_ldiff, int C, int K, int H,
                                     int W, const int* offset_list, double omf, double mlb, double prune_threshold,
                                     long long* h_mask, int* h_object_class, int* n_instances, long long* stats4) {
  g_last_error = MN_STATUS_OK;
  if (!h_logc || !h_lsame || !h_ldiff || !offset_list || !h_mask || !h_object_class || !n_instances || C <= 0 ||
      C >= MN_MAX_C || K <= 0 || K > MN_MAX_K || H <= 0 || W <= 0 || (long long)H * W * K > (1ll << 26)) {
    g_last_error = MN_STATUS_BAD_ARG;
    return MN_STATUS_BAD_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  const size_t N = (size_t)H * W, E = N * K;
  MnModeB m;
  memset(&m, 0, sizeof(m));
  m.C = C; m.K = K; m.H = H; m.W = W; m.N = (int)N; m.E = (long long)E;
  for (int k = 0; k < K; k++) { m.off_r[k] = offset_list[2 * k]; m.off_c[k] = offset_list[2 * k + 1]; }
  m.omf = omf; m.mlb = mlb; m.omf32 = (float)omf; m.mlb32 = (float)mlb; m.prune_threshold = prune_threshold;
  unsigned hm = 1;
  while ((size_t)hm < 2 * E + 16) hm <<= 1;
  m.h_mask = hm - 1;
  m.q_cap = (long long)(8 * E + 1024);
  // one allocation, carved up (256-byte aligned pieces)
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 255) / 256 * 256; return at; };
  const size_t a_logc = take(N * C * 4), a_ls = take(E * 4), a_ld = take(E * 4), a_npix = take(N * 4), a_cls = take(N * 4),
               a_clp = take(N * C * 8), a_osame = take(N * 4), a_alive = take(N), a_head = take(N * 4), a_tail = take(N * 4),
               a_pnext = take(N * 4), a_ptail = take(N * 4), a_o1 = take(E * 4), a_o2 = take(E * 4), a_oml = take(E * 4),
               a_same = take(E * 4), a_diff = take(E * 4), a_mp = take(E * 8), a_link = take(E * 24),
               a_hk = take((size_t)hm * 8), a_hv = take((size_t)hm * 4), a_qk = take((size_t)m.q_cap * 8),
               a_qr = take((size_t)m.q_cap * 4), a_mask = take(N * 8), a_ocls = take(N * 4), a_n = take(4),
               a_status = take(4), a_stats = take(64);
  unsigned char* d = nullptr;
  if (cudaMalloc(&d, o) != cudaSuccess) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  auto done = [&](int code) { cudaFree(d); g_last_error = code; return code; };
  m.logc = (const float*)(d + a_logc); m.lsame = (const float*)(d + a_ls); m.ldiff = (const float*)(d + a_ld);
  m.npix = (int*)(d + a_npix); m.cls = (int*)(d + a_cls); m.clp = (double*)(d + a_clp); m.osame = (float*)(d + a_osame);
  m.alive = d + a_alive; m.adj_head = (int*)(d + a_head); m.adj_tail = (int*)(d + a_tail);
  m.pix_next = (int*)(d + a_pnext); m.pix_tail = (int*)(d + a_ptail);
  m.r_o1 = (int*)(d + a_o1); m.r_o2 = (int*)(d + a_o2); m.r_oml = (float*)(d + a_oml); m.r_same = (float*)(d + a_same);
  m.r_diff = (float*)(d + a_diff); m.r_mp = (double*)(d + a_mp); m.r_link = (int*)(d + a_link);
  m.h_key = (unsigned long long*)(d + a_hk); m.h_val = (int*)(d + a_hv);
  m.q_key = (double*)(d + a_qk); m.q_rec = (int*)(d + a_qr); m.q_n = 0;
  m.out_mask = (long long*)(d + a_mask); m.out_cls = (int*)(d + a_ocls); m.out_n = (int*)(d + a_n);
  m.status = (int*)(d + a_status); m.stats = (long long*)(d + a_stats);
  cudaError_t e = cudaMemcpy((void*)m.logc, h_logc, N * C * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy((void*)m.lsame, h_lsame, E * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy((void*)m.ldiff, h_ldiff, E * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(m.h_key, 0xFF, (size_t)hm * 8);
  if (e == cudaSuccess) e = cudaMemset(d + a_n, 0, 4 + 252 + 4 + 252 + 64);  // n, status, stats (adjacent pieces)
  if (e == cudaSuccess) e = cudaMemset(m.out_cls, 0xFF, N * 4);
  if (e != cudaSuccess) return done(MN_STATUS_CUDA);
  mn_modeb_kernel<<<1, 32>>>(m);
  int status = 0;
  long long st[8] = {0};
  e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(&status, m.status, 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(st, m.stats, 32, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(n_instances, m.out_n, 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(h_mask, m.out_mask, N * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(h_object_class, m.out_cls, N * 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return done(MN_STATUS_CUDA);
  if (stats4) for (int i = 0; i < 4; i++) stats4[i] = st[i];
  if (status == 1) return done(MN_STATUS_Q_POOL);
  if (status == 2) return done(MN_STATUS_HASH_FULL);
  if (status == 3) return done(MN_STATUS_NO_BACKGROUND);
  return done(MN_STATUS_OK);
}

// ------------------------------------------------------------------------------------------------
// Tie-exact replay of the reference's C++ segmenter (mn_exact.cuh): the edge pass of the hot path, then ONE thread
// that replays segment.cc:539-727 with libstdc++'s heap and hash-table orders
__global__ void mn_exact_kernel(MnExact m) {
  if (threadIdx.x == 0 && blockIdx.x == 0) mnx_run(m);
}

extern "C" int mn_exact_segment_host(const float* h_class, int C, float* h_adj, int K, int H, int W, const int* offset_list,
                                     int clip, float sdb, float omf, float mlb, int* h_mask, int* h_object_class,
                                     int* n_instances, long long* stats4) {
  g_last_error = MN_STATUS_OK;
  if (!h_class || !h_adj || !offset_list || !h_mask || !h_object_class || !n_instances) {
    g_last_error = MN_STATUS_BAD_ARG;
    return MN_STATUS_BAD_ARG;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }  // the caller's current device
  mn_plan* p = nullptr;
  int rc = mn_plan_create(&p, 1, H, W, C, K, offset_list, dev);  // (checks the shape; its workspace receives the edge pass)
  if (rc) return rc;
  unsigned char* d = nullptr;
  auto done = [&](int code) { if (d) cudaFree(d); mn_plan_destroy(p); g_last_error = code; return code; };
  if (ensure_staging(p, 1)) return done(MN_STATUS_CUDA);
  const size_t N = p->N, E = (size_t)p->E;
  cudaStream_t s = p->stream;
  MnExact m;
  memset(&m, 0, sizeof(m));
  m.C = C; m.K = K; m.H = H; m.W = W; m.N = (int)N; m.E = (long long)E; m.omf = omf; m.mlb = mlb;
  for (int k = 0; k < K; k++) { m.off_r[k] = offset_list[2 * k]; m.off_c[k] = offset_list[2 * k + 1]; }
  m.heap.cap = (long long)(8 * E + 1024);  // (the reference's own queue holds 3.2 E entries over a whole 1024 x 2048 run)
  m.arena.half = (long long)(32 * N + 5 * E + 4096);
  if (const char* w = getenv("MN_EXACT_ARENA_WORDS")) {  // (test hook: a small half-space, so that the collection runs on the device too)
    const long long v = atoll(w);
    if (v > 0 && v < m.arena.half) m.arena.half = v;
  }
  m.arena.ntabs = (int)N + 1;
  // one allocation, carved up (256-byte aligned pieces)
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 255) / 256 * 256; return at; };
  const size_t a_npix = take(N * 4), a_cls = take(N * 4), a_pnext = take(N * 4), a_ptail = take(N * 4),
               a_tab = take((N + 1) * sizeof(MnStlTab)), a_onext = take(N * 4), a_o1 = take(E * 4), a_o2 = take(E * 4),
               a_oml = take(E * 4), a_mp = take(E * 4), a_merged = take(E * 4), a_ndn = take(2 * E * 4),
               a_ndk = take(2 * E * 8), a_qk = take((size_t)m.heap.cap * 4), a_qr = take((size_t)m.heap.cap * 4),
               a_bk = take((size_t)m.arena.half * 2 * 4), a_primes = take(MN_STL_NPRIMES * 4), a_mask = take(N * 4),
               a_ocls = take(N * 4), a_scalars = take(256);
  if (cudaMalloc(&d, o) != cudaSuccess) { d = nullptr; return done(MN_STATUS_CUDA); }
  m.npix = (int*)(d + a_npix); m.cls = (int*)(d + a_cls); m.pix_next = (int*)(d + a_pnext); m.pix_tail = (int*)(d + a_ptail);
  m.tab = (MnStlTab*)(d + a_tab); m.ob_next = (int*)(d + a_onext);
  m.r_o1 = (int*)(d + a_o1); m.r_o2 = (int*)(d + a_o2); m.r_oml = (float*)(d + a_oml); m.r_mp = (float*)(d + a_mp);
  m.r_merged = (int*)(d + a_merged); m.nd_next = (int*)(d + a_ndn); m.nd_key = (unsigned long long*)(d + a_ndk);
  m.heap.key = (float*)(d + a_qk); m.heap.rec = (int*)(d + a_qr); m.heap.n = 0;
  m.arena.bk = (int*)(d + a_bk); m.arena.tabs = m.tab; m.arena.primes = (const unsigned*)(d + a_primes);
  m.out_mask = (int*)(d + a_mask); m.out_cls = (int*)(d + a_ocls);
  // scalars: stats[8] | bump | base | out_n | status | overflow
  long long* sc = (long long*)(d + a_scalars);
  m.stats = sc; m.arena.collections = sc + 3; m.arena.bump = sc + 8; m.arena.base = sc + 9;
  m.out_n = (int*)(sc + 10); m.status = (int*)(sc + 11); m.arena.overflow = (int*)(sc + 12);
  static const unsigned primes[MN_STL_NPRIMES] = {MN_STL_PRIMES};
  cudaError_t e = cudaMemcpyAsync(p->d_in_class, h_class, (size_t)C * N * 4, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_in_adj, h_adj, (size_t)K * N * 4, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync((void*)m.arena.primes, primes, sizeof(primes), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(sc, 0, 256, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(m.out_mask, 0, N * 4, s);     // a failed run leaves (0, -1), like the drop-in symbol
  if (e == cudaSuccess) e = cudaMemsetAsync(m.out_cls, 0xFF, N * 4, s);
  if (e != cudaSuccess) return done(MN_STATUS_CUDA);
  // edge pass (+ the record init of the hot path, unused here) exactly as the drop-in symbol runs it
  // (like the drop-in symbol, any floats are accepted: out-of-domain maps take the kernel with libm's special values)
  rc = run_front(p, 0, 1, p->d_in_class, p->d_in_adj, clip | MN_INPUT_CHECK_DOMAIN, sdb, omf, mlb, s, false);
  if (rc) return done(rc);
  const MnImage& im = p->h_imgs[0];
  m.clp = im.clp; m.rec_same = im.rec_same; m.rec_diff = im.rec_diff;
  mn_exact_kernel<<<1, 32, 0, s>>>(m);
  int status = 0, n = 0;
  long long st[8] = {0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(st, sc, 64, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&n, m.out_n, 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&status, m.status, 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_mask, m.out_mask, N * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_object_class, m.out_cls, N * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && sdb != 0.0f)  // the reference rewrites the caller's buffer (segment.cc:187-191)
    e = cudaMemcpyAsync(h_adj, p->d_in_adj, (size_t)K * N * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return done(MN_STATUS_CUDA);
  *n_instances = n;
  if (stats4) for (int i = 0; i < 4; i++) stats4[i] = st[i];
  if (status == 1) return done(MN_STATUS_Q_POOL);
  if (status == 2) return done(MN_STATUS_PL_POOL);
  if (status == 3) return done(MN_STATUS_INTERNAL);
  if (status == 4) return done(MN_STATUS_BAD_ARG);
  return done(MN_STATUS_OK);
}

// ------------------------------------------------------------------------------------------------
// test hooks
extern "C" int mn_debug_edge_dump(int H, int W, int C, int K, const int* offset_list, const float* h_class,
                                  float* h_adj, float sdb, float omf, float mlb, float* clp, int* cls,
                                  float* same, float* diff, float* oml, float* mp, int* lo, int* hi) {
  mn_plan* p = nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }  // the caller's current device
  int rc = mn_plan_create(&p, 1, H, W, C, K, offset_list, dev);
  if (rc) return rc;
  auto done = [&](int code) { mn_plan_destroy(p); g_last_error = code; return code; };
  if (ensure_staging(p, 1)) return done(MN_STATUS_CUDA);
  const size_t N = p->N, E = (size_t)p->E;
  cudaStream_t s = p->stream;
  cudaMemcpyAsync(p->d_in_class, h_class, (size_t)C * N * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(p->d_in_adj, h_adj, (size_t)K * N * 4, cudaMemcpyHostToDevice, s);
  rc = run_front(p, 0, 1, p->d_in_class, p->d_in_adj, 0, sdb, omf, mlb, s, false);
  if (rc) return done(rc);
  std::vector<uint4> rec(E);
  std::vector<float> esame(E), ediff(E);
  const MnImage& im = p->h_imgs[0];
  cudaMemcpyAsync(clp, im.clp, N * C * 4, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(cls, im.cls, N * 4, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(rec.data(), im.rec, E * 16, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(esame.data(), im.rec_same, E * 4, cudaMemcpyDeviceToHost, s);  // (edge-pass outputs: no sort has reused the arena)
  cudaMemcpyAsync(ediff.data(), im.rec_diff, E * 4, cudaMemcpyDeviceToHost, s);
  if (sdb != 0.0f) cudaMemcpyAsync(h_adj, p->d_in_adj, (size_t)K * N * 4, cudaMemcpyDeviceToHost, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) return done(MN_STATUS_CUDA);
  for (size_t r = 0; r < E; r++) {
    const uint4 a = rec[r];
    lo[r] = mn_rec_lo(a.x); hi[r] = mn_rec_hi(a.y);
    bool v = lo[r] >= 0;
    if (!v) { lo[r] = -1; hi[r] = -1; }
    oml[r] = v ? mn_u2f(a.z) : 0.f; same[r] = v ? esame[r] : 0.f; diff[r] = v ? ediff[r] : 0.f; mp[r] = v ? mn_u2f(a.w) : 0.f;
  }
  return done(MN_STATUS_OK);
}

// dev hook: time the edge pass alone on B synthetic images (values uniform in the clipped domain)
__global__ void mn_fill_probs_kernel(float* p, size_t n, uint32_t seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    float v = (float)(h >> 8) * (1.0f / 16777216.0f);
    p[i] = fminf(fmaxf(v, 1.1920929e-07f), 0.99999988f);
  }
}
extern "C" int mn_debug_edge_bench(int H, int W, int C, int K, const int* offset_list, int B, int iters, int clip,
                                   float* ms_per_launch) {
  mn_plan* p = nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }  // the caller's current device
  int rc = mn_plan_create(&p, B, H, W, C, K, offset_list, dev);
  if (rc) return rc;
  auto done = [&](int code) { mn_plan_destroy(p); g_last_error = code; return code; };
  const size_t N = p->N;
  float *dc = nullptr, *da = nullptr;
  if (cudaMalloc(&dc, (size_t)B * C * N * 4) != cudaSuccess || cudaMalloc(&da, (size_t)B * K * N * 4) != cudaSuccess) {
    cudaFree(dc);
    return done(MN_STATUS_CUDA);
  }
  cudaStream_t s = p->stream;
  mn_fill_probs_kernel<<<2048, 256, 0, s>>>(dc, (size_t)B * C * N, 1u);
  mn_fill_probs_kernel<<<2048, 256, 0, s>>>(da, (size_t)B * K * N, 2u);
  for (int w = 0; w < 2; w++) rc = launch_edge(p, 0, B, dc, da, clip, 0.0f, s);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, s);
  for (int it = 0; it < iters && !rc; it++) rc = launch_edge(p, 0, B, dc, da, clip, 0.0f, s);
  cudaEventRecord(e1, s);
  cudaError_t e = cudaStreamSynchronize(s);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(dc); cudaFree(da);
  if (ms_per_launch) *ms_per_launch = ms / (float)(iters > 0 ? iters : 1);
  return done(rc ? rc : (e == cudaSuccess ? MN_STATUS_OK : MN_STATUS_CUDA));
}

extern "C" int mn_debug_libm(int which, unsigned first_bits, unsigned n, float bias, float* h_out) {
  g_last_error = MN_STATUS_OK;
  if (n == 0 || !h_out) return MN_STATUS_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  float* d = nullptr;
  MN_CUDA_OK(cudaMalloc(&d, (size_t)n * 4));
  mn_libm_kernel<<<1184, 256>>>(which, first_bits, n, bias, d);
  cudaError_t e = cudaMemcpy(h_out, d, (size_t)n * 4, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) { g_last_error = MN_STATUS_CUDA; return MN_STATUS_CUDA; }
  return MN_STATUS_OK;
}
