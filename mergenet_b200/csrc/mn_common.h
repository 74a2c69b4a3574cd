// mn_common.h -- types and arithmetic shared by every kernel of the merge-segmenter path.
//
// Everything here restates arithmetic of the reference so that results are bit-identical:
//   /root/reference/utils/csegment/segment.cc  (cc:LINE)   /root/reference/utils/csegment/segment.h (h:LINE)
// All fp32 operations are individually rounded (no FMA contraction): the reference is compiled for
// baseline x86-64 where every * + / is its own instruction.  The file compiles both as
// CUDA device code and as plain host C++ (tests/emul builds the scheduler for the host to unit-test
// its logic; that build is test infrastructure and is never reachable from the product API).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MN_HD __host__ __device__ __forceinline__
#define MN_D __device__ __forceinline__
#else
#define MN_HD inline
#define MN_D inline
#endif

// ---- fp32 ops that must never be contracted into FMAs --------------------------------------
#if defined(__CUDA_ARCH__)
#define MN_FADD(a, b) __fadd_rn((a), (b))
#define MN_FSUB(a, b) __fsub_rn((a), (b))
#define MN_FMUL(a, b) __fmul_rn((a), (b))
#define MN_FDIV(a, b) __fdiv_rn((a), (b))
#define MN_DADD(a, b) __dadd_rn((a), (b))
#define MN_DMUL(a, b) __dmul_rn((a), (b))
#else
// host build is compiled with -ffp-contract=off
#define MN_FADD(a, b) ((float)((float)(a) + (float)(b)))
#define MN_FSUB(a, b) ((float)((float)(a) - (float)(b)))
#define MN_FMUL(a, b) ((float)((float)(a) * (float)(b)))
#define MN_FDIV(a, b) ((float)((float)(a) / (float)(b)))
#define MN_DADD(a, b) ((double)(a) + (double)(b))
#define MN_DMUL(a, b) ((double)(a) * (double)(b))
#endif

#define MN_MAX_K 16   // offsets per image (live masks hold 2*K bits per pixel)
#define MN_MAX_C 256  // classes (class id packed in 8 bits)

MN_HD uint32_t mn_f2u(float f) {
  uint32_t u;
#if defined(__CUDA_ARCH__)
  u = __float_as_uint(f);
#else
  memcpy(&u, &f, 4);
#endif
  return u;
}
MN_HD float mn_u2f(uint32_t u) {
  float f;
#if defined(__CUDA_ARCH__)
  f = __uint_as_float(u);
#else
  memcpy(&f, &u, 4);
#endif
  return f;
}

// ---- queue order -----------------------------------------------------------------------------
// A queue entry pops before another iff (mp desc, tie(lo, hi) asc): a deterministic tie-break for
// what the reference leaves to libstdc++ heap layout and unordered_map iteration order (h:270-275).
// tie(lo, hi) = (u, D) compared lexicographically, D = hi - lo, u = (bitrev24(lo) + 0x9E3779 * D) mod
// 2^24.  (u, D) <-> (lo, hi) is a bijection, so the order is total; it scatters equal priorities
// (oracle-mode maps have millions) over the image AND over the records of one object, the way the
// reference's own hash-driven order does: consecutive pops rarely touch the same or neighbouring
// objects (few conflicts per round) and objects grow compactly.  For an initial record D is one of the
// K offset distances, so (u, rank of D) fits the 28-bit ordinal of the sorted initial keys.  Any
// fixed rule is equally valid against the reference; this one is fast on a GPU.
#define MN_TIE_MUL 0x9E3779u
MN_HD uint32_t mn_brev24(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __brev(v) >> 8;
#else
  uint32_t r = 0;
  for (int i = 0; i < 24; i++) { r = (r << 1) | (v & 1u); v >>= 1; }
  return r;
#endif
}
MN_HD uint32_t mn_tie_u(int lo, int hi) {
  return (mn_brev24((uint32_t)lo) + MN_TIE_MUL * (uint32_t)(hi - lo)) & 0xFFFFFFu;
}
MN_HD uint64_t mn_tie(int lo, int hi) {
  return ((uint64_t)mn_tie_u(lo, hi) << 24) | (uint64_t)(uint32_t)(hi - lo);
}
struct MnEnt {
  float mp;
  int lo, hi, rec;
};
MN_HD bool mn_before(float amp, int alo, int ahi, float bmp, int blo, int bhi) {
  if (amp != bmp) return amp > bmp;
  return mn_tie(alo, ahi) < mn_tie(blo, bhi);
}
MN_HD bool mn_ent_before(const MnEnt& a, const MnEnt& b) {
  return mn_before(a.mp, a.lo, a.hi, b.mp, b.lo, b.hi);
}

// ---- merge priority (cc:107-150) --------------------------------------------------------------
// c1/cls1/n1 belong to the LOWER-id object of the record (cc:49-56), c2/... to the higher.
// clp pointers may be null when the classes are equal.  Returns the priority; *merged = class the
// merged object would take (cc:121,136-138: first maximum of the joint vector).
MN_HD float mn_priority(float oml, float omf, float mlb, int C, int n1, int cls1, const float* c1,
                        int n2, int cls2, const float* c2, int* merged) {
  float cdl;
  int m;
  if (cls1 == cls2) {  // cc:108,120-121
    cdl = 0.0f;
    m = cls1;
  } else {  // cc:123-140
    float best = MN_FADD(c1[0], c2[0]);
    m = 0;
    for (int c = 1; c < C; c++) {
      float j = MN_FADD(c1[c], c2[c]);
      if (j > best) {
        best = j;
        m = c;
      }
    }
    cdl = MN_FSUB(best, c1[cls1]);
    cdl = MN_FSUB(cdl, c2[cls2]);
  }
  if (merged) *merged = m;
  float den = (float)(n1 + n2);                   // cc:147 (size_t -> float, exact below 2^24)
  float num = MN_FADD(MN_FMUL(oml, omf), cdl);    // cc:148
  return MN_FADD(MN_FDIV(num, den), mlb);         // cc:149
}

// ---- glibc-exact logf on the clipped domain [2^-23, 1-2^-23] ------------------------------------
// Public glibc / ARM optimized-routines algorithm, evaluated in fp64 with every
// product and sum individually rounded.  Pinned to the host libm exhaustively by
// tests/test_libm_parity.py (CPU: oracle restatement; GPU: this function on the device).
#define MN_LOGF_TABLE                                                                             \
  {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},   \
  {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},   \
  {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},      \
  {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},   \
  {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},                                \
  {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},     \
  {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},     \
  {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}

struct MnLogfTab {
  double invc, logc;
};

// tab: 16 entries of {invc, logc} (shared memory on the device).
// Inputs outside the positive normal range take glibc's own special cases (logf(0) = -inf, negative / NaN ->
// NaN, +inf -> +inf, subnormals rescaled by 2^23): a caller of the raw C ABI may hand in unclipped maps, and
// the same_different_bias transform can round a probability to exactly 1 or 0 (cc:183-195).
MN_HD float mn_logf_exact(float x, const MnLogfTab* tab) {
  uint32_t ix = mn_f2u(x);
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2u == 0u) return mn_u2f(0xff800000u);                       // log(+-0) = -inf
    if (ix == 0x7f800000u) return x;                                     // log(inf) = inf
    if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return mn_u2f(0x7fc00000u);  // negative or NaN
    ix = mn_f2u(x * 8388608.0f) - (23u << 23);                           // subnormal: normalise
  }
  uint32_t tmp = ix - 0x3f330000u;
  int i = (tmp >> 19) & 15;
  int k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & 0xff800000u);
  double z = (double)mn_u2f(iz);
  double r = MN_DADD(MN_DMUL(z, tab[i].invc), -1.0);
  double y0 = MN_DADD(tab[i].logc, MN_DMUL((double)k, 0x1.62e42fefa39efp-1));
  double r2 = MN_DMUL(r, r);
  double y = MN_DADD(MN_DMUL(0x1.5575b0be00b6ap-2, r), -0x1.ffffef20a4123p-2);
  y = MN_DADD(MN_DMUL(-0x1.00ea348b88334p-2, r2), y);
  y = MN_DADD(MN_DMUL(y, r2), MN_DADD(y0, r));
  return (float)y;
}
