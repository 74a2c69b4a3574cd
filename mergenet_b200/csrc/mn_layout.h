// mn_layout.h -- per-image device workspace of the merge segmenter and the small data-structure
// primitives every kernel shares (packed object word, (lo,hi)->record hash, SPMD/atomic shims).
//
// HBM layout per image (N = H*W pixels, E = N*K record slots, slot r = pixel*K + k):
//   objects   clp[N*C] f32 | obj[N] uint4 (npix | cls<<24, sameness sum, pixel-array offset, -) |
//             parent[N] | live_mask[N] | pix_pool: one contiguous pixel array per multi-pixel
//             object, capacity = next power of two >= npix (>= 4)          (Object, h:85-137)
//   records   rec[E] 16 B: (lo | hash slot << 24 | guard state << 29, hi, oml, mp); lo field all ones = dead
//             (AdjacencyRecord, h:175-232; the sameness / differentness sums of h:131,186-187 feed only the printed
//             total log-prob, which mn_logprob_scratch_kernel evaluates from the final partition instead)
//   hash      2-choice, 8-slot buckets of u32 (fingerprint<<26 | rec+1): (lo,hi) -> record
//             (replaces the per-object unordered_map lookups of cc:685-688)
//   queue     init_keys[E] u64, sorted: the initial priority-queue entries (cc:225-227);
//             a radix tree of unsorted entry chunks for entries created later (cc:564,697,705);
//             the entry pool aliases rec_same/rec_diff, which are dead after record init.
#pragma once
#include <stdint.h>

#include "mn_common.h"

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#else
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r = {x, y, z, w}; return r; }
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }
static inline int2 make_int2(int x, int y) { int2 r = {x, y}; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r = {x, y}; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
#endif

// ---- SPMD shims: the scheduler is written as phases of `for (i = tid; i < n; i += nt)` loops
// separated by block barriers, with all cross-thread traffic through shared/global arrays and
// atomics.  Compiled for the host (tests/emul) it runs with one logical thread. -----------------
#if defined(__CUDA_ARCH__)
#define MN_SYNC() __syncthreads()
#define MN_TID ((int)threadIdx.x)
#define MN_NT ((int)blockDim.x)
#define MN_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define MN_ATOMIC_SUB(p, v) atomicSub((p), (v))
#define MN_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define MN_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define MN_ATOMIC_OR(p, v) atomicOr((p), (v))
#define MN_ATOMIC_AND(p, v) atomicAnd((p), (v))
#define MN_ATOMIC_CAS(p, c, v) atomicCAS((p), (c), (v))
#else
#define MN_SYNC() ((void)0)
#define MN_TID 0
#define MN_NT 1
template <typename T> static inline T mn_h_add(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T mn_h_sub(T* p, T v) { T o = *p; *p = o - v; return o; }
template <typename T> static inline T mn_h_min(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T mn_h_max(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T mn_h_or(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T mn_h_and(T* p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T mn_h_cas(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
#define MN_ATOMIC_ADD(p, v) mn_h_add((p), (v))
#define MN_ATOMIC_SUB(p, v) mn_h_sub((p), (v))
#define MN_ATOMIC_MIN(p, v) mn_h_min((p), (v))
#define MN_ATOMIC_MAX(p, v) mn_h_max((p), (v))
#define MN_ATOMIC_OR(p, v) mn_h_or((p), (v))
#define MN_ATOMIC_AND(p, v) mn_h_and((p), (v))
#define MN_ATOMIC_CAS(p, c, v) mn_h_cas((p), (c), (v))
#endif

#define MN_ORD_BITS 28       // initial-entry tie-break ordinal: tie u << 4 | rank of the offset distance (K <= 16)
#define MN_QCH 64            // queue entries per tree chunk (16 B each)
#define MN_HASH_FP_SHIFT 26  // slot = fingerprint(6) << 26 | (rec + 1)
#define MN_TREE_BITS 3       // digit width of the queue tree below a root
#define MN_TREE_FANOUT (1 << MN_TREE_BITS)  // children per split: small, so that the children of a split
                                            // leaf (> MN_LEAFCAP entries) are still worth a load each

// queue tree roots: one per 2^15 float-bit patterns over [MN_ROOT_LO, MN_ROOT_HI)
#define MN_ROOT_SHIFT 15
#define MN_ROOT_LO_BITS 0x39800000u  // 2^-12
#define MN_ROOT_HI_BITS 0x43800000u  // 2^8
#define MN_NROOTS (((MN_ROOT_HI_BITS - MN_ROOT_LO_BITS) >> MN_ROOT_SHIFT) + 2)

enum MnStatus {
  MN_OK = 0,
  MN_ERR_BAD_ARG = 1,
  MN_ERR_PL_POOL = 2,     // pixel-array pool exhausted
  MN_ERR_Q_POOL = 3,      // queue entry pool exhausted
  MN_ERR_TREE_POOL = 4,   // queue tree node pool exhausted
  MN_ERR_HASH_FULL = 5,   // both hash buckets full and overflow area full
  MN_ERR_INTERNAL = 6,    // an invariant failed (the reference would exit(1): cc:42,667,672)
  MN_ERR_CUDA = 7,
  MN_ERR_LIMIT = 8        // iteration guard tripped
};

// per-image control block (global memory)
struct MnCtl {
  int status;
  int n_init;          // non-sentinel entries in init_keys
  int static_cursor;
  int pix_bump;
  int qc_bump, qc_free_top;
  int tn_bump;
  int hash_ovf_n;
  int tree_entries;    // entries currently stored in the tree
  int peak_entries, peak_chunks;  // high-water marks of the tree
  int n_instances;     // output: labels 1..n
  int fail_line;       // source line of the first failure (diagnostics)
  // statistics (north star: per-round latency, round count, merges/s)
  long long rounds, events, merges, restores, invalid_pops, solo_events;
  long long refills, flushes, splits, pairs, cuts_conflict, cuts_cascade, cuts_capacity;
  long long cycles_total;
  long long requeues;
  long long pix_gcs;   // collections of the pixel-array pool
  long long cyc[16];   // cycle buckets (MN_CY_*)
};

struct MnImage {
  // objects
  float* clp;
  int* cls;  // argmax class per pixel from the edge pass (consumed by record init)
  uint4* obj;  // x = npix | cls << 24, y = bits of the object's sameness sum (h:131), z = pixel array offset
               // (-1: single pixel); w = LIVE MASK OF PIXEL p (bit k: record slot p*K+k alive, bit 16+k: the
               // slot of the record arriving through offset k) -- per pixel, whatever object owns it
  int* parent;
  int* pix_pool;
  // records
  uint4* rec;       // one 16-byte record per slot r: x = lo (24 bits) | hash position (5 bits) << 24 | guard state
                    // (2 bits) << 29, y = hi, z = bits of oml, w = bits of mp; lo field all ones: dead   (h:175-232)
  float* rec_same;  // edge-pass outputs; dead after record init, then aliased by q_ent
  float* rec_diff;
  // hash
  uint32_t* hash;
  uint32_t* hash_ovf;  // small linear overflow area of rec+1 values
  uint32_t hash_nbuckets;
  uint32_t hash_ovf_cap;
  // queue
  uint64_t* init_keys;
  uint4* q_ent;  // [qc_cap * MN_QCH] (mp bits, rec, lo, hi): validated against the record on load.  One arena with
                 // init_keys: chunk c < qc_low_n overlays init_keys[2 * MN_QCH * c ...), usable once consumed
  int* qc_next;
  int* qc_free;
  int4* tn;     // tree nodes: (first chunk, last chunk, entries below, first child)
  int* tn_dir;  // first 8 chunk ids of every leaf
  int pix_cap, qc_cap, tn_cap;
  int qc_low_n;  // chunks covered by the initial-key array (arena prefix)
  // outputs
  int* out_mask;
  int* out_cls;
  MnCtl* ctl;
};

// ---- capacities of the per-image pools (one definition for the library, mn_api.cu: ws_layout, and for the host build
// of the scheduler that the CPU suite and the long seeded sweep run, tests/emul: a pool that is too small must show up
// there, not on a user's image) ---------------------------------------------------------------------
struct MnCaps {
  int pix_cap, qc_low_n, qc_cap, tn_cap;
  uint32_t hash_nbuckets, hash_ovf_cap;
};
MN_HD MnCaps mn_workspace_caps(size_t N, size_t E) {
  MnCaps c;
  // Pixel pool, two halves.  After a collection the live arrays take at most 2 N ints (capacity = pow2 >= npix < 2 npix),
  // and the new survivor arrays of ONE round at most 2 N more (its merges touch disjoint objects): 4 N ints per half
  // always fit.  (3 N did not: two objects of just over 2^k pixels each -- capacities 2^(k+1) -- merging into one of
  // capacity 2^(k+2) need 4 x their pixel count; found by tests/manual/soak_sweep.py on a 59 x 75 image that collapses
  // to one object, regression tests in tests/test_emul_scheduler.py and tests/test_gpu_parity.py)
  c.pix_cap = (int)(2 * (4 * N + 2048));
  // One arena holds, in turn, the edge pass outputs rec_same | rec_diff (8 E bytes, dead after record init), the sorted
  // initial keys (8 E bytes, written by the sort), and the queue chunks: the consumed prefix of the keys is recycled as
  // chunks (qc_low_n of them), and qc_cap - qc_low_n extra chunks cover the early demand (measured: 0.30 E entries at
  // 256x512).  qc_low_n is rounded UP: the first chunk of the scheduler's own starts behind the LAST key.
  // (Rounded down -- until the end of round 2 -- it overlaid the final E % 128 keys, which are sentinels of dormant /
  // out-of-image slots on every usual shape (and E % 128 = 0 on all BASELINE shapes), but real entries when nearly every
  // record starts with a priority >= 0 and few slots leave the image: one offset (0, 1) with a large
  // merge_logprob_bias; found by tests/manual/soak_sweep.py, same regression tests)
  c.qc_low_n = (int)((E * 8 + (size_t)MN_QCH * 16 - 1) / ((size_t)MN_QCH * 16));
  c.qc_cap = c.qc_low_n + (int)(E * 9 / 20 / MN_QCH + 4 * MN_NROOTS + 4096);
  const size_t splits = E / 512 > 4096 ? E / 512 : 4096;  // measured: < E / 1700 splits
  c.tn_cap = MN_NROOTS + MN_TREE_FANOUT * (int)splits;
  c.hash_nbuckets = (uint32_t)(E * 18 / 10 / 8 + 64);
  c.hash_ovf_cap = 16384;
  return c;
}

// ---- the 16-byte record --------------------------------------------------------------------------
// Guard state of a record = what is queued for it (the queue is lazy, see mn_merge.cuh):
//   NONE   no entry is expected (fresh, or its guard was consumed);
//   EXACT  an entry at exactly the stored priority (and key) is queued;
//   ABOVE  some entry with a priority strictly above the stored one is queued (the stored priority was lowered, or
//          went negative, after that entry was pushed): the first such entry to surface re-queues the record.
#define MN_G_NONE 0u
#define MN_G_EXACT 1u
#define MN_G_ABOVE 2u
#define MN_HS_OVF 16u    // hash position: 0..7 slot of bucket b1, 8..15 slot of bucket b2, 16 the overflow area
#define MN_HS_NONE 31u
#define MN_REC_DEAD 0xFFFFFFFFu
#define MN_REC(im, r) ((im).rec[(size_t)(r)])
MN_HD uint32_t mn_rec_pack_x(int lo, uint32_t hs, uint32_t g) { return (uint32_t)lo | (hs << 24) | (g << 29); }
MN_HD int mn_rec_lo(uint32_t x) { const uint32_t l = x & 0xFFFFFFu; return l == 0xFFFFFFu ? -1 : (int)l; }
MN_HD int mn_rec_hi(uint32_t y) { return (int)(y & 0xFFFFFFu); }
MN_HD uint32_t mn_rec_hs(uint32_t x) { return (x >> 24) & 31u; }
MN_HD uint32_t mn_rec_guard(uint32_t x) { return (x >> 29) & 3u; }
MN_HD uint32_t mn_rec_with_guard(uint32_t x, uint32_t g) { return (x & ~(3u << 29)) | (g << 29); }
MN_HD uint4 mn_load_rec(const MnImage& im, int r) { return im.rec[(size_t)r]; }
MN_HD void mn_store_rec(const MnImage& im, int r, uint4 v) { im.rec[(size_t)r] = v; }
MN_HD int2 mn_rec_key(const MnImage& im, int r) {  // (lo, hi); lo = -1: dead
  const uint2 k = *reinterpret_cast<const uint2*>(&im.rec[(size_t)r]);
  return make_int2(mn_rec_lo(k.x), mn_rec_hi(k.y));
}
// one 8-slot hash bucket (32 bytes)
MN_HD void mn_load_bucket(const MnImage& im, uint32_t b, uint32_t* out) {
#if defined(__CUDA_ARCH__)
  asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(out[0]), "=r"(out[1]), "=r"(out[2]), "=r"(out[3]), "=r"(out[4]), "=r"(out[5]), "=r"(out[6]), "=r"(out[7])
               : "l"(im.hash + (size_t)b * 8));
#else
  const uint4* p = reinterpret_cast<const uint4*>(im.hash + (size_t)b * 8);
  uint4 x = p[0], y = p[1];
  out[0] = x.x; out[1] = x.y; out[2] = x.z; out[3] = x.w;
  out[4] = y.x; out[5] = y.y; out[6] = y.z; out[7] = y.w;
#endif
}
MN_HD uint32_t mn_pack_nc(int npix, int cls) { return (uint32_t)npix | ((uint32_t)cls << 24); }
MN_HD int mn_nc_npix(uint32_t nc) { return (int)(nc & 0xFFFFFFu); }
MN_HD int mn_nc_cls(uint32_t nc) { return (int)(nc >> 24); }
// capacity of the pixel array of an object of n pixels (0: the single pixel is the object id itself)
MN_HD int mn_pix_cap(int n) {
  if (n <= 1) return 0;
  int c = 4;
  while (c < n) c <<= 1;
  return c;
}

// ---- (lo,hi) -> record hash ---------------------------------------------------------------------
MN_HD uint64_t mn_mix64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}
struct MnHashPos {
  uint32_t b1, b2, fp;
};
MN_HD MnHashPos mn_hash_pos(uint32_t nbuckets, int lo, int hi) {
  uint64_t h = mn_mix64(((uint64_t)(uint32_t)lo << 32) | (uint32_t)hi);
  MnHashPos p;
  p.b1 = (uint32_t)(((h & 0xffffffffu) * (uint64_t)nbuckets) >> 32);
  uint32_t d = (uint32_t)((((h >> 32) & 0x3ffffffu) * (uint64_t)(nbuckets - 1)) >> 26);
  p.b2 = p.b1 + 1 + d;
  if (p.b2 >= nbuckets) p.b2 -= nbuckets;
  p.fp = (uint32_t)(h >> 58);
  return p;
}
// 5-bit hash position kept in the record <-> global slot index (-1: the overflow area)
MN_HD uint32_t mn_hs_of_slot(const MnHashPos& p, int slot) {
  if (slot < 0) return MN_HS_OVF;
  return ((uint32_t)slot >> 3) == p.b1 ? ((uint32_t)slot & 7u) : (8u + ((uint32_t)slot & 7u));
}
MN_HD int mn_slot_of_hs(const MnHashPos& p, uint32_t hs) {
  if (hs >= MN_HS_OVF) return -1;
  return (int)((hs < 8u ? p.b1 : p.b2) * 8u + (hs & 7u));
}
// returns record id or -1
MN_HD int mn_hash_find(const MnImage& im, int lo, int hi) {
  MnHashPos p = mn_hash_pos(im.hash_nbuckets, lo, hi);
  for (int w = 0; w < 2; w++) {
    const uint32_t* bk = im.hash + (size_t)(w ? p.b2 : p.b1) * 8;
    for (int s = 0; s < 8; s++) {
      uint32_t v = bk[s];
      if (v != 0 && (v >> MN_HASH_FP_SHIFT) == p.fp) {
        int r = (int)(v & ((1u << MN_HASH_FP_SHIFT) - 1)) - 1;
        int2 lh = mn_rec_key(im, r);
        if (lh.x == lo && lh.y == hi) return r;
      }
    }
  }
  int n = im.ctl->hash_ovf_n;
  for (int i = 0; i < n; i++) {
    uint32_t v = im.hash_ovf[i];
    if (v != 0) {
      int r = (int)v - 1;
      int2 lh = mn_rec_key(im, r);
      if (lh.x == lo && lh.y == hi) return r;
    }
  }
  return -1;
}
// Safe against concurrent inserts/erases of other keys.  Two-choice placement: the emptier of the two
// candidate buckets first.  Returns the global slot index, or -1 when the record went to the overflow
// area (the caller stores it in the record: erasing then needs no lookup).
MN_HD int mn_hash_insert(const MnImage& im, int lo, int hi, int rec) {
  MnHashPos p = mn_hash_pos(im.hash_nbuckets, lo, hi);
  uint32_t val = (p.fp << MN_HASH_FP_SHIFT) | (uint32_t)(rec + 1);
  uint32_t* bk1 = im.hash + (size_t)p.b1 * 8;
  uint32_t* bk2 = im.hash + (size_t)p.b2 * 8;
  int f1 = 0, f2 = 0;
  for (int s = 0; s < 8; s++) { f1 += bk1[s] == 0; f2 += bk2[s] == 0; }
  for (int w = 0; w < 2; w++) {
    uint32_t* bk = ((w == 0) == (f1 >= f2)) ? bk1 : bk2;
    for (int s = 0; s < 8; s++) {
      if (bk[s] == 0 && MN_ATOMIC_CAS(&bk[s], 0u, val) == 0u) return (int)(bk - im.hash) + s;
    }
  }
  int i = MN_ATOMIC_ADD(&im.ctl->hash_ovf_n, 1);
  if ((uint32_t)i < im.hash_ovf_cap) {
    im.hash_ovf[i] = (uint32_t)(rec + 1);
  } else {
    im.ctl->status = MN_ERR_HASH_FULL;
  }
  return -1;
}
MN_HD void mn_hash_erase(const MnImage& im, int lo, int hi, int rec) {
  MnHashPos p = mn_hash_pos(im.hash_nbuckets, lo, hi);
  uint32_t val = (p.fp << MN_HASH_FP_SHIFT) | (uint32_t)(rec + 1);
  for (int w = 0; w < 2; w++) {
    uint32_t* bk = im.hash + (size_t)(w ? p.b2 : p.b1) * 8;
    for (int s = 0; s < 8; s++) {
      if (bk[s] == val) {
        bk[s] = 0;
        return;
      }
    }
  }
  int n = im.ctl->hash_ovf_n;
  for (int i = 0; i < n && (uint32_t)i < im.hash_ovf_cap; i++) {
    if (im.hash_ovf[i] == (uint32_t)(rec + 1)) {
      im.hash_ovf[i] = 0;
      return;
    }
  }
}
