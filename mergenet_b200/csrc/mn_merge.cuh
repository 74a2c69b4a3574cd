// mn_merge.cuh -- order-exact merge scheduler: one persistent CTA per image (sm_100a).
//
// Reproduces the reference's RunSegmentation / Merge loop (cc:539-573, cc:602-727) exactly, up to
// the tie order among equal priorities (deterministic here: mp desc, then the (u, D) scatter rule of mn_common.h).
//
// The reference's lazy heap is observationally an indexed map  record -> stored priority  holding
// the records with stored mp >= 0 (cc:554-565).  A ROUND takes the next MN_H queue entries in pop
// order, PLANS each against the round-start state (read only), then COMMITS the longest prefix
// that the sequential heap would provably execute in exactly this order with exactly these results:
//   (a) no member reads or writes an object written by an earlier member, and writes none an
//       earlier member read (objects written by a merge: both endpoints; read: every neighbour of
//       the absorbed object, whose records are rewired -- cc:650-707; a non-merging pop reads its
//       two endpoints and re-stores only its own priority -- cc:560-565);
//   (b) no earlier member stores a priority that would pop before a later member.
// Rule (a)+(b) was validated against the sequential restatement on the host
// (oracle/mergenet_oracle.c: mno_run_rounds_model).  Member 0 always commits, so every round makes
// progress.  All state of an image is private to its CTA: no inter-CTA communication.
//
// A round is organised around the number of DEPENDENT global-memory round trips, not around
// instruction count: every load a phase needs is issued by a different thread in the same phase
// (stage: record + both objects + class vectors + hash buckets of a candidate at once; pixels ->
// live masks; pair: record -> {neighbour object, four hash buckets} -> partner record).
//
// Queue.  `hot` (shared memory, sorted) holds every entry that pops before-or-at `bound`; colder
// entries live in HBM: the sorted initial entries (init_keys, cursor) and, for entries created
// later, a lazily split radix tree of unsorted chunks keyed by (mp bits, lo, hi).  The queue is lazy
// in two ways.  Like the reference's heap, an entry that no longer describes its record is dropped
// when it surfaces (cc:554-559).  Unlike it, a record whose priority is re-stored LATER in pop order
// than an entry it already has gets no new entry: two bits of the record (its GUARD STATE, mn_layout.h)
// remember whether an entry at exactly the stored priority is queued (EXACT) or only one strictly
// above it (ABOVE, the "guard"); when a guard surfaces and the stored priority is lower, an exact
// entry is queued then ("requeue").  A priority is stored with a new entry only when it RISES (or when
// the record has no queued entry, or is re-keyed without falling): every queued entry of a record was
// pushed at a priority the record had, so the one pushed at its running maximum since the last push
// is still queued whenever the state is not NONE.  Invariant: every live record with stored
// mp >= 0 has a queued entry popping before-or-at its true position, so no pop can be missed, and
// the sequence of EXACT pops -- the only ones with side effects -- is the reference's.
//
// The file is written as SPMD phases (see mn_layout.h) and also compiles for the host, where
// tests/emul runs it single-threaded to unit-test the logic; that build is test infrastructure.
#pragma once
#include <limits.h>

#include "mn_common.h"
#include "mn_layout.h"

#ifndef MN_H
#define MN_H 32          // candidates per round (<= 64)
#endif
#define MN_PW 1024       // pixel work-list capacity
#define MN_CW 64         // queue-chunk work-list capacity
#define MN_WL 960        // (candidate, record) pair work-list capacity
#define MN_HC 1024       // hot capacity
#define MN_NE 1024       // new hot-bound entries per round (>= MN_WL + MN_H, <= MN_SB)
#ifndef MN_IC
#define MN_IC 2048       // insert-buffer capacity
#endif
#define MN_SB 1024       // sort buffer capacity
#ifndef MN_LEAFCAP
#define MN_LEAFCAP 512   // tree leaves larger than this are split before they are loaded
#endif
#define MN_CT 2048       // conflict-table slots (power of two)
#define MN_GCL 512        // objects scanned per garbage-collection step (one per thread)
#ifndef MN_LF
#define MN_LF 128         // leaves one refill may load
#endif
#define MN_OVF 128        // records in the hash overflow area (cached in shared memory)
#ifndef MN_REFILL_TARGET
#define MN_REFILL_TARGET 256      // stop loading tree leaves once this many entries are staged (A/B at 1024x2048, merge ms: 64: 11212, 128: 10986, 192: 10950, 256: 10925-10983, 384: 11128, 640: 11618, 832: 12042)
#endif
#ifndef MN_REFILL_STATIC_MIN
#define MN_REFILL_STATIC_MIN 128  // sort-buffer slots always left for initial entries
#endif
#define MN_NEG_INF (-3.0e38f)
#ifndef MN_RANK_MAX
#define MN_RANK_MAX 256  // up to this many entries are ordered by brute-force ranking (no barriers)
#endif
// cycle accounting buckets (thread 0, clock64)
#define MN_NCYC 16
#define MN_CY_CONFLICT 0 // footprints into the conflict table + conflict check
#define MN_CY_PLAN 1
#define MN_CY_ACCEPT 2
#define MN_CY_COMMIT 3
#define MN_CY_HOT 4
#define MN_CY_FLUSH 5
#define MN_CY_REFILL 6
#define MN_CY_SPLIT 7
#define MN_CY_SOLO 8
#define MN_CY_PAIRLIST 9 // capacity cut by pairs + pair lists
#define MN_CY_RF_LEAVES 10  // refill: tree leaves
#define MN_CY_RF_INIT 11    // refill: initial entries, bound, bulk requeue
#define MN_CY_RF_SORT 12    // refill: ordering + merge into hot
#define MN_CY_SEL_STAGE 13  // select: staging loads
#define MN_CY_SEL_CLASS 14  // select: classify, merged vectors, pixel capacity
#define MN_CY_SEL_PIX 15    // select: pixels, masks, pair lists

// candidate kinds
#define MN_K_DROP 0      // stale entry: dropped when consumed
#define MN_K_RESTORE 1   // exact pop whose recomputed priority differs: re-store (cc:563-565)
#define MN_K_MERGE 2     // exact pop that merges (cc:561-562)
#define MN_K_REQUEUE 3   // guard entry of a record whose stored priority is lower: queue the exact entry
#define MN_K_UNGUARD 4   // guard entry of a dormant record (stored mp < 0): forget the guard

struct MnOffsets {
  int K;
  int delta[MN_MAX_K];      // linear pixel delta of offset k: dr*W + dc
  int k_of_rank[MN_MAX_K];  // offset index with the rank-th smallest |delta|
};

struct MnSm {
  // queue
  float hot_mp[2][MN_HC]; int hot_lo[2][MN_HC]; int hot_hi[2][MN_HC]; int hot_rec[2][MN_HC];
  int hsel;  // which hot buffer is current
  float ins_mp[MN_IC]; int ins_lo[MN_IC]; int ins_hi[MN_IC]; int ins_rec[MN_IC];
  float ne_mp[MN_NE]; int ne_lo[MN_NE]; int ne_hi[MN_NE]; int ne_rec[MN_NE]; int ne_pos[MN_NE];
  float sb_mp[MN_SB]; int sb_lo[MN_SB]; int sb_hi[MN_SB]; int sb_rec[MN_SB];
  int sb_node[MN_SB];
  uint32_t root_bits[(MN_NROOTS + 31) / 32];
  uint32_t root_sum[((MN_NROOTS + 31) / 32 + 31) / 32];
  int cw_chunk[MN_CW];
  int lf_start[MN_LF + 1]; int lf_cnt[MN_LF]; int lf_base[MN_LF]; int nlf;  // leaves of one refill: sb ranges
  int red_idx[1024];  // argmax reduction scratch
  int4 scan_nd[MN_TREE_FANOUT]; int4 leaf_nd; int path_cnt[32];
  int qc_free_top, qc_bump, tn_bump, tree_entries, static_cursor, n_init;  // mirrors of MnCtl
  int qc_low_bump, qc_low_avail;  // chunks recycled from the consumed prefix of the initial keys
  int peak_entries, peak_chunks;
  // candidates: staged loads
  int c_rec[MN_H]; float c_key[MN_H]; int c_lo[MN_H]; int c_hi[MN_H]; int c_kind[MN_H];
  uint4 c_recw[MN_H]; int2 c_lh[MN_H]; uint4 c_obj[MN_H][2];  // the staged record (raw words), its key, both objects
  int c_dup[MN_H];
  // candidates: classification
  float c_newmp[MN_H]; int c_merged[MN_H]; int c_surv[MN_H]; int c_abs[MN_H]; int c_na[MN_H]; int c_nb[MN_H];
  int c_ptra[MN_H]; int c_ptrb[MN_H]; int c_newptr[MN_H]; int c_cpbase[MN_H];
  int c_eslot[MN_H];
  int c_pwbase[MN_H]; int c_npairs[MN_H]; int c_pbase[MN_H]; int c_pfill[MN_H];
  unsigned long long c_maxnew[MN_H];  // pop-order key (mn_pop_key) of the earliest-popping entry the candidate stores; 0: none
  int c_conflict[MN_H]; int c_accept[MN_H];
  int m_list[MN_H]; int m_base[MN_H + 1]; int nm;      // merging candidates and their pixel ranges
  int cp_list[MN_H]; int cp_base[MN_H + 1]; int ncp;   // accepted merges whose survivor array moves
  // work lists
  int pw_pix[MN_PW]; uint32_t pw_mask[MN_PW]; int pw_cand[MN_PW]; int pw_cnt[MN_PW]; int pw_off[MN_PW];
  // the (candidate, record) pair plan of a round and the scratch of distribute() are never live together
  union {
    struct {
      int cand[MN_WL]; int t[MN_WL]; int x[MN_WL]; int u[MN_WL];
      float oml[MN_WL]; int g[MN_WL]; float mpold[MN_WL];  // g / mpold: guard state and stored priority of the record that takes the sum
      float mp[MN_WL];   // (refill parks its sorted leaf run in mp / lo / hi / eslot: keep them beyond w.ds)
      int lo[MN_WL]; int hi[MN_WL]; int eslot[MN_WL]; int islot[MN_WL];
    } pr;
    // distribute(): per entry (group << 16 | index in group); per group node / count / old tail /
    // (old fill | directory base << 8); directory of freshly allocated chunks
    struct {
      int el[MN_SB]; int node[MN_SB]; int cnt[MN_SB]; int tail[MN_SB]; int fd[MN_SB];
      int dir[MN_SB + 128];
    } ds;
  } w;
  // conflict table: object -> (min writer candidate, min reader candidate)
  int ct_obj[MN_CT]; int ct_w[MN_CT]; int ct_r[MN_CT];
  // scalars
  int nhot, nins, nne, ncw, npw, npr, ncand, nacc, cutpos, solo, first, tmp0, tmp1, tmp2, tmp3;
  int ds_ngroups, ds_ndir;  // distribute() counters
  int pix_bump;             // pixel-array pool bump pointer (mirrors MnCtl)
  int pix_hi;               // end of the active half of the pool (semi-space: the live arrays never exceed 2 N ints)
  int need_gc, gc_tried;    // the round's arrays do not fit: collect, then redo the round
  int gc_list[MN_GCL]; int gc_dst[MN_GCL]; int gc_src[MN_GCL]; int gc_n;  // large arrays of one scan block, copied cooperatively
  long long st_gcs;
  int hash_ovf_n;           // entries of the overflow cache below (tombstones included)
  int ovf_lo[MN_OVF]; int ovf_hi[MN_OVF]; int ovf_rec[MN_OVF];  // records that fit neither hash bucket
  int failed;               // sticky copy of ctl->status != 0
  int cold_empty;  // 1: nothing outside `hot` -> every new entry goes to hot
  float b_mp; int b_lo; int b_hi;  // bound: entries popping strictly after it are cold
  int path[32]; int path_n;
  long long cyc[MN_NCYC]; long long cyc_t0;
  long long st_rounds, st_events, st_merges, st_restores, st_invalid, st_solo, st_refills,
      st_flushes, st_splits, st_pairs, st_cut_conf, st_cut_casc, st_cut_cap, st_requeues;
};

struct MnMergeArgs {
  int C, K, N, W;
  int H;  // candidates per round (<= MN_H; smaller when the staged class vectors would not fit)
  float omf, mlb;
  MnOffsets off;
  long long max_rounds;  // safety guard (0 = none)
};

#define HOT_MP(i) sm.hot_mp[sm.hsel][i]
#define HOT_LO(i) sm.hot_lo[sm.hsel][i]
#define HOT_HI(i) sm.hot_hi[sm.hsel][i]
#define HOT_REC(i) sm.hot_rec[sm.hsel][i]
#if defined(__CUDA_ARCH__)
#define MN_CLZ(x) __clz((int)(x))
#define MN_POPC(x) __popc((unsigned)(x))
#else
#define MN_CLZ(x) __builtin_clz((unsigned)(x))
#define MN_POPC(x) __builtin_popcount((unsigned)(x))
#endif
#if defined(__CUDA_ARCH__) && defined(MN_PHASE_CYCLES)  // per-phase cycle buckets cost ~4 %: profiling builds only
#define MN_TIC() do { if (MN_T0) sm.cyc_t0 = clock64(); } while (0)
#define MN_TOC(k) do { if (MN_T0) { long long t__ = clock64(); sm.cyc[k] += t__ - sm.cyc_t0; sm.cyc_t0 = t__; } } while (0)
#else
#define MN_TIC() ((void)0)
#define MN_TOC(k) ((void)0)
#endif
#if defined(MN_EMUL_WATCH)
#define MN_WATCH(rec, ...) do { if ((rec) == MN_EMUL_WATCH) { fprintf(stderr, "W%d: ", (int)(rec)); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)
#else
#define MN_WATCH(rec, ...) ((void)0)
#endif
#define MN_FOR(i, n) for (int i = MN_TID; i < (n); i += MN_NT)
#define MN_T0 (MN_TID == 0)

MN_D void mn_fail_at(const MnImage& im, int code, int line) {
  if (im.ctl->status == MN_OK) { im.ctl->status = code; im.ctl->fail_line = line; }
}
#define mn_fail(im, code) do { mn_fail_at((im), (code), __LINE__); sm.failed = 1; } while (0)

// ------------------------------------------------------------------------------------------------
// sorting
MN_D void mn_sort_sb(MnSm& sm, int n2) {  // n2 = power of two, pads carry mp = MN_NEG_INF
  for (int k = 2; k <= n2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      MN_FOR(i, n2) {
        int l = i ^ j;
        if (l > i) {
          bool up = ((i & k) == 0);
          bool lbi = mn_before(sm.sb_mp[l], sm.sb_lo[l], sm.sb_hi[l], sm.sb_mp[i], sm.sb_lo[i], sm.sb_hi[i]);
          bool ibl = mn_before(sm.sb_mp[i], sm.sb_lo[i], sm.sb_hi[i], sm.sb_mp[l], sm.sb_lo[l], sm.sb_hi[l]);
          if (up ? lbi : ibl) {
            float t = sm.sb_mp[i]; sm.sb_mp[i] = sm.sb_mp[l]; sm.sb_mp[l] = t;
            int u = sm.sb_lo[i]; sm.sb_lo[i] = sm.sb_lo[l]; sm.sb_lo[l] = u;
            u = sm.sb_hi[i]; sm.sb_hi[i] = sm.sb_hi[l]; sm.sb_hi[l] = u;
            u = sm.sb_rec[i]; sm.sb_rec[i] = sm.sb_rec[l]; sm.sb_rec[l] = u;
          }
        }
      }
      MN_SYNC();
    }
}
MN_D int mn_pow2_ge(int n) { int p = 1; while (p < n) p <<= 1; return p; }

// position from a shared counter, one atomic per group of converged lanes (all of them name the same counter)
MN_D int mn_agg_inc(int* ctr) {
#if defined(__CUDA_ARCH__)
  const unsigned act = __activemask();
  const int lane = threadIdx.x & 31, leader = __ffs(act) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(ctr, __popc(act));
  base = __shfl_sync(act, base, leader);
  return base + __popc(act & ((1u << lane) - 1u));
#else
  return MN_ATOMIC_ADD(ctr, 1);
#endif
}

// ------------------------------------------------------------------------------------------------
// queue tree
MN_D int mn_root_of(float mp) {
  uint32_t b = mn_f2u(mp);
  if (b < MN_ROOT_LO_BITS) return 0;
  if (b >= MN_ROOT_HI_BITS) return MN_NROOTS - 1;
  return (int)((b - MN_ROOT_LO_BITS) >> MN_ROOT_SHIFT) + 1;
}
// MN_TREE_BITS-bit digit `level` (1-based, below the root) of the 80-bit pop-order key [~mpbits:32][tie u:24][D:24].
// Regular roots fix the top 32 - MN_ROOT_SHIFT key bits, where their digits start; the two open-ended roots
// start at bit 0.  Smaller digit = pops first.
MN_D int mn_digit(int root, int level, float mp, int lo, int hi) {
  int start = (root == 0 || root == MN_NROOTS - 1) ? 0 : 32 - MN_ROOT_SHIFT;
  int pos = start + MN_TREE_BITS * (level - 1);  // bit offset from the top of the 80-bit key
  const unsigned long long tie = mn_tie(lo, hi);  // 48 bits: (u, D)
  unsigned long long hi64 = ((unsigned long long)(~mn_f2u(mp)) << 32) | (tie >> 16);
  unsigned long long lo16 = tie & 0xFFFFull;
  int d = 0;
  for (int b = 0; b < MN_TREE_BITS; b++) {
    int p = pos + b;
    int bit;
    if (p < 64) bit = (int)((hi64 >> (63 - p)) & 1ull);
    else if (p < 80) bit = (int)((lo16 >> (79 - p)) & 1ull);
    else bit = 0;
    d = (d << 1) | bit;
  }
  return d;
}
MN_D int mn_max_level(int root) {
  const int start = (root == 0 || root == MN_NROOTS - 1) ? 0 : 32 - MN_ROOT_SHIFT;
  return (80 - start + MN_TREE_BITS - 1) / MN_TREE_BITS;
}

// Queue chunks come from one arena that starts with the sorted initial keys: the keys below the cursor
// are consumed, so their 1 KB blocks (MN_QCH 16-byte entries = 2 * MN_QCH keys) are handed out as chunks
// (ids below qc_low_n, available up to qc_low_avail); only the early demand, before enough of the
// initial array is consumed, needs chunks of its own (ids from qc_low_n up to qc_cap).
MN_D int mn_qc_alloc(const MnImage& im, MnSm& sm) {  // called only from phases that never free
  int t = MN_ATOMIC_SUB(&sm.qc_free_top, 1);
  if (t > 0) return im.qc_free[t - 1];
  MN_ATOMIC_ADD(&sm.qc_free_top, 1);
  if (sm.qc_low_bump < sm.qc_low_avail) {
    int c = MN_ATOMIC_ADD(&sm.qc_low_bump, 1);
    if (c < sm.qc_low_avail) return c;
    MN_ATOMIC_SUB(&sm.qc_low_bump, 1);
  }
  int c = MN_ATOMIC_ADD(&sm.qc_bump, 1);
  if (c >= im.qc_cap) { mn_fail(im, MN_ERR_Q_POOL); return -1; }
  return c;
}
MN_D void mn_qc_free(const MnImage& im, MnSm& sm, int c) {  // called only from phases that never allocate
  int t = MN_ATOMIC_ADD(&sm.qc_free_top, 1);
  im.qc_free[t] = c;
}

// conflict-table helpers are also used as a small node -> group hash by distribute()
MN_D int mn_ct_slot(MnSm& sm, int obj);
MN_D int mn_ct_find(const MnSm& sm, int obj);

// Tree node (int4): x = first chunk, y = last chunk, z = entries below (a leaf: its own), w = first of
// its 64 children (-1: leaf).  A leaf's entries fill its chunks in order, so the fill of the last
// chunk follows from z; tn_dir holds the first 8 chunk ids of a leaf (8 * MN_QCH = MN_LEAFCAP
// entries: what a refill loads) so that they can be fetched together instead of chain-walked.

// Append the n entries staged in sm.sb_* (sb_node[i] = destination leaf) to their leaves.
// No sorting: entries are grouped by destination through a shared-memory hash, one thread per
// group reserves the space (linking fresh chunks), then every entry writes itself.
MN_D void mn_distribute(const MnImage& im, MnSm& sm, int n) {
  if (n <= 0) return;
  MN_FOR(i, MN_CT) { sm.ct_obj[i] = -1; sm.ct_w[i] = -1; }
  if (MN_T0) { sm.ds_ngroups = 0; sm.ds_ndir = 0; }
  MN_SYNC();
  MN_FOR(i, n) { if (mn_ct_slot(sm, sm.sb_node[i]) < 0) mn_fail(im, MN_ERR_INTERNAL); }
  MN_SYNC();
  MN_FOR(s, MN_CT) {
    if (sm.ct_obj[s] != -1) {
      int g = MN_ATOMIC_ADD(&sm.ds_ngroups, 1);
      sm.ct_w[s] = g; sm.w.ds.node[g] = sm.ct_obj[s]; sm.w.ds.cnt[g] = 0;
    }
  }
  MN_SYNC();
  MN_FOR(i, n) {
    int s = mn_ct_find(sm, sm.sb_node[i]);
    int g = s >= 0 ? sm.ct_w[s] : 0;
    int local = MN_ATOMIC_ADD(&sm.w.ds.cnt[g], 1);
    sm.w.ds.el[i] = (g << 16) | local;
  }
  MN_SYNC();
  const int ngroups = sm.ds_ngroups;
  MN_FOR(g, ngroups) {
    const int node = sm.w.ds.node[g], cnt = sm.w.ds.cnt[g];
    const int4 nd = im.tn[node];
    const int have = nd.z, tail = nd.y;
    const int fill = (tail >= 0 && have > 0) ? ((have - 1) % MN_QCH) + 1 : MN_QCH;
    const int room = MN_QCH - fill;
    const int nch0 = (tail >= 0 && have > 0) ? (have + MN_QCH - 1) / MN_QCH : 0;
    const int nnew = cnt > room ? (cnt - room + MN_QCH - 1) / MN_QCH : 0;
    const int base = nnew ? MN_ATOMIC_ADD(&sm.ds_ndir, nnew) : 0;
    sm.w.ds.tail[g] = tail;
    sm.w.ds.fd[g] = fill | (base << 8);
    int prev = tail, head = (tail >= 0 && have > 0) ? nd.x : -1;
    if (prev >= 0 && !(have > 0)) prev = -1;
    for (int j = 0; j < nnew; j++) {
      int c = mn_qc_alloc(im, sm);
      if (base + j < MN_SB + 128) sm.w.ds.dir[base + j] = c;
      if (c < 0) break;
      im.qc_next[c] = -1;
      if (prev >= 0) im.qc_next[prev] = c; else head = c;
      if (nch0 + j < 8) im.tn_dir[(size_t)node * 8 + nch0 + j] = c;
      prev = c;
    }
    im.tn[node] = make_int4(head, prev, have + cnt, nd.w);
  }
  MN_SYNC();
  MN_FOR(i, n) {
    int g = sm.w.ds.el[i] >> 16, local = sm.w.ds.el[i] & 0xffff;
    int fill = sm.w.ds.fd[g] & 0xff, base = sm.w.ds.fd[g] >> 8;
    int pos = fill + local, chunk, slot;
    if (pos < MN_QCH) { chunk = sm.w.ds.tail[g]; slot = pos; }
    else { int q = pos - MN_QCH; chunk = sm.w.ds.dir[base + q / MN_QCH]; slot = q % MN_QCH; }
    if (chunk >= 0)
      im.q_ent[(size_t)chunk * MN_QCH + slot] = make_uint4(mn_f2u(sm.sb_mp[i]), (uint32_t)sm.sb_rec[i], (uint32_t)sm.sb_lo[i], (uint32_t)sm.sb_hi[i]);
  }
  MN_SYNC();
}

// destination leaf of an entry: descend from its root through split nodes, counting it on the way
MN_D int mn_descend_for_insert(const MnImage& im, MnSm& sm, float mp, int lo, int hi) {
  int root = mn_root_of(mp);
  MN_ATOMIC_OR(&sm.root_bits[root >> 5], 1u << (root & 31));
  MN_ATOMIC_OR(&sm.root_sum[root >> 10], 1u << ((root >> 5) & 31));
  int node = root, level = 0;
  int child = im.tn[node].w;
  while (child >= 0) {
    MN_ATOMIC_ADD(&im.tn[node].z, 1);
    level++;
    node = child + mn_digit(root, level, mp, lo, hi);
    child = im.tn[node].w;
  }
  return node;
}

// flush the insert buffer into the tree (batches of MN_SB)
MN_D void mn_flush_ins(const MnImage& im, MnSm& sm) {
  MN_SYNC();
  const int total = sm.nins;
  if (total == 0) return;
  for (int off = 0; off < total; off += MN_SB) {
    const int n = total - off < MN_SB ? total - off : MN_SB;
    MN_FOR(i, n) {
      int s = off + i;
      sm.sb_mp[i] = sm.ins_mp[s]; sm.sb_lo[i] = sm.ins_lo[s]; sm.sb_hi[i] = sm.ins_hi[s]; sm.sb_rec[i] = sm.ins_rec[s];
      sm.sb_node[i] = mn_descend_for_insert(im, sm, sm.ins_mp[s], sm.ins_lo[s], sm.ins_hi[s]);
    }
    MN_SYNC();
    mn_distribute(im, sm, n);
  }
  if (MN_T0) {
    sm.tree_entries += total;
    sm.nins = 0;
    sm.st_flushes++;
    if (sm.tree_entries > sm.peak_entries) sm.peak_entries = sm.tree_entries;
    const int used = (sm.qc_bump - im.qc_low_n) + sm.qc_low_bump - (sm.qc_free_top > 0 ? sm.qc_free_top : 0);
    if (used > sm.peak_chunks) sm.peak_chunks = used;
  }
  MN_SYNC();
}

// What a queue entry (emp, elo, ehi) is against its record (raw words rw, mn_layout.h):
//   DROP     the record is dead, or no entry is expected (NONE), or it is not the expected one;
//   exact    state EXACT, emp == mp and the key matches: the reference's valid pop (cc:554-559);
//   REQUEUE  state ABOVE, emp > mp >= 0: a guard of a record whose priority was re-stored lower;
//   UNGUARD  state ABOVE, mp < 0: a guard of a record that went dormant.
// Returns MN_K_DROP / MN_K_RESTORE (meaning "exact") / MN_K_REQUEUE / MN_K_UNGUARD.
MN_D int mn_entry_state(float emp, int elo, int ehi, uint4 rw) {
  const int lo = mn_rec_lo(rw.x);
  if (lo < 0) return MN_K_DROP;
  const uint32_t g = mn_rec_guard(rw.x);
  const float mp = mn_u2f(rw.w);
  if (g == MN_G_EXACT) return (mp == emp && lo == elo && mn_rec_hi(rw.y) == ehi) ? MN_K_RESTORE : MN_K_DROP;
  if (g == MN_G_ABOVE && emp > mp) return mp >= 0.0f ? MN_K_REQUEUE : MN_K_UNGUARD;
  return MN_K_DROP;
}

// chunk ids [k0, k0 + n) of leaf `node` into sm.cw_chunk (thread 0 follows the chain past the directory;
// `prev` = id of chunk k0 - 1 or -1)
MN_D void mn_leaf_chunks(const MnImage& im, MnSm& sm, int node, int k0, int n, int prev) {
  MN_FOR(i, n) if (k0 + i < 8) sm.cw_chunk[i] = im.tn_dir[(size_t)node * 8 + k0 + i];
  MN_SYNC();
  if (MN_T0 && k0 + n > 8) {
    int i0 = k0 < 8 ? 8 - k0 : 0;
    int c = i0 > 0 ? sm.cw_chunk[i0 - 1] : prev;
    for (int i = i0; i < n; i++) { c = c >= 0 ? im.qc_next[c] : -1; sm.cw_chunk[i] = c; }
  }
  MN_SYNC();
}

// clear the bit of an emptied root (thread 0)
MN_D void mn_root_clear(MnSm& sm, int root) {
  sm.root_bits[root >> 5] &= ~(1u << (root & 31));
  if (sm.root_bits[root >> 5] == 0) sm.root_sum[root >> 10] &= ~(1u << ((root >> 5) & 31));
}

// Split leaf `node` (at depth `level` under `root`, `have` entries): its entries move to 64 children by
// the next digit.  Entries are validated on the way (dead ones are dropped: garbage collection).
MN_D void mn_split_leaf(const MnImage& im, MnSm& sm, int root, int node, int level, int have) {
  if (MN_T0) {
    int base = MN_ATOMIC_ADD(&sm.tn_bump, MN_TREE_FANOUT);
    if (base + MN_TREE_FANOUT > im.tn_cap) { mn_fail(im, MN_ERR_TREE_POOL); base = -1; }
    sm.tmp0 = base;
    sm.tmp2 = 0;  // surviving entries
    sm.st_splits++;
  }
  MN_SYNC();
  const int base = sm.tmp0;
  if (base < 0) return;
  MN_FOR(i, MN_TREE_FANOUT) im.tn[base + i] = make_int4(-1, -1, 0, -1);
  MN_SYNC();
  const int nch = (have + MN_QCH - 1) / MN_QCH;
  const int per = MN_SB / MN_QCH < MN_CW ? MN_SB / MN_QCH : MN_CW;
  int prev = -1;
  for (int k0 = 0; k0 < nch; k0 += per) {
    const int n = nch - k0 < per ? nch - k0 : per;
    mn_leaf_chunks(im, sm, node, k0, n, prev);
    if (MN_T0) sm.npr = 0;
    MN_SYNC();
    prev = sm.cw_chunk[n - 1];
    const int ents = (have - k0 * MN_QCH) < n * MN_QCH ? (have - k0 * MN_QCH) : n * MN_QCH;
    MN_FOR(i, ents) {
      const int c = sm.cw_chunk[i / MN_QCH], s = i % MN_QCH;
      if (c < 0) { mn_fail(im, MN_ERR_INTERNAL); continue; }
      uint4 e = im.q_ent[(size_t)c * MN_QCH + s];
      int rec = (int)e.y;
      const uint4 rw = mn_load_rec(im, rec);
      float emp = mn_u2f(e.x);
      int st = mn_entry_state(emp, (int)e.z, (int)e.w, rw);
      if (st == MN_K_UNGUARD) MN_REC(im, rec).x = mn_rec_with_guard(rw.x, MN_G_NONE);
      else if (st != MN_K_DROP) {  // the entry keeps its own key (a guard stays where it is)
        int p = mn_agg_inc(&sm.npr);
        sm.sb_mp[p] = emp; sm.sb_lo[p] = (int)e.z; sm.sb_hi[p] = (int)e.w; sm.sb_rec[p] = rec;
        sm.sb_node[p] = base + mn_digit(root, level + 1, emp, (int)e.z, (int)e.w);
      }
    }
    MN_SYNC();
    const int m = sm.npr;
    mn_distribute(im, sm, m);
    MN_FOR(i, n) if (sm.cw_chunk[i] >= 0) mn_qc_free(im, sm, sm.cw_chunk[i]);
    if (MN_T0) sm.tmp2 += m;
    MN_SYNC();
    if (sm.failed) return;
  }
  if (MN_T0) {
    // the node now holds only the survivors, below it; fix the counts on the path
    const int removed = have - sm.tmp2;
    for (int i = 0; i < sm.path_n; i++) { sm.path_cnt[i] -= removed; MN_ATOMIC_SUB(&im.tn[sm.path[i]].z, removed); }
    sm.tree_entries -= removed;
    im.tn[node] = make_int4(-1, -1, sm.tmp2, base);
    const int root_cnt = sm.path_n > 0 ? sm.path_cnt[0] : sm.tmp2;
    if (root_cnt <= 0) mn_root_clear(sm, root);
  }
  MN_SYNC();
}

// Find the first non-empty leaf in pop order; split it while it is too large.  Leaves the path
// (ancestors, root first, with their counts) in sm.path / sm.path_cnt, the leaf's node in sm.leaf_nd,
// and returns the leaf, -1 when the tree is empty, or -2 when the leaf needs a split that the caller
// did not allow.  The root bitmap is exact, so finding the root costs no memory access.
MN_D int mn_top_leaf(const MnImage& im, MnSm& sm, int* root_out, bool allow_split) {
  for (int guard = 0; guard < 4096; guard++) {
    MN_SYNC();
    if (MN_T0) {
      int root = -1;
      const int nsum = ((MN_NROOTS + 31) / 32 + 31) / 32;
      for (int s = nsum - 1; s >= 0 && root < 0; s--) {
        uint32_t sv = sm.root_sum[s];
        if (sv) {
          int w = s * 32 + 31 - MN_CLZ(sv);
          uint32_t bits = sm.root_bits[w];
          if (bits) root = w * 32 + 31 - MN_CLZ(bits);
          else sm.root_sum[s] = sv & ~(1u << (w & 31));  // (stale summary bit)
          if (root < 0) s++;                                // look at this summary word again
        }
      }
      sm.tmp0 = root;
      sm.path_n = 0;
    }
    MN_SYNC();
    const int root = sm.tmp0;
    if (root < 0) return -1;
    int node = root, level = 0;
    int4 nd = im.tn[node];
    if (nd.z <= 0) {  // (defensive: a root bit without entries)
      MN_SYNC();
      if (MN_T0) mn_root_clear(sm, root);
      continue;
    }
    bool bad = false;
    while (nd.w >= 0) {  // descend: the 64 children of a split node are read by 64 threads at once
      const int cb = nd.w;
      MN_FOR(d, MN_TREE_FANOUT) sm.scan_nd[d] = im.tn[cb + d];
      MN_SYNC();
      if (MN_T0) {
        int found = -1;
        for (int d = 0; d < MN_TREE_FANOUT; d++)
          if (sm.scan_nd[d].z > 0) { found = d; break; }
        sm.path[sm.path_n] = node; sm.path_cnt[sm.path_n] = nd.z; sm.path_n++;
        sm.tmp1 = found;
      }
      MN_SYNC();
      const int found = sm.tmp1;
      if (found < 0 || level > 30) { bad = true; break; }
      nd = sm.scan_nd[found];
      node = cb + found;
      level++;
      MN_SYNC();  // scan_nd is reused by the next level
    }
    if (bad) { mn_fail(im, MN_ERR_INTERNAL); return -1; }
    *root_out = root;
    if (nd.z <= MN_LEAFCAP || level >= mn_max_level(root)) {
      MN_SYNC();
      if (MN_T0) sm.leaf_nd = nd;
      MN_SYNC();
      return node;
    }
    if (!allow_split) return -2;  // the caller's staging buffers are in use
    MN_TOC(MN_CY_REFILL);
    mn_split_leaf(im, sm, root, node, level, nd.z);
    MN_TOC(MN_CY_SPLIT);
    if (sm.failed) return -1;
  }
  mn_fail(im, MN_ERR_LIMIT);
  return -1;
}

// decode the i-th sorted initial entry
MN_D void mn_decode_init(const MnMergeArgs& A, uint64_t key, float* mp, int* lo, int* hi, int* rec) {
  uint32_t ord = (uint32_t)(key & ((1ull << MN_ORD_BITS) - 1));
  *mp = mn_u2f(~(uint32_t)(key >> MN_ORD_BITS));
  const int rank = (int)(ord & 15u);
  int k = A.off.k_of_rank[rank];
  int d = A.off.delta[k];
  const uint32_t D = (uint32_t)(d > 0 ? d : -d);
  int l = (int)mn_brev24(((ord >> 4) - MN_TIE_MUL * D) & 0xFFFFFFu);  // (the 24-bit reversal is an involution)
  int h = l + (int)D;
  *lo = l; *hi = h;
  int p = d > 0 ? l : h;
  *rec = p * A.K + k;
}

// exclusive prefix sum of n non-negative ints (all threads call it); returns the total
MN_D int mn_exclusive_scan(MnSm& sm, const int* vals, int* out, int n) {
  const int nt = MN_NT, t = MN_TID;
  const int per = (n + nt - 1) / nt;
  const int b = t * per, e = b + per < n ? b + per : n;
  int ssum = 0;
  for (int i = b; i < e; i++) ssum += vals[i];
  if (t < MN_SB) sm.w.ds.cnt[t] = ssum;
  MN_SYNC();
#if defined(__CUDA_ARCH__)
  if (t < 32) {  // warp 0: each lane scans a run of the per-thread sums, the runs are joined by a shuffle scan
    const int lim = nt < MN_SB ? nt : MN_SB;
    const int run = (lim + 31) / 32;
    const int b2 = t * run, e2 = b2 + run < lim ? b2 + run : lim;
    int s2 = 0;
    for (int k = b2; k < e2; k++) s2 += sm.w.ds.cnt[k];
    int incl = s2;
    for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, incl, d); if (t >= d) incl += o; }
    int acc = incl - s2;
    for (int k = b2; k < e2; k++) { int v = sm.w.ds.cnt[k]; sm.w.ds.cnt[k] = acc; acc += v; }
    if (t == 31) sm.tmp3 = incl;
  }
#else
  if (MN_T0) {
    int acc = 0;
    for (int k = 0; k < nt && k < MN_SB; k++) { int v = sm.w.ds.cnt[k]; sm.w.ds.cnt[k] = acc; acc += v; }
    sm.tmp3 = acc;
  }
#endif
  MN_SYNC();
  int acc = t < MN_SB ? sm.w.ds.cnt[t] : 0;
  for (int i = b; i < e; i++) { int v = vals[i]; out[i] = acc; acc += v; }
  MN_SYNC();
  return sm.tmp3;
}

MN_D void mn_hot_update(const MnImage& im, MnSm& sm, int cut, bool lead_sync = true, bool trail_sync = true);
MN_D void mn_push_entry(MnSm& sm, float mp, int lo, int hi, int rec);
// Refill the (empty) hot buffer.  Afterwards: hot holds every entry popping before-or-at `bound`.
MN_D void mn_refill(const MnImage& im, MnSm& sm, const MnMergeArgs& A) {
  if (MN_T0) sm.st_refills++;
  for (int guard = 0; guard < (1 << 20); guard++) {
    MN_SYNC();
    mn_flush_ins(im, sm);  // (also on retries: a previous pass may have pushed leaf entries back)
    MN_TOC(MN_CY_FLUSH);
    if (sm.failed) return;
    // ---- load + validate successive top leaves (each pops entirely before the next) until a
    //      useful number of entries is staged; their chunks are recycled ----
    int nleaf = 0;
    if (MN_T0) { sm.npr = 0; sm.nlf = 0; sm.lf_start[0] = 0; }
    MN_SYNC();
    // (MN_LF bounds the leaves that CONTRIBUTE entries -- lf_start has a slot per contributing leaf --, not the
    //  leaves visited: late in a run whole leaves are stale, and stopping after MN_LF fruitless leaves with
    //  nothing staged would read "no leaf entry" below as "the tree is empty" and lose what is behind them)
    for (int lguard = 0; lguard < (1 << 22) && sm.nlf < MN_LF && nleaf < MN_REFILL_TARGET; lguard++) {
      int root = 0;
      int leaf = mn_top_leaf(im, sm, &root, nleaf == 0);  // splitting reuses the staging buffers
      if (nleaf == 0) { MN_SYNC(); if (MN_T0) sm.npr = 0; MN_SYNC(); }  // (a split used npr)
      if (leaf < 0) break;
      if (sm.failed) return;
      const int have = sm.leaf_nd.z;
      if (nleaf + have > MN_SB - MN_REFILL_STATIC_MIN) {  // keep room for initial entries
        if (nleaf == 0) mn_fail(im, MN_ERR_INTERNAL);  // an unsplittable leaf larger than the buffer
        break;
      }
      const int nch = (have + MN_QCH - 1) / MN_QCH;
      if (nch > MN_CW) { mn_fail(im, MN_ERR_INTERNAL); break; }
      mn_leaf_chunks(im, sm, leaf, 0, nch, -1);
      MN_FOR(i, have) {
        const int c = sm.cw_chunk[i / MN_QCH], s = i % MN_QCH;
        if (c < 0) { mn_fail(im, MN_ERR_INTERNAL); continue; }
        uint4 e = im.q_ent[(size_t)c * MN_QCH + s];
        int rec = (int)e.y;
        const uint4 rw = mn_load_rec(im, rec);
        float emp = mn_u2f(e.x);
        int st = mn_entry_state(emp, (int)e.z, (int)e.w, rw);
        if (st == MN_K_UNGUARD) MN_REC(im, rec).x = mn_rec_with_guard(rw.x, MN_G_NONE);
        else if (st != MN_K_DROP) {
          int p = mn_agg_inc(&sm.npr);
          if (p < MN_SB) { sm.sb_mp[p] = emp; sm.sb_lo[p] = (int)e.z; sm.sb_hi[p] = (int)e.w; sm.sb_rec[p] = rec; }
        }
      }
      MN_SYNC();
      MN_FOR(i, nch) if (sm.cw_chunk[i] >= 0) mn_qc_free(im, sm, sm.cw_chunk[i]);
      if (MN_T0) {
        for (int i = 0; i < sm.path_n; i++) MN_ATOMIC_SUB(&im.tn[sm.path[i]].z, have);
        im.tn[leaf] = make_int4(-1, -1, 0, -1);
        sm.tree_entries -= have;
        const int root_cnt = sm.path_n > 0 ? sm.path_cnt[0] - have : 0;
        if (root_cnt <= 0) mn_root_clear(sm, root);
        if (sm.npr > MN_SB) { mn_fail(im, MN_ERR_INTERNAL); sm.npr = MN_SB; }
        if (sm.npr > sm.lf_start[sm.nlf]) { sm.nlf++; sm.lf_start[sm.nlf] = sm.npr; }  // its entries: a contiguous sb range
      }
      MN_SYNC();
      nleaf = sm.npr;
    }
    MN_TOC(MN_CY_RF_LEAVES);
    // ---- the leaf's last entry in pop order bounds what may be taken from the sorted initial
    //      entries: everything else in the tree pops after it ----
    {  // (leaves were loaded in pop order: the last one holds it; tournament over its sb range)
      const int r0 = sm.nlf > 0 ? sm.lf_start[sm.nlf - 1] : 0, m = nleaf - r0;
      MN_FOR(i, m) sm.red_idx[i] = r0 + i;
      MN_SYNC();
      for (int len = m; len > 1;) {
        const int half = (len + 1) >> 1;
        MN_FOR(i, len - half) {
          const int a = sm.red_idx[i], b = sm.red_idx[half + i];
          if (mn_before(sm.sb_mp[a], sm.sb_lo[a], sm.sb_hi[a], sm.sb_mp[b], sm.sb_lo[b], sm.sb_hi[b])) sm.red_idx[i] = b;
        }
        MN_SYNC();
        len = half;
      }
      if (MN_T0) { sm.tmp0 = m > 0 ? sm.red_idx[0] : -1; sm.tmp2 = 0; }
      MN_SYNC();
    }
    const int w = sm.tmp0;
    const float lmp = w >= 0 ? sm.sb_mp[w] : 0.f; const int llo = w >= 0 ? sm.sb_lo[w] : 0, lhi = w >= 0 ? sm.sb_hi[w] : 0;
    const int sc = sm.static_cursor, ninit = sm.n_init;
    const int room = MN_HC - nleaf;
    int navail = ninit - sc; if (navail > room) navail = room;
    // count the initial entries (a prefix, they are sorted) that pop before the leaf's last entry
    MN_FOR(i, navail) {
      float mp; int lo, hi, rec;
      mn_decode_init(A, im.init_keys[sc + i], &mp, &lo, &hi, &rec);
      bool take = (w < 0) || mn_before(mp, lo, hi, lmp, llo, lhi);
      if (take) mn_agg_inc(&sm.tmp2);
    }
    MN_SYNC();
    const int ntake = sm.tmp2;
    // does an untaken initial entry still pop before the leaf's last entry?
    bool more_before = false;
    if (w >= 0 && ntake == navail && sc + navail < ninit) {
      float mp; int lo, hi, rec;
      mn_decode_init(A, im.init_keys[sc + navail], &mp, &lo, &hi, &rec);
      more_before = mn_before(mp, lo, hi, lmp, llo, lhi);
    }
    MN_FOR(i, ntake) {
      float mp; int lo, hi, rec;
      mn_decode_init(A, im.init_keys[sc + i], &mp, &lo, &hi, &rec);
      const uint4 rw = mn_load_rec(im, rec);
      int st = mn_entry_state(mp, lo, hi, rw);
      if (st == MN_K_UNGUARD) { MN_REC(im, rec).x = mn_rec_with_guard(rw.x, MN_G_NONE); st = MN_K_DROP; }
      int p = nleaf + i;
      sm.sb_mp[p] = st != MN_K_DROP ? mp : MN_NEG_INF; sm.sb_lo[p] = lo; sm.sb_hi[p] = hi; sm.sb_rec[p] = rec;
    }
    MN_SYNC();
    // the bound
    if (MN_T0) {
      if (w < 0) {  // tree empty: the last taken initial entry bounds the rest of the array
        if (ntake > 0) {
          float mp; int lo, hi, rec;
          mn_decode_init(A, im.init_keys[sc + ntake - 1], &mp, &lo, &hi, &rec);
          sm.b_mp = mp; sm.b_lo = lo; sm.b_hi = hi;
        }
        sm.cold_empty = (sc + ntake >= ninit) ? 1 : 0;
      } else if (more_before && ntake == 0) {
        mn_fail(im, MN_ERR_INTERNAL);
      } else if (!more_before) {
        sm.b_mp = lmp; sm.b_lo = llo; sm.b_hi = lhi;
        sm.cold_empty = 0;
      } else {  // hot is full of earlier initial entries: leaf entries after the last taken one stay cold
        float mp; int lo, hi, rec;
        mn_decode_init(A, im.init_keys[sc + ntake - 1], &mp, &lo, &hi, &rec);
        sm.b_mp = mp; sm.b_lo = lo; sm.b_hi = hi;
        sm.cold_empty = 0;
      }
      sm.static_cursor = sc + ntake;
      sm.qc_low_avail = (int)(((long long)sm.static_cursor * 8) / (MN_QCH * 16));
      if (sm.qc_low_avail > im.qc_low_n) sm.qc_low_avail = im.qc_low_n;
      sm.tmp1 = 0; sm.tmp0 = 0;
    }
    MN_SYNC();
    // ---- guards in the batch (record re-stored lower since the entry was queued) are replaced by
    //      their exact entry now, in bulk, instead of costing a window slot later ----
    MN_FOR(i, nleaf + ntake) {
      if (sm.sb_mp[i] > MN_NEG_INF) {
        const int rec = sm.sb_rec[i];
        const uint4 rw = mn_load_rec(im, rec);
        if (mn_entry_state(sm.sb_mp[i], sm.sb_lo[i], sm.sb_hi[i], rw) == MN_K_REQUEUE) {
          MN_REC(im, rec).x = mn_rec_with_guard(rw.x, MN_G_EXACT);
          MN_ATOMIC_ADD(&sm.tmp1, 1);
          mn_push_entry(sm, mn_u2f(rw.w), mn_rec_lo(rw.x), mn_rec_hi(rw.y), rec);  // cold -> insert buffer; still hot-bound -> ne, merged below
          sm.sb_mp[i] = MN_NEG_INF;
        }
      }
    }
    MN_SYNC();
    if (MN_T0) { sm.st_requeues += sm.tmp1; sm.tmp1 = 0; }
    if (more_before) {
      // push the leaf entries that pop after the bound back to the insert buffer
      MN_FOR(i, nleaf) {
        if (sm.sb_mp[i] > MN_NEG_INF && mn_before(sm.b_mp, sm.b_lo, sm.b_hi, sm.sb_mp[i], sm.sb_lo[i], sm.sb_hi[i])) {
          int p = MN_ATOMIC_ADD(&sm.nins, 1);
          sm.ins_mp[p] = sm.sb_mp[i]; sm.ins_lo[p] = sm.sb_lo[i]; sm.ins_hi[p] = sm.sb_hi[i]; sm.ins_rec[p] = sm.sb_rec[i];
          sm.sb_mp[i] = MN_NEG_INF;
        }
      }
      MN_SYNC();
    }
#ifdef MN_EMUL_TRACE
    fprintf(stderr, "refill: nleaf %d ntake %d more_before %d nins %d cold_empty %d bound %.9g %d %d sc %d ninit %d tree %d\n", nleaf, ntake, (int)more_before, sm.nins, sm.cold_empty, sm.b_mp, sm.b_lo, sm.b_hi, sm.static_cursor, ninit, sm.tree_entries);
#endif
    MN_TOC(MN_CY_RF_INIT);
    const int n = nleaf + ntake;
    if (n == 0) {
      if (MN_T0) sm.nhot = 0;
      MN_SYNC();
      return;  // nothing left anywhere (cold_empty set above)
    }
    {
      // ---- order the batch: the initial entries are a sorted run already (compacted with a scan); the
      //      leaf entries are ranked inside their own leaf (leaves were loaded in pop order, so a
      //      leaf's entries all pop before the next leaf's), which needs no barrier-heavy sort; then
      //      the two runs are merged by binary search.  Duplicates of one record stay adjacent and
      //      are dropped when they are popped. ----
      const int nlf = sm.nlf;
      MN_FOR(k, nlf) sm.lf_cnt[k] = 0;
      MN_FOR(i, ntake) sm.sb_node[i] = sm.sb_mp[nleaf + i] > MN_NEG_INF ? 1 : 0;
      MN_SYNC();
      const int ns = mn_exclusive_scan(sm, sm.sb_node, sm.w.ds.el, ntake);
      const int cur = sm.hsel, dst = sm.hsel ^ 1;
      MN_FOR(i, ntake) {
        if (sm.sb_node[i]) {
          int p = sm.w.ds.el[i], q = nleaf + i;
          sm.hot_mp[cur][p] = sm.sb_mp[q]; sm.hot_lo[cur][p] = sm.sb_lo[q]; sm.hot_hi[cur][p] = sm.sb_hi[q]; sm.hot_rec[cur][p] = sm.sb_rec[q];
        }
      }
      MN_FOR(i, nleaf) {
        int rk = -1, k = 0;
        if (sm.sb_mp[i] > MN_NEG_INF) {
          int a = 0, bnd = nlf;  // leaf of entry i: last k with lf_start[k] <= i
          while (a + 1 < bnd) { int mid = (a + bnd) >> 1; if (sm.lf_start[mid] <= i) a = mid; else bnd = mid; }
          k = a;
          rk = 0;
          const float mp = sm.sb_mp[i]; const int lo = sm.sb_lo[i], hi = sm.sb_hi[i];
          for (int q = sm.lf_start[k]; q < sm.lf_start[k + 1]; q++) {
            const float qmp = sm.sb_mp[q];
            if (q == i || !(qmp > MN_NEG_INF)) continue;
            bool qb = mn_before(qmp, sm.sb_lo[q], sm.sb_hi[q], mp, lo, hi);
            bool ib = mn_before(mp, lo, hi, qmp, sm.sb_lo[q], sm.sb_hi[q]);
            if (qb || (!ib && q < i)) rk++;
          }
          MN_ATOMIC_ADD(&sm.lf_cnt[k], 1);
        }
        sm.w.ds.node[i] = k; sm.w.ds.tail[i] = rk;
      }
      MN_SYNC();
      if (MN_T0) { int acc = 0; for (int k = 0; k < nlf; k++) { sm.lf_base[k] = acc; acc += sm.lf_cnt[k]; } sm.tmp0 = acc; }
      MN_SYNC();
      const int nl = sm.tmp0;
      // the sorted leaf run goes to the (idle) tail of the pair-plan arrays, which distribute()'s scratch does not cover
      float* Lmp = sm.w.pr.mp; int* Llo = sm.w.pr.lo; int* Lhi = sm.w.pr.hi; int* Lrec = sm.w.pr.eslot;
      MN_FOR(i, nleaf) {
        const int rk = sm.w.ds.tail[i];
        if (rk >= 0) {
          const int p = sm.lf_base[sm.w.ds.node[i]] + rk;
          Lmp[p] = sm.sb_mp[i]; Llo[p] = sm.sb_lo[i]; Lhi[p] = sm.sb_hi[i]; Lrec[p] = sm.sb_rec[i];
        }
      }
      MN_SYNC();
      MN_FOR(i, nl) {  // leaf entry i -> i + #initial entries popping before-or-equal it
        float mp = Lmp[i]; int lo = Llo[i], hi = Lhi[i];
        int a = 0, bnd = ns;
        while (a < bnd) { int mid = (a + bnd) >> 1; if (mn_before(mp, lo, hi, sm.hot_mp[cur][mid], sm.hot_lo[cur][mid], sm.hot_hi[cur][mid])) bnd = mid; else a = mid + 1; }
        int p = i + a;
        sm.hot_mp[dst][p] = mp; sm.hot_lo[dst][p] = lo; sm.hot_hi[dst][p] = hi; sm.hot_rec[dst][p] = Lrec[i];
      }
      MN_FOR(j, ns) {  // initial entry j -> j + #leaf entries popping strictly before it
        float mp = sm.hot_mp[cur][j]; int lo = sm.hot_lo[cur][j], hi = sm.hot_hi[cur][j];
        int a = 0, bnd = nl;
        while (a < bnd) { int mid = (a + bnd) >> 1; if (mn_before(Lmp[mid], Llo[mid], Lhi[mid], mp, lo, hi)) a = mid + 1; else bnd = mid; }
        int p = j + a;
        sm.hot_mp[dst][p] = mp; sm.hot_lo[dst][p] = lo; sm.hot_hi[dst][p] = hi; sm.hot_rec[dst][p] = sm.hot_rec[cur][j];
      }
      MN_SYNC();
      if (MN_T0) { sm.hsel = dst; sm.nhot = nl + ns; }
      MN_SYNC();
    }
    if (sm.nne > 0) mn_hot_update(im, sm, 0);  // exact entries of requeued guards that are still hot-bound
    MN_TOC(MN_CY_RF_SORT);
    if (sm.nhot > 0 || (sm.cold_empty && sm.nins == 0)) return;
    // everything loaded was invalid: lower the bound again
  }
  mn_fail(im, MN_ERR_LIMIT);
}

// ------------------------------------------------------------------------------------------------
// conflict table
MN_D int mn_ct_slot(MnSm& sm, int obj) {
  uint32_t h = ((uint32_t)obj * 2654435761u) >> 21;  // 11 bits
  for (int i = 0; i < MN_CT; i++) {
    int s = (int)((h + (uint32_t)i) & (MN_CT - 1));
    int cur = sm.ct_obj[s];
    if (cur == obj) return s;
    if (cur == -1) {
      int old = MN_ATOMIC_CAS(&sm.ct_obj[s], -1, obj);
      if (old == -1 || old == obj) return s;
    }
  }
  return -1;
}
MN_D int mn_ct_find(const MnSm& sm, int obj) {
  uint32_t h = ((uint32_t)obj * 2654435761u) >> 21;
  for (int i = 0; i < MN_CT; i++) {
    int s = (int)((h + (uint32_t)i) & (MN_CT - 1));
    int cur = sm.ct_obj[s];
    if (cur == obj) return s;
    if (cur == -1) return -1;
  }
  return -1;
}

// record slot of the `bit`-th live-mask bit of pixel p
MN_D int mn_rec_of_bit(const MnMergeArgs& A, int p, int bit) {
  return bit < 16 ? p * A.K + bit : (p - A.off.delta[bit - 16]) * A.K + (bit - 16);
}
// clear the two live-mask bits of record slot r
MN_D void mn_clear_live(const MnImage& im, const MnMergeArgs& A, int r) {
  int p = r / A.K, k = r - p * A.K;
  MN_ATOMIC_AND(&im.obj[p].w, ~(1u << k));
  MN_ATOMIC_AND(&im.obj[p + A.off.delta[k]].w, ~(1u << (16 + k)));
}
// live-mask bits of pixel `pix` that belong to record slot r (0 when r is not a slot of pix)
MN_D uint32_t mn_own_bits(const MnMergeArgs& A, int pix, int r) {
  int p = r / A.K, k = r - p * A.K;
  uint32_t m = 0;
  if (pix == p) m |= 1u << k;
  if (pix == p + A.off.delta[k]) m |= 1u << (16 + k);
  return m;
}

// queue a created entry (cc:564,697,705): hot-bound entries are staged in ne_*, colder ones go to ins
MN_D void mn_push_entry(MnSm& sm, float mp, int lo, int hi, int rec) {
  MN_WATCH(rec, "push mp %.9g key %d %d", mp, lo, hi);
  bool cold = !sm.cold_empty && mn_before(sm.b_mp, sm.b_lo, sm.b_hi, mp, lo, hi);
#if defined(__CUDA_ARCH__)
  // the converged lanes split into the cold and the hot-bound group; each group takes its slots with one atomic
  const unsigned act = __activemask();
  const unsigned mc = __ballot_sync(act, cold);
  const unsigned mine = cold ? mc : (act & ~mc);
  const int lane = threadIdx.x & 31, leader = __ffs(mine) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(cold ? &sm.nins : &sm.nne, __popc(mine));
  base = __shfl_sync(mine, base, leader);
  const int p = base + __popc(mine & ((1u << lane) - 1u));
#else
  const int p = MN_ATOMIC_ADD(cold ? &sm.nins : &sm.nne, 1);
#endif
  if (cold) { sm.ins_mp[p] = mp; sm.ins_lo[p] = lo; sm.ins_hi[p] = hi; sm.ins_rec[p] = rec; }
  else { sm.ne_mp[p] = mp; sm.ne_lo[p] = lo; sm.ne_hi[p] = hi; sm.ne_rec[p] = rec; }
}
// Store priority `mp` on a record whose stored priority was `mp_old` and whose guard state was `g_old`
// (MN_G_NONE: no queued entry): queue an entry unless one popping before-or-at the new position is known to be
// queued -- that is, unless the priority did not rise (a re-keyed record: fell; its old entries carry the old
// key, which orders equal priorities).  Returns the new guard state.
MN_D uint32_t mn_store_priority(MnSm& sm, float mp, float mp_old, uint32_t g_old, bool rekey, int lo, int hi, int rec) {
  if (!(mp >= 0.0f)) return g_old == MN_G_NONE ? MN_G_NONE : MN_G_ABOVE;  // dormant; a queued entry stays queued
  const bool push = g_old == MN_G_NONE || (rekey ? mp >= mp_old : mp > mp_old);
  if (push) {
    mn_push_entry(sm, mp, lo, hi, rec);
    return MN_G_EXACT;
  }
  return (g_old == MN_G_EXACT && mp == mp_old) ? MN_G_EXACT : MN_G_ABOVE;
}

// ---- hash bucket helpers on buckets already in registers ------------------------------------------
MN_D void mn_load_bucket4(const MnImage& im, uint32_t b, uint32_t* out) { mn_load_bucket(im, b, out); }
// global slot index of value `val` in the two loaded buckets (-1: not there)
MN_D int mn_bucket_find_val(const MnHashPos& p, const uint32_t* bk /*16*/, uint32_t val) {
  for (int s = 0; s < 8; s++) if (bk[s] == val) return (int)(p.b1 * 8 + s);
  for (int s = 0; s < 8; s++) if (bk[8 + s] == val) return (int)(p.b2 * 8 + s);
  return -1;
}
// the (tiny) overflow area of the hash lives in shared memory: no memory access on a lookup miss
MN_D int mn_ovf_find(const MnSm& sm, int n, int lo, int hi) {
  for (int i = 0; i < n; i++)
    if (sm.ovf_lo[i] == lo && sm.ovf_hi[i] == hi) return sm.ovf_rec[i];
  return -1;
}
MN_D void mn_ovf_erase(MnSm& sm, int rec) {
  const int n = sm.hash_ovf_n < MN_OVF ? sm.hash_ovf_n : MN_OVF;
  for (int i = 0; i < n; i++)
    if (sm.ovf_rec[i] == rec && sm.ovf_lo[i] >= 0) { sm.ovf_lo[i] = -1; sm.ovf_hi[i] = -1; return; }
}
// insert with a hint: `islot` was free when the buckets were read.  Returns the slot taken (-1: overflow)
MN_D int mn_hash_insert_hint(const MnImage& im, MnSm& sm, int lo, int hi, int rec, int islot) {
  MnHashPos p = mn_hash_pos(im.hash_nbuckets, lo, hi);
  uint32_t val = (p.fp << MN_HASH_FP_SHIFT) | (uint32_t)(rec + 1);
  uint32_t* bk1 = im.hash + (size_t)p.b1 * 8;
  uint32_t* bk2 = im.hash + (size_t)p.b2 * 8;
  if (islot >= 0 && MN_ATOMIC_CAS(&im.hash[islot], 0u, val) == 0u) return islot;
  for (int w = 0; w < 2; w++) {
    uint32_t* bk = w ? bk2 : bk1;
    for (int s = 0; s < 8; s++)
      if (bk[s] == 0 && MN_ATOMIC_CAS(&bk[s], 0u, val) == 0u) return (int)(bk - im.hash) + s;
  }
  int i = MN_ATOMIC_ADD(&sm.hash_ovf_n, 1);
  if (i < MN_OVF) { sm.ovf_lo[i] = lo; sm.ovf_hi[i] = hi; sm.ovf_rec[i] = rec; }
  else mn_fail(im, MN_ERR_HASH_FULL);
  return -1;
}

// 64-bit pop-order key of a queue entry for the cascade test of the accept pass: larger = pops earlier.
// [mp bits + 1 : 32][~(top 32 bits of the 48-bit tie) : 32].  Entries that differ only in the low 16 tie
// bits get the same key; the test treats equal keys as "pops first", which is the safe side.
MN_D unsigned long long mn_pop_key(float mp, int lo, int hi) {
  return (((unsigned long long)mn_f2u(mp) + 1ull) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)(mn_tie(lo, hi) >> 16));
}

// ---- plan the pairs [p0, p1) (record t of candidate j's absorbed object), cc:650-707 -----------
// Dependent round trips: record t -> {neighbour object, new-key buckets, old-key buckets} ->
// partner record (+ the neighbour's class vector when classes differ).
MN_D void mn_plan_pairs(const MnImage& im, MnSm& sm, const MnMergeArgs& A, const float* c_clp, int p0, int p1) {
  const int novf = sm.hash_ovf_n < MN_OVF ? sm.hash_ovf_n : MN_OVF;
  MN_FOR(ii, p1 - p0) {
    int i = p0 + ii;
    int j = sm.w.pr.cand[i], t = sm.w.pr.t[i];
    int a = sm.c_surv[j], b = sm.c_abs[j];
    const uint4 tw = mn_load_rec(im, t);  // key, hash position, guard state, oml, mp
    const int2 lh = make_int2(mn_rec_lo(tw.x), mn_rec_hi(tw.y));
    int x = lh.x == b ? lh.y : lh.x;
    if ((lh.x != b && lh.y != b) || x == a || x < 0) {  // cc:665-673
      mn_fail(im, MN_ERR_INTERNAL);
      sm.w.pr.x[i] = a; sm.w.pr.u[i] = -1; sm.w.pr.mp[i] = -1.0f; sm.w.pr.eslot[i] = -1; sm.w.pr.islot[i] = -1;
      sm.w.pr.lo[i] = 0; sm.w.pr.hi[i] = 0; sm.w.pr.oml[i] = 0; sm.w.pr.g[i] = 0; sm.w.pr.mpold[i] = -1.0f;
      continue;
    }
    int nlo = a < x ? a : x, nhi = a < x ? x : a;
    MnHashPos pn = mn_hash_pos(im.hash_nbuckets, nlo, nhi);
    uint32_t bn[16];
    mn_load_bucket4(im, pn.b1, bn); mn_load_bucket4(im, pn.b2, bn + 8);
    uint4 ox = im.obj[x];
    // where t sits in the hash under its old key (cc:680 erases it)
    const int eslot = mn_slot_of_hs(mn_hash_pos(im.hash_nbuckets, lh.x, lh.y), mn_rec_hs(tw.x));
    // partner record u = (a, x) if the survivor is already linked to x (cc:685-686)
    int u = -1, islot = -1, f1 = 0, f2 = 0;
    for (int s = 0; s < 8; s++) { f1 += bn[s] == 0; f2 += bn[8 + s] == 0; }
    int c0 = -1, c1 = -1, srest = 16;
    for (int s = 0; s < 16; s++) {
      uint32_t hv = bn[s];
      if (hv != 0 && (hv >> MN_HASH_FP_SHIFT) == pn.fp) {
        int r = (int)(hv & ((1u << MN_HASH_FP_SHIFT) - 1)) - 1;
        if (c0 < 0) c0 = r; else if (c1 < 0) c1 = r; else { srest = s; break; }
      }
    }
    uint4 uw = make_uint4(MN_REC_DEAD, 0u, 0u, 0u);
    if (c0 >= 0) {  // fingerprint matches: key and values of (up to) two candidates in one round trip
      const uint4 w0 = mn_load_rec(im, c0);
      uint4 w1 = make_uint4(MN_REC_DEAD, 0u, 0u, 0u);
      if (c1 >= 0) w1 = mn_load_rec(im, c1);
      if (mn_rec_lo(w0.x) == nlo && mn_rec_hi(w0.y) == nhi) { u = c0; uw = w0; }
      else if (c1 >= 0 && mn_rec_lo(w1.x) == nlo && mn_rec_hi(w1.y) == nhi) { u = c1; uw = w1; }
    }
    for (int s = srest; s < 16 && u < 0; s++) {  // (a third fingerprint match: practically never)
      uint32_t hv = bn[s];
      if (hv != 0 && (hv >> MN_HASH_FP_SHIFT) == pn.fp) {
        int r = (int)(hv & ((1u << MN_HASH_FP_SHIFT) - 1)) - 1;
        const uint4 w2 = mn_load_rec(im, r);
        if (mn_rec_lo(w2.x) == nlo && mn_rec_hi(w2.y) == nhi) { u = r; uw = w2; }
      }
    }
    if (u < 0 && novf > 0) {
      u = mn_ovf_find(sm, novf, nlo, nhi);
      if (u >= 0) uw = mn_load_rec(im, u);
    }
    if (u < 0) {  // free slot for the re-keyed record: the emptier bucket first
      int first = (f1 >= f2) ? 0 : 8;
      for (int s = 0; s < 8 && islot < 0; s++) if (bn[first + s] == 0) islot = (int)((first ? pn.b2 : pn.b1) * 8 + s);
      for (int s = 0; s < 8 && islot < 0; s++) if (bn[(8 - first) + s] == 0) islot = (int)((first ? pn.b1 : pn.b2) * 8 + s);
    }
    // the record that carries the pair from now on: u (cc:690-692: that += this) or the re-keyed t
    float oml = mn_u2f(tw.z), mpold = mn_u2f(tw.w);
    uint32_t g = mn_rec_guard(tw.x);
    if (u >= 0) {
      oml = MN_FADD(mn_u2f(uw.z), oml);
      mpold = mn_u2f(uw.w); g = mn_rec_guard(uw.x) | (mn_rec_hs(uw.x) << 2);  // (u keeps its hash position)
    }
    int nx = mn_nc_npix(ox.x), cx = mn_nc_cls(ox.x);
    const float* clpa = c_clp + (size_t)(j * 3 + 2) * A.C;
    const float* clpx = im.clp + (size_t)x * A.C;
    float mp;
    if (a < x) mp = mn_priority(oml, A.omf, A.mlb, A.C, sm.c_na[j], sm.c_merged[j], clpa, nx, cx, clpx, nullptr);
    else mp = mn_priority(oml, A.omf, A.mlb, A.C, nx, cx, clpx, sm.c_na[j], sm.c_merged[j], clpa, nullptr);
    sm.w.pr.x[i] = x; sm.w.pr.u[i] = u; sm.w.pr.oml[i] = oml; sm.w.pr.g[i] = (int)g; sm.w.pr.mpold[i] = mpold;
    sm.w.pr.mp[i] = mp; sm.w.pr.lo[i] = nlo; sm.w.pr.hi[i] = nhi;
    sm.w.pr.eslot[i] = eslot; sm.w.pr.islot[i] = islot;
    if (mp >= 0.0f) MN_ATOMIC_MAX(&sm.c_maxnew[j], mn_pop_key(mp, nlo, nhi));
  }
}

// ---- commit the pairs [p0, p1) of accepted candidates -------------------------------------------
MN_D void mn_commit_pairs(const MnImage& im, MnSm& sm, const MnMergeArgs& A, int p0, int p1) {
  MN_FOR(ii, p1 - p0) {
    int i = p0 + ii;
    int j = sm.w.pr.cand[i];
    if (!sm.c_accept[j]) continue;
    int t = sm.w.pr.t[i], u = sm.w.pr.u[i];
    if (sm.w.pr.eslot[i] >= 0) im.hash[sm.w.pr.eslot[i]] = 0;  // cc:680
    else mn_ovf_erase(sm, t);
    const float mp = sm.w.pr.mp[i];
    const int lo = sm.w.pr.lo[i], hi = sm.w.pr.hi[i];
    const uint32_t gw = (uint32_t)sm.w.pr.g[i];
    if (u >= 0) {  // cc:690-698: fold t into u, t dies
      const uint32_t g = mn_store_priority(sm, mp, sm.w.pr.mpold[i], gw & 3u, false, lo, hi, u);
      MN_WATCH(u, "fold-into mp %.9g mp_old %.9g g_old %u g_new %u key %d %d (t=%d)", mp, sm.w.pr.mpold[i], gw & 3u, g, lo, hi, t);
      MN_WATCH(t, "folded (dies) into %d", u);
      mn_store_rec(im, u, make_uint4(mn_rec_pack_x(lo, gw >> 2, g), (uint32_t)hi, mn_f2u(sm.w.pr.oml[i]), mn_f2u(mp)));
      MN_REC(im, t).x = MN_REC_DEAD;  // cc:694
      mn_clear_live(im, A, t);
    } else {  // cc:659-664,677,700-706: t is re-keyed to (survivor, x)
      const uint32_t g = mn_store_priority(sm, mp, sm.w.pr.mpold[i], gw & 3u, true, lo, hi, t);
      MN_WATCH(t, "adopt mp %.9g mp_old %.9g g_old %u g_new %u key %d %d", mp, sm.w.pr.mpold[i], gw & 3u, g, lo, hi);
      *reinterpret_cast<uint2*>(&MN_REC(im, t)) = make_uint2(mn_rec_pack_x(lo, MN_HS_NONE, g), (uint32_t)hi);  // (the hash verifies keys through the record)
      const int hs = mn_hash_insert_hint(im, sm, lo, hi, t, sm.w.pr.islot[i]);
      mn_store_rec(im, t, make_uint4(mn_rec_pack_x(lo, mn_hs_of_slot(mn_hash_pos(im.hash_nbuckets, lo, hi), hs), g), (uint32_t)hi,
                                     mn_f2u(sm.w.pr.oml[i]), mn_f2u(mp)));
    }
  }
}

// cc:635-647 for accepted merge candidate j: object-level part of Merge (one thread)
MN_D void mn_commit_merge_object(const MnImage& im, MnSm& sm, const MnMergeArgs& A, int j) {
  int a = sm.c_surv[j], b = sm.c_abs[j], r = sm.c_rec[j];
  // cc:635-642 (obj.w, the live mask of pixel a, is updated concurrently by atomics: leave it alone)
  im.obj[a].x = mn_pack_nc(sm.c_na[j], sm.c_merged[j]);
  im.obj[a].z = (uint32_t)sm.c_newptr[j];
  im.parent[b] = a;                                        // cc:724-725
  if (sm.c_eslot[j] >= 0) im.hash[sm.c_eslot[j]] = 0;      // cc:645-647
  else mn_ovf_erase(sm, r);
  MN_REC(im, r).x = MN_REC_DEAD;                           // cc:726
  mn_clear_live(im, A, r);
}

// ------------------------------------------------------------------------------------------------
// Merge the new hot-bound entries (ne_*) into hot after dropping the first `cut` hot entries.
// Output goes to the other hot buffer; what does not fit spills to the insert buffer and the bound
// moves up to the last kept entry.
// lead_sync / trail_sync: whether the caller needs the barrier before / after (the round loop has its own on
// both sides: the commit phase ends with one, the loop starts with one)
MN_D void mn_hot_update(const MnImage& im, MnSm& sm, int cut, bool lead_sync, bool trail_sync) {
  if (lead_sync) MN_SYNC();
  const int m = sm.nne;
  const int nh = sm.nhot - cut;
  if (m == 0 && cut == 0) return;
  if (m > 0) {
    if (m <= MN_RANK_MAX) {  // rank by brute force; the ranked entry goes straight to its place (ne is read-only here)
      MN_FOR(i, m) {
        const float mp = sm.ne_mp[i]; const int lo = sm.ne_lo[i], hi = sm.ne_hi[i];
        const unsigned long long ti = mn_tie(lo, hi);
        int rk = 0;
        for (int q = 0; q < m; q++) {
          const float qmp = sm.ne_mp[q];
          // q pops before i: higher priority, or the same priority and the smaller tie (ties of equal keys: index)
          bool before = qmp > mp;
          if (qmp == mp) {
            const unsigned long long tq = mn_tie(sm.ne_lo[q], sm.ne_hi[q]);
            before = tq < ti || (tq == ti && q < i);
          }
          rk += before ? 1 : 0;
        }
        sm.sb_mp[rk] = mp; sm.sb_lo[rk] = lo; sm.sb_hi[rk] = hi; sm.sb_rec[rk] = sm.ne_rec[i];
      }
      MN_SYNC();
    } else {
      int n2 = mn_pow2_ge(m);
      MN_FOR(i, n2) {
        if (i < m) { sm.sb_mp[i] = sm.ne_mp[i]; sm.sb_lo[i] = sm.ne_lo[i]; sm.sb_hi[i] = sm.ne_hi[i]; sm.sb_rec[i] = sm.ne_rec[i]; }
        else { sm.sb_mp[i] = MN_NEG_INF; sm.sb_lo[i] = INT_MAX; sm.sb_hi[i] = INT_MAX; sm.sb_rec[i] = -1; }
      }
      MN_SYNC();
      mn_sort_sb(sm, n2);
    }
  }
  const int src = sm.hsel, dst = sm.hsel ^ 1;
  // old hot element i moves to i + (#new entries popping strictly before it)
  MN_FOR(i, nh) {
    int s = cut + i;
    float mp = sm.hot_mp[src][s]; int lo = sm.hot_lo[src][s], hi = sm.hot_hi[src][s], rec = sm.hot_rec[src][s];
    int a = 0, bnd = m;
    while (a < bnd) { int mid = (a + bnd) >> 1; if (mn_before(sm.sb_mp[mid], sm.sb_lo[mid], sm.sb_hi[mid], mp, lo, hi)) a = mid + 1; else bnd = mid; }
    int p = i + a;
    if (p < MN_HC) { sm.hot_mp[dst][p] = mp; sm.hot_lo[dst][p] = lo; sm.hot_hi[dst][p] = hi; sm.hot_rec[dst][p] = rec; }
    else { int q = MN_ATOMIC_ADD(&sm.nins, 1); sm.ins_mp[q] = mp; sm.ins_lo[q] = lo; sm.ins_hi[q] = hi; sm.ins_rec[q] = rec; }
  }
  // new element q moves to q + (#hot entries popping before-or-equal it)
  MN_FOR(q, m) {
    float mp = sm.sb_mp[q]; int lo = sm.sb_lo[q], hi = sm.sb_hi[q], rec = sm.sb_rec[q];
    int a = 0, bnd = nh;
    while (a < bnd) { int mid = (a + bnd) >> 1; int s = cut + mid; if (mn_before(mp, lo, hi, sm.hot_mp[src][s], sm.hot_lo[src][s], sm.hot_hi[src][s])) bnd = mid; else a = mid + 1; }
    int p = q + a;
    if (p < MN_HC) { sm.hot_mp[dst][p] = mp; sm.hot_lo[dst][p] = lo; sm.hot_hi[dst][p] = hi; sm.hot_rec[dst][p] = rec; }
    else { int z = MN_ATOMIC_ADD(&sm.nins, 1); sm.ins_mp[z] = mp; sm.ins_lo[z] = lo; sm.ins_hi[z] = hi; sm.ins_rec[z] = rec; }
  }
  MN_SYNC();
  if (MN_T0) {
    const int total = nh + m;
    sm.hsel = dst;
    if (total > MN_HC) {
      sm.nhot = MN_HC;
      sm.b_mp = sm.hot_mp[dst][MN_HC - 1]; sm.b_lo = sm.hot_lo[dst][MN_HC - 1]; sm.b_hi = sm.hot_hi[dst][MN_HC - 1];
      sm.cold_empty = 0;
    } else {
      sm.nhot = total;
    }
    sm.nne = 0;
  }
  if (trail_sync) MN_SYNC();
}

// ------------------------------------------------------------------------------------------------
// round phases

// Stage everything the classification of the first n hot entries needs: one dependent round trip.
// c_clp holds per candidate [lo vector | hi vector | merged vector], C floats each.
MN_D void mn_stage_candidates(const MnImage& im, MnSm& sm, const MnMergeArgs& A, float* c_clp, int n) {
  MN_FOR(w, n * 4) {
    const int j = w >> 2, role = w & 3;
    const int rec = HOT_REC(j), lo = HOT_LO(j), hi = HOT_HI(j);
    if (role == 0) {  // the record: 16 bytes
      const uint4 rw = mn_load_rec(im, rec);
      sm.c_recw[j] = rw;
      sm.c_lh[j] = make_int2(mn_rec_lo(rw.x), mn_rec_hi(rw.y));
      sm.c_rec[j] = rec; sm.c_key[j] = HOT_MP(j); sm.c_lo[j] = lo; sm.c_hi[j] = hi;
      sm.c_kind[j] = MN_K_DROP; sm.c_npairs[j] = 0; sm.c_pfill[j] = 0; sm.c_maxnew[j] = 0; sm.c_conflict[j] = 0;
      sm.c_nb[j] = 0; sm.c_accept[j] = 0; sm.c_cpbase[j] = -1;
    } else if (role == 1) sm.c_obj[j][0] = im.obj[lo];
    else if (role == 2) sm.c_obj[j][1] = im.obj[hi];
    else {
      // an earlier window entry of the same record: bit 0; bit 1: also the same priority and key
      const float mp = HOT_MP(j);
      int d = 0;
      for (int i = 0; i < j; i++) {
        if (HOT_REC(i) != rec) continue;  // (one shared-memory read per earlier entry; the rest only on a hit)
        d |= 1 | ((HOT_MP(i) == mp && HOT_LO(i) == lo && HOT_HI(i) == hi) ? 2 : 0);
      }
      sm.c_dup[j] = d;
    }
  }
  const int C = A.C;
  const int total = n * 2 * C;
  for (int w0 = MN_TID; w0 < total; w0 += 4 * MN_NT) {  // loads first, then stores: one round trip
    float v[4]; int dst[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int w = w0 + q * MN_NT;
      dst[q] = -1;
      if (w < total) {
        const int j = w / (2 * C), rem = w - j * 2 * C;
        const int side = rem / C, c = rem - side * C;
        const int o = side ? HOT_HI(j) : HOT_LO(j);
        v[q] = im.clp[(size_t)o * C + c];
        dst[q] = (j * 3 + side) * C + c;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) if (dst[q] >= 0) c_clp[dst[q]] = v[q];
  }
}

// Classify candidate j from the staged data (cc:554-561).  One thread per candidate.
MN_D void mn_classify(const MnImage& im, MnSm& sm, const MnMergeArgs& A, const float* c_clp, int j) {
  const float mp = sm.c_key[j]; const int lo = sm.c_lo[j], hi = sm.c_hi[j];
  const int2 lh = sm.c_lh[j];
  const uint4 rw = sm.c_recw[j];
  int st = mn_entry_state(mp, lo, hi, rw);
  // a second entry of a record in one window: an exact duplicate (same priority and key) of an exact entry, or --
  // the earlier one then is a guard too, and guards the record alone -- any later entry of a guarded record
  if (st != MN_K_DROP && (sm.c_dup[j] & (st == MN_K_RESTORE ? 2 : 1))) st = MN_K_DROP;
  if (st != MN_K_RESTORE) {
    sm.c_kind[j] = st;
    if (st == MN_K_REQUEUE) sm.c_maxnew[j] = mn_pop_key(mn_u2f(rw.w), lh.x, lh.y);
    // a guard touches its record, whose endpoints may have moved since the entry was queued
    if (st != MN_K_DROP) { sm.c_lo[j] = lh.x; sm.c_hi[j] = lh.y; }
    return;
  }
  const uint4 o1 = sm.c_obj[j][0], o2 = sm.c_obj[j][1];
  const int n1 = mn_nc_npix(o1.x), n2 = mn_nc_npix(o2.x), cl1 = mn_nc_cls(o1.x), cl2 = mn_nc_cls(o2.x);
  int merged;
  const float nmp = mn_priority(mn_u2f(rw.z), A.omf, A.mlb, A.C, n1, cl1, c_clp + (size_t)(j * 3) * A.C, n2, cl2,
                                c_clp + (size_t)(j * 3 + 1) * A.C, &merged);  // cc:560
  sm.c_newmp[j] = nmp;
  sm.c_merged[j] = merged;
  if (nmp == mp) {  // cc:561-562 -> Merge; cc:612-616: the larger object survives, lower id on ties
    sm.c_kind[j] = MN_K_MERGE;
    const bool swap = n1 < n2;
    sm.c_surv[j] = swap ? hi : lo; sm.c_abs[j] = swap ? lo : hi;
    sm.c_na[j] = n1 + n2; sm.c_nb[j] = swap ? n1 : n2;
    const uint4 oa = swap ? o2 : o1, ob = swap ? o1 : o2;
    sm.c_ptra[j] = (int)oa.z; sm.c_ptrb[j] = (int)ob.z;
    // where the record sits in the hash (cc:645-647 erases it)
    sm.c_eslot[j] = mn_slot_of_hs(mn_hash_pos(im.hash_nbuckets, lh.x, lh.y), mn_rec_hs(rw.x));
  } else {  // cc:563-565
    sm.c_kind[j] = MN_K_RESTORE;
    if (nmp >= 0.0f) sm.c_maxnew[j] = mn_pop_key(nmp, lo, hi);
  }
}

// merged class vector of the merging candidates [0, n) (cc:640: this += other)
MN_D void mn_stage_merged_clp(MnSm& sm, const MnMergeArgs& A, float* c_clp, int n) {
  const int C = A.C;
  MN_FOR(w, n * C) {
    const int j = w / C, c = w - j * C;
    if (sm.c_kind[j] == MN_K_MERGE) {
      const int sa = sm.c_surv[j] == sm.c_lo[j] ? 0 : 1;
      c_clp[(size_t)(j * 3 + 2) * C + c] = MN_FADD(c_clp[(size_t)(j * 3 + sa) * C + c], c_clp[(size_t)(j * 3 + 1 - sa) * C + c]);
    }
  }
}

// pixels [i0, i0 + n) of candidate j's absorbed object -> pw slots [s0, s0 + n): pixel, live mask
// without the merging record's own bits, pair count (one more dependent round trip: array -> masks)
MN_D void mn_load_pixels(const MnImage& im, MnSm& sm, const MnMergeArgs& A, int j, int i0, int s0, int n) {
  MN_FOR(k, n) {
    const int b = sm.c_abs[j];
    int pix; uint32_t m;
    if (sm.c_nb[j] == 1) { pix = b; m = sm.c_obj[j][b == sm.c_lo[j] ? 0 : 1].w; }
    else { pix = im.pix_pool[sm.c_ptrb[j] + i0 + k]; m = im.obj[pix].w; }
    m &= ~mn_own_bits(A, pix, sm.c_rec[j]);
    const int cnt = MN_POPC(m);
    sm.pw_pix[s0 + k] = pix; sm.pw_mask[s0 + k] = m; sm.pw_cand[s0 + k] = j; sm.pw_cnt[s0 + k] = cnt;
    if (cnt) MN_ATOMIC_ADD(&sm.c_npairs[j], cnt);
  }
}

// survivor pixel arrays of the accepted merges: copy the arrays that move, append the absorbed pixels
MN_D void mn_commit_pixels(const MnImage& im, MnSm& sm, int npw) {
  const int ncp = sm.ncp;
  const int ncopy = sm.cp_base[ncp];
  MN_FOR(i, ncopy) {
    int e = 0;
    while (e + 1 < ncp && sm.cp_base[e + 1] <= i) e++;
    const int j = sm.cp_list[e], idx = i - sm.cp_base[e];
    const int n_a = sm.c_na[j] - sm.c_nb[j];
    const int src = n_a == 1 ? sm.c_surv[j] : im.pix_pool[sm.c_ptra[j] + idx];
    im.pix_pool[sm.c_newptr[j] + idx] = src;
  }
  MN_FOR(i, npw) {
    const int j = sm.pw_cand[i];
    if (!sm.c_accept[j]) continue;
    const int n_a = sm.c_na[j] - sm.c_nb[j];
    im.pix_pool[sm.c_newptr[j] + n_a + (i - sm.c_pwbase[j])] = sm.pw_pix[i];
  }
}

// pixel array of the survivor of merge candidate j (thread 0): keep it when the merged object still
// fits its capacity class, otherwise take a fresh array and schedule the copy
MN_D void mn_alloc_pixels(const MnImage& im, MnSm& sm, int j) {
  const int na = sm.c_na[j], n_a = na - sm.c_nb[j];
  const int capn = mn_pix_cap(na), capo = mn_pix_cap(n_a);
  if (capn != capo) {
    int ptr = sm.pix_bump;
    if (ptr + capn > sm.pix_hi) { sm.need_gc = 1; ptr = 0; }
    else sm.pix_bump = ptr + capn;
    sm.c_newptr[j] = ptr;
    sm.cp_list[sm.ncp] = j; sm.cp_base[sm.ncp + 1] = sm.cp_base[sm.ncp] + n_a; sm.ncp++;
  } else {
    sm.c_newptr[j] = sm.c_ptra[j];
  }
}

// Solo mode: the first live candidate f is a merge whose absorbed object does not fit the work
// lists.  It is the next event of the sequential order whatever else is queued, so it is planned
// and committed in slices, alone.
MN_D void mn_solo_merge(const MnImage& im, MnSm& sm, const MnMergeArgs& A, float* c_clp, int f) {
  MN_SYNC();
  const int C = A.C;
  // candidate f becomes candidate 0 of a one-member round
  if (f != 0) {
    MN_FOR(c, 3 * C) c_clp[c] = c_clp[(size_t)f * 3 * C + c];
    if (MN_T0) {
      sm.c_rec[0] = sm.c_rec[f]; sm.c_key[0] = sm.c_key[f]; sm.c_lo[0] = sm.c_lo[f]; sm.c_hi[0] = sm.c_hi[f];
      sm.c_newmp[0] = sm.c_newmp[f]; sm.c_merged[0] = sm.c_merged[f];
      sm.c_surv[0] = sm.c_surv[f]; sm.c_abs[0] = sm.c_abs[f]; sm.c_na[0] = sm.c_na[f]; sm.c_nb[0] = sm.c_nb[f];
      sm.c_ptra[0] = sm.c_ptra[f]; sm.c_ptrb[0] = sm.c_ptrb[f]; sm.c_eslot[0] = sm.c_eslot[f];
    }
  }
  MN_SYNC();
  if (MN_T0) {
    sm.c_kind[0] = MN_K_MERGE; sm.c_accept[0] = 1; sm.c_maxnew[0] = 0; sm.c_pbase[0] = 0; sm.c_pwbase[0] = 0;
    // entries of the consumed prefix that forget their guard
    sm.st_solo++; sm.st_events++; sm.st_merges++; sm.st_rounds++;
    sm.gc_tried = 0;
    sm.nne = 0;
    sm.ncp = 0; sm.cp_base[0] = 0;
    mn_alloc_pixels(im, sm, 0);
  }
  MN_SYNC();
  mn_hot_update(im, sm, f + 1);  // drop the consumed prefix (stale entries and f itself)
  {  // move the survivor's pixel array if it must grow
    const int ncopy = sm.cp_base[sm.ncp];
    const int n_a = sm.c_na[0] - sm.c_nb[0];
    MN_FOR(i, ncopy) im.pix_pool[sm.c_newptr[0] + i] = n_a == 1 ? sm.c_surv[0] : im.pix_pool[sm.c_ptra[0] + i];
  }
  MN_SYNC();
  const int nb = sm.c_nb[0];
  const int n_a = sm.c_na[0] - nb;
  for (int base = 0; base < nb; base += MN_PW) {
    const int cnt = nb - base < MN_PW ? nb - base : MN_PW;
    if (MN_T0) sm.c_npairs[0] = 0;
    MN_SYNC();
    mn_load_pixels(im, sm, A, 0, base, 0, cnt);
    MN_SYNC();
    MN_FOR(i, cnt) im.pix_pool[sm.c_newptr[0] + n_a + base + i] = sm.pw_pix[i];
    mn_exclusive_scan(sm, sm.pw_cnt, sm.pw_off, cnt);
    // sub-slices of at most MN_WL pairs
    int w0 = 0;
    while (w0 < cnt) {
      MN_SYNC();
      if (MN_T0) {
        const int o0 = sm.pw_off[w0];
        int a = w0, bnd = cnt;  // last w with off[w] + cnt[w] - o0 <= MN_WL
        while (a < bnd) { int mid = (a + bnd) >> 1; if (sm.pw_off[mid] + sm.pw_cnt[mid] - o0 <= MN_WL) a = mid + 1; else bnd = mid; }
        sm.tmp2 = a; sm.tmp3 = (a < cnt ? sm.pw_off[a] : sm.pw_off[cnt - 1] + sm.pw_cnt[cnt - 1]) - o0;
      }
      MN_SYNC();
      const int w1 = sm.tmp2, npr = sm.tmp3;
      if (w1 == w0) { mn_fail(im, MN_ERR_INTERNAL); return; }
      const int o0 = sm.pw_off[w0];
      MN_FOR(ii, w1 - w0) {
        const int i = w0 + ii;
        uint32_t m = sm.pw_mask[i];
        int slot = sm.pw_off[i] - o0;
        while (m) {
          int bit = 31 - MN_CLZ(m);
          m &= ~(1u << bit);
          sm.w.pr.cand[slot] = 0; sm.w.pr.t[slot] = mn_rec_of_bit(A, sm.pw_pix[i], bit);
          slot++;
        }
      }
      MN_SYNC();
      mn_plan_pairs(im, sm, A, c_clp, 0, npr);
      MN_SYNC();
      mn_commit_pairs(im, sm, A, 0, npr);
      MN_SYNC();
      if (MN_T0) sm.st_pairs += npr;
      mn_hot_update(im, sm, 0);
      if (sm.nins > MN_IC - MN_NE - 64) mn_flush_ins(im, sm);
      if (sm.failed) return;
      w0 = w1;
    }
    MN_SYNC();
  }
  MN_SYNC();
  if (MN_T0) mn_commit_merge_object(im, sm, A, 0);
  MN_FOR(c, C) im.clp[(size_t)sm.c_surv[0] * C + c] = c_clp[(size_t)2 * C + c];
  MN_SYNC();
}

MN_D void mn_pix_gc(const MnImage& im, MnSm& sm, const MnMergeArgs& A);
// The solo merge of candidate f allocates its survivor array up front: collect the pool first when it
// would not fit.  Returns true when the round has to be planned again (or the image failed).
MN_D bool mn_solo_needs_gc(const MnImage& im, MnSm& sm, const MnMergeArgs& A, int f) {
  MN_SYNC();
  if (MN_T0) {
    const int na = sm.c_na[f], n_a = na - sm.c_nb[f];
    const int capn = mn_pix_cap(na), capo = mn_pix_cap(n_a);
    sm.tmp3 = (capn != capo && sm.pix_bump + capn > sm.pix_hi) ? 1 : 0;
  }
  MN_SYNC();
  if (!sm.tmp3) return false;
  if (sm.gc_tried) { if (MN_T0) mn_fail(im, MN_ERR_PL_POOL); MN_SYNC(); return true; }
  mn_pix_gc(im, sm, A);
  if (MN_T0) sm.gc_tried = 1;
  MN_SYNC();
  return true;
}

// consume the non-event entries of the window prefix [0, n): forget guards of dormant records
MN_D void mn_consume_unguard(const MnImage& im, MnSm& sm, int n) {
  MN_FOR(j, n) {
    if (sm.c_kind[j] == MN_K_UNGUARD) MN_REC(im, sm.c_rec[j]).x = mn_rec_with_guard(sm.c_recw[j].x, MN_G_NONE);
  }
}

// ---- the three scans over the (<= 32) candidates of a round --------------------------------------
// On the device they run on warp 0 with one lane per candidate (ballots and shuffles: a thread-0
// loop over shared memory costs ~30 cycles per access); on the host they are plain loops.
#if defined(__CUDA_ARCH__)
// Warp 0 runs the scans with MN_NH candidates per lane: lane l holds the candidates l, l + 32, ...; masks over
// the window are 64 bits wide.
#define MN_NH ((MN_H + 31) / 32)
#if MN_NH > 2
#error "the accept / capacity passes hold at most two candidates per lane (MN_H <= 64)"
#endif
typedef unsigned long long mn_wmask;
MN_D mn_wmask mn_ballot_w(const bool* p) {
  mn_wmask m = 0;
#pragma unroll
  for (int h = 0; h < MN_NH; h++) m |= (mn_wmask)__ballot_sync(0xffffffffu, p[h]) << (32 * h);
  return m;
}
MN_D int mn_ffs_w(mn_wmask m) { return __ffsll((long long)m) - 1; }  // -1: empty
MN_D mn_wmask mn_below_w(int n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }
MN_D int mn_wscan_incl(int v, int lane) {
  for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += o; }
  return v;
}
#endif

// Capacity cut of the window [0, n0) by pixels (vals = c_nb) or by pairs (vals = c_npairs) of its
// merging members: members that do not fit `cap` wait for a later round.  When the first event does not
// fit on its own it is merged alone, in slices ("solo") -- unless a requeue ahead of it may put
// another record in front, in which case this round only consumes that prefix.
// Writes: base[j] per merging member, sm.ncand, sm.solo, and returns nothing; with `pixels` also the
// list of merging members (m_list, m_base), sm.first, sm.npw; otherwise sm.npr.
MN_D void mn_pass_capacity(MnSm& sm, int n0, const int* vals, int* base, int cap, bool pixels) {
#if defined(__CUDA_ARCH__)
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  int k[MN_NH], v[MN_NH], incl[MN_NH];
  bool isM[MN_NH], ev[MN_NH], rq[MN_NH], ov[MN_NH];
  int carry = 0;
#pragma unroll
  for (int h = 0; h < MN_NH; h++) {
    const int j = lane + 32 * h;
    k[h] = j < n0 ? sm.c_kind[j] : MN_K_DROP;
    isM[h] = k[h] == MN_K_MERGE;
    v[h] = isM[h] ? vals[j] : 0;
    incl[h] = mn_wscan_incl(v[h], lane) + carry;
    carry = __shfl_sync(0xffffffffu, incl[h], 31);
    ev[h] = k[h] == MN_K_RESTORE || k[h] == MN_K_MERGE;
    rq[h] = k[h] == MN_K_REQUEUE;
    ov[h] = isM[h] && incl[h] > cap;
  }
  const mn_wmask evm = mn_ballot_w(ev), rqm = mn_ballot_w(rq), mm = mn_ballot_w(isM), om = mn_ballot_w(ov);
  const int f = pixels ? mn_ffs_w(evm) : sm.first;
  int n = n0, solo = 0;
  if (om) {
    const int jo = mn_ffs_w(om);
    n = jo;
    const int nrq = __popcll(rqm & mn_below_w(jo));
    solo = (jo == f && nrq == 0) ? 1 : 0;
  }
  int tot = 0;  // pixels / pairs of the members below the cut
#pragma unroll
  for (int h = 0; h < MN_NH; h++) {
    const int src = n - 1 - 32 * h;
    const int t = __shfl_sync(0xffffffffu, incl[h], src & 31);
    if (src >= 0 && src < 32) tot = t;
  }
#pragma unroll
  for (int h = 0; h < MN_NH; h++) {
    const int j = lane + 32 * h;
    if (isM[h] && j < n) {
      base[j] = incl[h] - v[h];
      if (pixels) { const int mi = __popcll(mm & mn_below_w(j)); sm.m_list[mi] = j; sm.m_base[mi] = incl[h] - v[h]; }
    }
  }
  if (lane == 0) {
    if (pixels) {
      const int nm = __popcll(mm & mn_below_w(n));
      sm.m_base[nm] = tot; sm.nm = nm; sm.first = f; sm.npw = tot; sm.npr = 0;
    } else {
      if (n < n0) sm.st_cut_cap++;
      sm.npr = tot;
    }
    sm.ncand = n; sm.solo = solo;
  }
#else
  int tot = 0, n = 0, f = pixels ? -1 : sm.first, solo = 0, nm = 0, nrq = 0;
  for (int j = 0; j < n0; j++) {
    const int k = sm.c_kind[j];
    if (pixels && (k == MN_K_RESTORE || k == MN_K_MERGE) && f < 0) f = j;
    if (k == MN_K_REQUEUE) nrq++;
    if (k == MN_K_MERGE) {
      if (tot + vals[j] > cap) { if (j == f && nrq == 0) solo = 1; break; }
      base[j] = tot;
      if (pixels) { sm.m_list[nm] = j; sm.m_base[nm] = tot; nm++; }
      tot += vals[j];
    }
    n = j + 1;
  }
  if (pixels) { sm.m_base[nm] = tot; sm.nm = nm; sm.first = f; sm.npw = tot; sm.npr = 0; }
  else { if (n < n0) sm.st_cut_cap++; sm.npr = tot; }
  sm.ncand = n; sm.solo = solo;
#endif
}

// Accept the longest provably sequential prefix of the window [0, ncand) (rules (a) and (b) of the
// header) and give every accepted merge the pixel array of its survivor: the old one when the merged
// object still fits its capacity class, else a fresh one (cp_list / cp_base schedule the copies).
MN_D void mn_pass_accept(const MnImage& im, MnSm& sm, int ncand, int npr) {
#if defined(__CUDA_ARCH__)
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  int k[MN_NH];
  bool member[MN_NH], event[MN_NH], cc[MN_NH], cconf[MN_NH];
  unsigned long long exmax[MN_NH];
  unsigned long long carry = 0;  // running maximum of the pop keys stored by the members of the earlier halves
#pragma unroll
  for (int h = 0; h < MN_NH; h++) {
    const int j = lane + 32 * h;
    k[h] = j < ncand ? sm.c_kind[j] : MN_K_DROP;
    member[h] = k[h] != MN_K_DROP;
    event[h] = k[h] == MN_K_RESTORE || k[h] == MN_K_MERGE;
  }
  const mn_wmask memm = mn_ballot_w(member);
#pragma unroll
  for (int h = 0; h < MN_NH; h++) {
    const int j = lane + 32 * h;
    unsigned long long x = member[h] ? sm.c_maxnew[j] : 0ull;
    for (int d = 1; d < 32; d <<= 1) { unsigned long long o = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d && o > x) x = o; }
    unsigned long long e = __shfl_up_sync(0xffffffffu, x, 1);  // inclusive -> exclusive
    if (lane == 0) e = 0ull;
    exmax[h] = e > carry ? e : carry;
    const unsigned long long tot = __shfl_sync(0xffffffffu, x, 31);
    if (tot > carry) carry = tot;
    const bool before = (memm & mn_below_w(j)) != 0;
    cconf[h] = member[h] && before && sm.c_conflict[j] != 0;
    // rule (b), in the full pop order (mp, then tie): an entry stored by an earlier member pops before this event
    const bool ccasc = member[h] && before && !cconf[h] && event[h] && exmax[h] != 0 &&
                       exmax[h] >= mn_pop_key(sm.c_key[j], sm.c_lo[j], sm.c_hi[j]);
    cc[h] = cconf[h] || ccasc;
  }
  const mn_wmask cm = mn_ballot_w(cc);
  const int cut = cm ? mn_ffs_w(cm) : ncand;
  bool acc[MN_NH], am[MN_NH], ar[MN_NH], aq[MN_NH], au[MN_NH], ad[MN_NH], cf[MN_NH], nd[MN_NH];
  int need[MN_NH], n_a[MN_NH], need_incl[MN_NH], cp_incl[MN_NH];
  int need_carry = 0, cp_carry = 0;
#pragma unroll
  for (int h = 0; h < MN_NH; h++) {
    const int j = lane + 32 * h;
    acc[h] = member[h] && j < cut;
    if (j < ncand) sm.c_accept[j] = acc[h] ? 1 : 0;
    am[h] = acc[h] && k[h] == MN_K_MERGE; ar[h] = acc[h] && k[h] == MN_K_RESTORE;
    aq[h] = acc[h] && k[h] == MN_K_REQUEUE; au[h] = acc[h] && k[h] == MN_K_UNGUARD;
    ad[h] = j < cut && j < ncand && k[h] == MN_K_DROP;
    cf[h] = cconf[h] && j == cut;
    // pixel arrays
    need[h] = 0; n_a[h] = 0;
    if (am[h]) {
      const int na = sm.c_na[j];
      n_a[h] = na - sm.c_nb[j];
      const int capn = mn_pix_cap(na), capo = mn_pix_cap(n_a[h]);
      if (capn != capo) need[h] = capn;
    }
    nd[h] = need[h] != 0;
    need_incl[h] = mn_wscan_incl(need[h], lane) + need_carry;
    cp_incl[h] = mn_wscan_incl(need[h] ? n_a[h] : 0, lane) + cp_carry;
    need_carry = __shfl_sync(0xffffffffu, need_incl[h], 31);
    cp_carry = __shfl_sync(0xffffffffu, cp_incl[h], 31);
  }
  const mn_wmask accm = memm & mn_below_w(cut);
  const mn_wmask mergem = mn_ballot_w(am), restm = mn_ballot_w(ar), reqm = mn_ballot_w(aq), ungm = mn_ballot_w(au),
                 dropm = mn_ballot_w(ad), confcut = mn_ballot_w(cf), needm = mn_ballot_w(nd);
  const int total_need = need_carry, total_cp = cp_carry;
  const int bump = sm.pix_bump;
  const bool fits = bump + total_need <= sm.pix_hi;
#pragma unroll
  for (int h = 0; h < MN_NH; h++) {
    const int j = lane + 32 * h;
    if (am[h]) {
      if (need[h]) {
        const int ci = __popcll(needm & mn_below_w(j));
        sm.c_newptr[j] = fits ? bump + need_incl[h] - need[h] : 0;
        sm.cp_list[ci] = j; sm.cp_base[ci] = cp_incl[h] - n_a[h];
      } else {
        sm.c_newptr[j] = sm.c_ptra[j];
      }
    }
  }
  if (lane == 0) {
    const int ncp = __popcll(needm);
    sm.ncp = ncp; sm.cp_base[ncp] = total_cp;
    sm.cutpos = cut; sm.nacc = __popcll(accm);
    if (fits) {
      sm.pix_bump = bump + total_need;
      sm.st_merges += __popcll(mergem); sm.st_restores += __popcll(restm); sm.st_requeues += __popcll(reqm);
      sm.st_events += __popcll(mergem) + __popcll(restm);
      sm.st_invalid += __popcll(ungm) + __popcll(dropm);
      if (cm) { if (confcut) sm.st_cut_conf++; else sm.st_cut_casc++; }
      sm.st_rounds++; sm.st_pairs += npr;
    } else {
      sm.need_gc = 1;  // nothing of this round is committed: collect the pixel pool and plan it again
    }
  }
#else
  unsigned long long runmax = 0;  // pop key of the earliest-popping entry stored by an accepted member
  int cut = ncand, nacc = 0;
#ifdef MN_EMUL_STATS
  int runmax_j = -1;
#endif
  sm.ncp = 0; sm.cp_base[0] = 0;
  const long long s0 = sm.st_invalid, s1 = sm.st_cut_conf, s2 = sm.st_cut_casc, s3 = sm.st_merges, s4 = sm.st_events,
                  s5 = sm.st_restores, s6 = sm.st_requeues;
  const int bump0 = sm.pix_bump;
  for (int j = 0; j < ncand; j++) {
    const int k = sm.c_kind[j];
    if (k == MN_K_DROP) { sm.st_invalid++; continue; }
    const bool event = (k == MN_K_RESTORE || k == MN_K_MERGE);
    if (sm.c_conflict[j] && nacc > 0) { cut = j; sm.st_cut_conf++; break; }
    if (event && runmax != 0 && runmax >= mn_pop_key(sm.c_key[j], sm.c_lo[j], sm.c_hi[j]) && nacc > 0) {
      cut = j; sm.st_cut_casc++;
#ifdef MN_EMUL_STATS
      MN_EMUL_STATS(sm, runmax_j, j);
#endif
      break;
    }
    sm.c_accept[j] = 1;
    nacc++;
#ifdef MN_EMUL_STATS
    if (sm.c_maxnew[j] > runmax) runmax_j = j;
#endif
    if (sm.c_maxnew[j] > runmax) runmax = sm.c_maxnew[j];
    if (k == MN_K_MERGE) { sm.st_merges++; sm.st_events++; mn_alloc_pixels(im, sm, j); }
    else if (k == MN_K_RESTORE) { sm.st_restores++; sm.st_events++; }
    else if (k == MN_K_REQUEUE) sm.st_requeues++;
    else sm.st_invalid++;
  }
  sm.cutpos = cut; sm.nacc = nacc;
  if (sm.need_gc) {  // nothing of this round is committed: collect the pixel pool and plan it again
    sm.st_invalid = s0; sm.st_cut_conf = s1; sm.st_cut_casc = s2; sm.st_merges = s3; sm.st_events = s4;
    sm.st_restores = s5; sm.st_requeues = s6; sm.pix_bump = bump0;
  } else {
    sm.st_rounds++; sm.st_pairs += npr;
  }
#endif
}

// Semi-space collection of the pixel pool.  Every merge that outgrows its survivor's array abandons the
// old one (and the absorbed object's), so the bump pointer runs far ahead of what is alive; but the live
// arrays never exceed 2 N ints (capacity < 2 * npix, and the pixels of all objects add up to N).  When a
// round's arrays do not fit the active half, the live arrays are copied, packed, to the other half: one
// scan over the objects (MN_GCL per step, one per thread), small arrays copied by their thread, large
// ones by the whole block.  Costs ~1 ms and happens a handful of times per image.  A half holds 4 N ints: the packed
// live arrays (<= 2 N) plus the new survivor arrays of one round (<= 2 N: its merges touch disjoint objects).
MN_D void mn_pix_gc(const MnImage& im, MnSm& sm, const MnMergeArgs& A) {
  MN_SYNC();
  const int half = im.pix_cap / 2;
  const int new_lo = sm.pix_hi > half ? 0 : half;
  if (MN_T0) { sm.tmp0 = new_lo; sm.st_gcs++; }
  MN_SYNC();
  const int STEP = 8 * MN_GCL;  // objects per step: 8 per thread slot, parents fetched together
  for (int base = 0; base < A.N; base += STEP) {
    if (MN_T0) sm.gc_n = 0;
    MN_SYNC();
    MN_FOR(i, MN_GCL) {
      int par[8];
#pragma unroll
      for (int q = 0; q < 8; q++) { const int p = base + q * MN_GCL + i; par[q] = p < A.N ? im.parent[p] : -1; }
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int p = base + q * MN_GCL + i;
        if (par[q] != p) continue;
        const uint4 o = im.obj[p];
        const int n = mn_nc_npix(o.x);
        if (n <= 1) continue;
        const int dst = MN_ATOMIC_ADD(&sm.tmp0, mn_pix_cap(n));
        int e = -1;
        if (n > 32) { e = MN_ATOMIC_ADD(&sm.gc_n, 1); if (e >= MN_GCL) e = -1; }
        if (e >= 0) { sm.gc_list[e] = p; sm.gc_dst[e] = dst; sm.gc_src[e] = (int)o.z; }
        else for (int t = 0; t < n; t++) im.pix_pool[dst + t] = im.pix_pool[(int)o.z + t];  // small (or list full)
        im.obj[p].z = (uint32_t)dst;
      }
    }
    MN_SYNC();
    const int nl = sm.gc_n < MN_GCL ? sm.gc_n : MN_GCL;
    for (int e = 0; e < nl; e++) {  // large arrays: the whole block copies each
      const int p = sm.gc_list[e], dst = sm.gc_dst[e], src = sm.gc_src[e];
      const int n = mn_nc_npix(im.obj[p].x);
      MN_FOR(q, n) im.pix_pool[dst + q] = im.pix_pool[src + q];
    }
    MN_SYNC();
  }
  if (MN_T0) {
    sm.pix_bump = sm.tmp0;
    sm.pix_hi = new_lo + half;
    if (sm.pix_bump > sm.pix_hi) mn_fail(im, MN_ERR_PL_POOL);  // (cannot happen: live <= 2 N <= half)
  }
  MN_SYNC();
}

// The scheduler for one image.  c_clp: MN_H * 3 * C floats of shared memory.
MN_D void mn_merge_image(const MnImage& im, MnSm& sm, const MnMergeArgs& A, float* c_clp) {
  // ---- init ----
  if (MN_T0) {
    sm.hsel = 0; sm.nhot = 0; sm.nins = 0; sm.nne = 0; sm.cold_empty = 0;
    sm.b_mp = 0; sm.b_lo = 0; sm.b_hi = 0; sm.path_n = 0; sm.failed = 0;
    sm.pix_bump = 0; sm.pix_hi = im.pix_cap / 2; sm.need_gc = 0; sm.gc_tried = 0; sm.st_gcs = 0; sm.hash_ovf_n = im.ctl->hash_ovf_n;
    sm.qc_free_top = 0; sm.qc_bump = im.qc_low_n; sm.qc_low_bump = 0; sm.qc_low_avail = 0; sm.tn_bump = MN_NROOTS; sm.tree_entries = 0; sm.peak_entries = 0; sm.peak_chunks = 0;
    sm.st_rounds = sm.st_events = sm.st_merges = sm.st_restores = sm.st_invalid = sm.st_solo = 0;
    sm.st_refills = sm.st_flushes = sm.st_splits = sm.st_pairs = sm.st_cut_conf = sm.st_cut_casc = sm.st_cut_cap = 0;
    sm.st_requeues = 0;
    for (int i = 0; i < MN_NCYC; i++) sm.cyc[i] = 0;
    sm.cyc_t0 = 0;
    // number of real (non-sentinel) initial entries: first index whose key is the sentinel
    long long E = (long long)A.N * A.K;
    long long a = 0, b = E;
    while (a < b) { long long mid = (a + b) >> 1; if (im.init_keys[mid] == ~0ull) b = mid; else a = mid + 1; }
    sm.n_init = (int)a;
    sm.static_cursor = 0;
    if (im.ctl->status != MN_OK) sm.failed = 1;
  }
  MN_SYNC();
  {  // the records record-init could not place in a bucket
    const int n = sm.hash_ovf_n;
    if (n > MN_OVF || (uint32_t)n > im.hash_ovf_cap) { if (MN_T0) mn_fail(im, MN_ERR_HASH_FULL); }
    else MN_FOR(i, n) {
      const int rec = (int)im.hash_ovf[i] - 1;
      int2 lh = make_int2(-1, -1);
      if (rec >= 0) lh = mn_rec_key(im, rec);
      sm.ovf_lo[i] = lh.x; sm.ovf_hi[i] = lh.y; sm.ovf_rec[i] = rec;
    }
  }
  MN_FOR(i, (int)((MN_NROOTS + 31) / 32)) sm.root_bits[i] = 0;
  MN_FOR(i, (int)(((MN_NROOTS + 31) / 32 + 31) / 32)) sm.root_sum[i] = 0;
  MN_SYNC();

#if defined(__CUDA_ARCH__) && defined(MN_CALIBRATE)
  {  // latency calibration (debug builds only): cyc[GC] = 64 dependent random loads by thread 0, then
     // cyc[REFILL] = 16 x (one random load per thread + barrier)
    const long long E = (long long)A.N * A.K;
    if (MN_T0) {
      long long t0 = clock64();
      unsigned idx = 12345u;
      for (int i = 0; i < 64; i++) { int2 v = mn_rec_key(im, (int)(idx % (unsigned)E)); idx = idx * 1664525u + 1013904223u + (unsigned)v.x; }
      sm.cyc[MN_CY_PAIRLIST] = clock64() - t0 + (idx == 7u ? 1 : 0);
    }
    MN_SYNC();
    long long t0 = clock64();
    unsigned idx = 777u * (MN_TID + 1);
    for (int i = 0; i < 16; i++) {
      int2 v = mn_rec_key(im, (int)(idx % (unsigned)E));
      idx = idx * 1664525u + 1013904223u + (unsigned)v.x;
      sm.ct_obj[MN_TID % MN_CT] = (int)idx;
      MN_SYNC();
    }
    if (MN_T0) sm.cyc[MN_CY_REFILL] = clock64() - t0;
    MN_SYNC();
  }
#endif
  for (long long round = 0;; round++) {
    MN_SYNC();
    if (sm.failed) break;
    if (A.max_rounds > 0 && round >= A.max_rounds) { if (MN_T0) mn_fail(im, MN_ERR_LIMIT); break; }
#ifdef MN_EMUL_ROUND_HOOK
    MN_EMUL_ROUND_HOOK(im, sm, A, round);  // (host test builds only)
#endif
    MN_TIC();
    if (sm.nins > MN_IC - MN_NE - 64) { mn_flush_ins(im, sm); MN_TOC(MN_CY_FLUSH); }
    if (sm.nhot == 0) {
      mn_refill(im, sm, A);
      MN_TOC(MN_CY_REFILL);
      if (sm.failed) break;
      if (sm.nhot == 0) break;  // queue empty: cc:542
    }
    const int ncand0 = sm.nhot < A.H ? sm.nhot : A.H;
    // ---- phase 1: stage (one round trip), classify ----
    mn_stage_candidates(im, sm, A, c_clp, ncand0);
    MN_FOR(i, MN_CT) { sm.ct_obj[i] = -1; sm.ct_w[i] = INT_MAX; sm.ct_r[i] = INT_MAX; }
    MN_SYNC();
    MN_TOC(MN_CY_SEL_STAGE);
#if defined(__CUDA_ARCH__)
    // classify (one thread per candidate) and the capacity cut by pixels (warp 0) run back to back: no block
    // barrier between them
    if (threadIdx.x < 32 * MN_NH) {
      if ((int)threadIdx.x < ncand0) mn_classify(im, sm, A, c_clp, (int)threadIdx.x);
#if MN_NH == 1
      __syncwarp();
#else
      asm volatile("bar.sync 1, %0;" :: "n"(32 * MN_NH) : "memory");  // the classifying warps only
#endif
    }
#else
    MN_FOR(j, ncand0) mn_classify(im, sm, A, c_clp, j);
#endif
    // ---- phase 2: capacity cut by pixels ----
    mn_pass_capacity(sm, ncand0, sm.c_nb, sm.c_pwbase, MN_PW, true);
    MN_SYNC();
    if (sm.solo) {
      const int f = sm.first;
      if (mn_solo_needs_gc(im, sm, A, f)) continue;
      mn_consume_unguard(im, sm, f);
      mn_stage_merged_clp(sm, A, c_clp, ncand0);
      mn_solo_merge(im, sm, A, c_clp, f);
      MN_TOC(MN_CY_SOLO);
      continue;
    }
    MN_TOC(MN_CY_SEL_CLASS);
    // ---- phase 3: pixels of the absorbed objects, their live masks, and -- optimistically -- the pair list
    //      itself: slots come from one counter (a warp takes its slots with one atomic); only when the pairs
    //      of the window exceed the work list does the capacity cut by pairs (below) redo the list ----
    {
      const int npw = sm.npw, nm = sm.nm;
#if defined(__CUDA_ARCH__)
      const int lane = threadIdx.x & 31;
      for (int i0 = (int)(threadIdx.x & ~31u); i0 < npw; i0 += MN_NT) {
        const int i = i0 + lane;
#else
      for (int i = 0; i < npw; i++) {
#endif
        int cnt = 0, pix = 0, j = 0; uint32_t m = 0;
        if (i < npw) {
          int a = 0, bnd = nm;  // last e with m_base[e] <= i
          while (a + 1 < bnd) { int mid = (a + bnd) >> 1; if (sm.m_base[mid] <= i) a = mid; else bnd = mid; }
          j = sm.m_list[a];
          const int idx = i - sm.m_base[a];
          const int b = sm.c_abs[j];
          if (sm.c_nb[j] == 1) { pix = b; m = sm.c_obj[j][b == sm.c_lo[j] ? 0 : 1].w; }
          else { pix = im.pix_pool[sm.c_ptrb[j] + idx]; m = im.obj[pix].w; }
          m &= ~mn_own_bits(A, pix, sm.c_rec[j]);
          cnt = MN_POPC(m);
          sm.pw_pix[i] = pix; sm.pw_mask[i] = m; sm.pw_cand[i] = j; sm.pw_cnt[i] = cnt;
          if (cnt) MN_ATOMIC_ADD(&sm.c_npairs[j], cnt);
        }
#if defined(__CUDA_ARCH__)
        const int incl = mn_wscan_incl(cnt, lane);
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int base = 0;
        if (lane == 31 && total) base = atomicAdd(&sm.npr, total);
        base = __shfl_sync(0xffffffffu, base, 31);
        int slot = base + incl - cnt;
#else
        int slot = sm.npr; sm.npr += cnt;
#endif
        if (cnt && slot + cnt <= MN_WL) {
          while (m) {
            int bit = 31 - MN_CLZ(m);
            m &= ~(1u << bit);
            sm.w.pr.cand[slot] = j; sm.w.pr.t[slot] = mn_rec_of_bit(A, pix, bit);
            slot++;
          }
        }
      }
      mn_stage_merged_clp(sm, A, c_clp, ncand0);
    }
    MN_SYNC();
    MN_TOC(MN_CY_SEL_PIX);
    if (sm.npr > MN_WL) {  // (rare: 1-2 % of the rounds) the window's pairs do not fit: cut it, list the pairs again
      MN_SYNC();
      mn_pass_capacity(sm, sm.ncand, sm.c_npairs, sm.c_pbase, MN_WL, false);  // capacity cut by pairs
      MN_SYNC();
      if (sm.solo) {
        const int f = sm.first;
        if (mn_solo_needs_gc(im, sm, A, f)) continue;
        mn_consume_unguard(im, sm, f);
        mn_solo_merge(im, sm, A, c_clp, f);
        MN_TOC(MN_CY_SOLO);
        continue;
      }
      const int ncand_c = sm.ncand, npw = sm.npw;
      MN_FOR(i, npw) {
        const int j = sm.pw_cand[i];
        if (j >= ncand_c) continue;
        uint32_t m = sm.pw_mask[i];
        const int cnt = sm.pw_cnt[i];
        if (!cnt) continue;
        int slot = sm.c_pbase[j] + MN_ATOMIC_ADD(&sm.c_pfill[j], cnt);
        while (m) {
          int bit = 31 - MN_CLZ(m);
          m &= ~(1u << bit);
          sm.w.pr.cand[slot] = j; sm.w.pr.t[slot] = mn_rec_of_bit(A, sm.pw_pix[i], bit);
          slot++;
        }
      }
      MN_SYNC();
    }
    const int ncand = sm.ncand, npr = sm.npr, npw = sm.npw;
    MN_SYNC();
    MN_TOC(MN_CY_PAIRLIST);
    // ---- phase 4: plan ----
    mn_plan_pairs(im, sm, A, c_clp, 0, npr);
    MN_SYNC();
    MN_TOC(MN_CY_PLAN);
    // ---- phase 5: footprints into the conflict table ----
    MN_FOR(j, ncand) {
      const int k = sm.c_kind[j];
      if (k == MN_K_DROP) continue;
      int s1 = mn_ct_slot(sm, sm.c_lo[j]), s2 = mn_ct_slot(sm, sm.c_hi[j]);
      if (s1 < 0 || s2 < 0) { sm.c_conflict[j] = 1; continue; }
      if (k == MN_K_MERGE) { MN_ATOMIC_MIN(&sm.ct_w[s1], j); MN_ATOMIC_MIN(&sm.ct_w[s2], j); }
      else { MN_ATOMIC_MIN(&sm.ct_r[s1], j); MN_ATOMIC_MIN(&sm.ct_r[s2], j); }
    }
    MN_FOR(i, npr) {
      int s = mn_ct_slot(sm, sm.w.pr.x[i]);
      if (s < 0) sm.c_conflict[sm.w.pr.cand[i]] = 1; else MN_ATOMIC_MIN(&sm.ct_r[s], sm.w.pr.cand[i]);
    }
    MN_SYNC();
    MN_FOR(j, ncand) {
      const int k = sm.c_kind[j];
      if (k == MN_K_DROP) continue;
      int s1 = mn_ct_find(sm, sm.c_lo[j]), s2 = mn_ct_find(sm, sm.c_hi[j]);
      bool cf = false;
      if (s1 >= 0) cf = cf || sm.ct_w[s1] < j || (k == MN_K_MERGE && sm.ct_r[s1] < j);
      if (s2 >= 0) cf = cf || sm.ct_w[s2] < j || (k == MN_K_MERGE && sm.ct_r[s2] < j);
      if (cf) sm.c_conflict[j] = 1;
    }
    MN_FOR(i, npr) {
      int s = mn_ct_find(sm, sm.w.pr.x[i]);
      if (s >= 0 && sm.ct_w[s] < sm.w.pr.cand[i]) sm.c_conflict[sm.w.pr.cand[i]] = 1;
    }
    MN_SYNC();
    MN_TOC(MN_CY_CONFLICT);
    // ---- phase 6: accept the longest provably sequential prefix ----
    mn_pass_accept(im, sm, ncand, npr);
    MN_SYNC();
    if (sm.need_gc) {
      if (sm.gc_tried) { if (MN_T0) mn_fail(im, MN_ERR_PL_POOL); continue; }  // even a packed pool is too small
      mn_pix_gc(im, sm, A);
      if (MN_T0) { sm.need_gc = 0; sm.gc_tried = 1; }
      continue;  // nothing was committed: the same window is staged and planned again
    }
    if (MN_T0) sm.gc_tried = 0;
#ifdef MN_EMUL_TRACE
    fprintf(stderr, "round: ncand %d nacc %d cut %d nhot %d nins %d npr %d cold_empty %d key0 %.9g\n", ncand, sm.nacc, sm.cutpos, sm.nhot, sm.nins, npr, sm.cold_empty, sm.c_key[0]);
#endif
    MN_TOC(MN_CY_ACCEPT);
    // ---- phase 7: commit ----
    MN_FOR(j, ncand) {
      if (!sm.c_accept[j]) continue;
      const int k = sm.c_kind[j];
      const int rec = sm.c_rec[j];
      MN_WATCH(rec, "commit cand j=%d kind %d key %.9g (%d,%d) mp %.9g guard %u newmp %.9g", j, k, sm.c_key[j], sm.c_lo[j], sm.c_hi[j], mn_u2f(sm.c_recw[j].w), mn_rec_guard(sm.c_recw[j].x), sm.c_newmp[j]);
      const uint4 rw = sm.c_recw[j];
      if (k == MN_K_RESTORE) {  // cc:563-565: the consumed entry was the record's only expected one
        const float nmp = sm.c_newmp[j];
        const uint32_t g = mn_store_priority(sm, nmp, mn_u2f(rw.w), MN_G_NONE, false, sm.c_lo[j], sm.c_hi[j], rec);
        mn_store_rec(im, rec, make_uint4(mn_rec_with_guard(rw.x, g), rw.y, rw.z, mn_f2u(nmp)));  // (no other accepted member touches it)
      } else if (k == MN_K_MERGE) {
        mn_commit_merge_object(im, sm, A, j);
      } else if (k == MN_K_REQUEUE) {  // the guard is consumed: the exact entry takes its place
        const uint32_t g = mn_store_priority(sm, mn_u2f(rw.w), mn_u2f(rw.w), MN_G_NONE, false, sm.c_lh[j].x, sm.c_lh[j].y, rec);
        MN_REC(im, rec).x = mn_rec_with_guard(rw.x, g);
      } else if (k == MN_K_UNGUARD) {
        MN_REC(im, rec).x = mn_rec_with_guard(rw.x, MN_G_NONE);
      }
    }
    MN_FOR(i, ncand * A.C) {
      int j = i / A.C, c = i % A.C;
      if (sm.c_accept[j] && sm.c_kind[j] == MN_K_MERGE) im.clp[(size_t)sm.c_surv[j] * A.C + c] = c_clp[(size_t)(j * 3 + 2) * A.C + c];
    }
    mn_commit_pixels(im, sm, npw);
    mn_commit_pairs(im, sm, A, 0, npr);
    MN_SYNC();
    MN_TOC(MN_CY_COMMIT);
    // ---- phase 8: queue maintenance ----
    mn_hot_update(im, sm, sm.cutpos, false, false);  // (barriers: the one above, and the one the loop starts with)
    MN_TOC(MN_CY_HOT);
  }
  MN_SYNC();
  if (MN_T0) {
    MnCtl* c = im.ctl;
    c->rounds = sm.st_rounds; c->events = sm.st_events; c->merges = sm.st_merges; c->restores = sm.st_restores;
    c->invalid_pops = sm.st_invalid; c->solo_events = sm.st_solo; c->refills = sm.st_refills;
    c->flushes = sm.st_flushes; c->splits = sm.st_splits; c->pairs = sm.st_pairs;
    c->cuts_conflict = sm.st_cut_conf; c->cuts_cascade = sm.st_cut_casc; c->cuts_capacity = sm.st_cut_cap;
    c->requeues = sm.st_requeues; c->pix_gcs = sm.st_gcs;
    c->pix_bump = sm.pix_bump; c->hash_ovf_n = sm.hash_ovf_n;
    c->qc_bump = sm.qc_bump; c->qc_free_top = sm.qc_free_top; c->tn_bump = sm.tn_bump; c->tree_entries = sm.tree_entries;
    c->static_cursor = sm.static_cursor; c->n_init = sm.n_init;
    c->peak_entries = sm.peak_entries; c->peak_chunks = sm.peak_chunks;
    for (int i = 0; i < MN_NCYC; i++) c->cyc[i] = sm.cyc[i];
  }
  MN_SYNC();
}
