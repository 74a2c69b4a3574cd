// mn_merge.cuh -- order-exact merge scheduler: one persistent CTA per image (sm_100a).
//
// Reproduces the reference's RunSegmentation / Merge loop (cc:539-573, cc:602-727) exactly, up to
// the tie order among equal priorities (deterministic here: mp desc, lo asc, hi asc).
//
// The reference's lazy heap is observationally an indexed map  record -> stored priority  holding
// the records with stored mp >= 0 (cc:554-565).  A ROUND takes the next MN_H valid entries in pop
// order, PLANS each against the round-start state (read only), then COMMITS the longest prefix
// that the sequential heap would provably execute in exactly this order with exactly these results:
//   (a) no member reads or writes an object written by an earlier member, and writes none an
//       earlier member read (objects written by a merge: both endpoints; read: every neighbour of
//       the absorbed object, whose records are rewired -- cc:650-707; a non-merging pop reads its
//       two endpoints and re-stores only its own priority -- cc:560-565);
//   (b) no earlier member creates a queue entry that would pop before a later member.
// Rule (a)+(b) was validated against the sequential restatement on the host
// (oracle/mergenet_oracle.c: mno_run_rounds_model).  Member 0 always commits, so every round makes
// progress.  All state of an image is private to its CTA: no inter-CTA communication.
//
// Queue: `hot` (shared memory, sorted) holds every entry that pops before-or-at `bound`; colder
// entries live in HBM: the sorted initial entries (init_keys, cursor) and, for entries created
// later, a lazily split radix tree of unsorted chunks keyed by (mp bits, lo, hi).  Entries are
// validated lazily against the record's stored (mp, lo, hi) when they are loaded and again when
// they are popped, which is what the reference's `merge_priority != arec->GetPriority()` test does.
//
// The file is written as SPMD phases (see mn_layout.h) and also compiles for the host, where
// tests/emul runs it single-threaded to unit-test the logic; that build is test infrastructure.
#pragma once
#include <limits.h>

#include "mn_common.h"
#include "mn_layout.h"

#define MN_H 32          // candidates per round
#define MN_PW 1024       // pixel work-list capacity
#define MN_CW 256        // pixel-chunk work-list capacity
#define MN_WL 960        // (candidate, record) pair work-list capacity
#define MN_HC 1024       // hot capacity
#define MN_IC 2048       // insert-buffer capacity
#define MN_NE 1024       // new hot-bound entries per round (>= MN_WL + MN_H, <= MN_SB)
#define MN_SB 1024       // sort buffer capacity
#define MN_LEAFCAP 512   // tree leaves larger than this are split before they are loaded
#define MN_CT 2048       // conflict-table slots (power of two)
#define MN_PLCACHE 64    // pre-popped pixel-list chunks per round
#define MN_REFILL_TARGET 384      // stop loading tree leaves once this many entries are staged
#define MN_REFILL_STATIC_MIN 128  // sort-buffer slots always left for initial entries
#define MN_NEG_INF (-3.0e38f)
#define MN_RANK_MAX 640  // up to this many entries are ordered by brute-force ranking (no barriers)
// cycle accounting buckets (thread 0, clock64)
#define MN_NCYC 10
#define MN_CY_SELECT 0   // classify + work lists
#define MN_CY_PLAN 1
#define MN_CY_ACCEPT 2
#define MN_CY_COMMIT 3
#define MN_CY_HOT 4
#define MN_CY_FLUSH 5
#define MN_CY_REFILL 6
#define MN_CY_SPLIT 7
#define MN_CY_SOLO 8
#define MN_CY_GC 9

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
  // distribute() scratch: per entry (group << 16 | index in group); per group node / count / old tail /
  // (old fill | directory base << 8); directory of freshly allocated chunks
  int ds_el[MN_SB]; int ds_node[MN_SB]; int ds_cnt[MN_SB]; int ds_tail[MN_SB]; int ds_fd[MN_SB];
  int ds_dir[MN_SB + 128];
  uint32_t root_bits[(MN_NROOTS + 31) / 32];
  uint32_t root_sum[((MN_NROOTS + 31) / 32 + 31) / 32];
  // candidates
  int c_rec[MN_H]; float c_key[MN_H]; int c_lo[MN_H]; int c_hi[MN_H]; int c_kind[MN_H];
  float c_newmp[MN_H]; int c_merged[MN_H]; int c_surv[MN_H]; int c_abs[MN_H]; int c_na[MN_H];
  float c_rsame[MN_H]; int c_npairs[MN_H]; int c_pbase[MN_H]; int c_pfill[MN_H];
  uint32_t c_maxnew[MN_H]; int c_conflict[MN_H]; int c_npix[MN_H]; int c_accept[MN_H];
  // work lists
  int cw_cand[MN_CW]; int cw_chunk[MN_CW];
  int pw_cand[MN_PW]; int pw_pix[MN_PW];
  int pr_cand[MN_WL]; int pr_t[MN_WL]; int pr_x[MN_WL]; int pr_u[MN_WL];
  float pr_oml[MN_WL]; float pr_same[MN_WL]; float pr_diff[MN_WL]; float pr_mp[MN_WL];
  int pr_lo[MN_WL]; int pr_hi[MN_WL];
  // conflict table: object -> (min writer candidate, min reader candidate)
  int ct_obj[MN_CT]; int ct_w[MN_CT]; int ct_r[MN_CT];
  int plcache[MN_PLCACHE];
  // scalars
  int nhot, nins, nne, ncw, npw, npr, ncand, nacc, cutpos, solo, done, tmp0, tmp1, tmp2, tmp3;
  int plcache_n, plcache_used;
  int ds_ngroups, ds_ndir;  // distribute() counters
  int cold_empty;  // 1: nothing outside `hot` -> every new entry goes to hot
  float b_mp; int b_lo; int b_hi;  // bound: entries popping strictly after it are cold
  int path[20]; int path_n;
  long long cyc[MN_NCYC]; long long cyc_t0;
  long long st_rounds, st_events, st_merges, st_restores, st_invalid, st_solo, st_refills,
      st_flushes, st_splits, st_pairs, st_cut_conf, st_cut_casc, st_cut_cap;
};

struct MnMergeArgs {
  int C, K, N, W;
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
#else
#define MN_CLZ(x) __builtin_clz((unsigned)(x))
#endif
#if defined(__CUDA_ARCH__)
#define MN_TIC() do { if (MN_T0) sm.cyc_t0 = clock64(); } while (0)
#define MN_TOC(k) do { if (MN_T0) { long long t__ = clock64(); sm.cyc[k] += t__ - sm.cyc_t0; sm.cyc_t0 = t__; } } while (0)
#else
#define MN_TIC() ((void)0)
#define MN_TOC(k) ((void)0)
#endif
#define MN_FOR(i, n) for (int i = MN_TID; i < (n); i += MN_NT)
#define MN_T0 (MN_TID == 0)

MN_D void mn_fail_at(const MnImage& im, int code, int line) {
  if (im.ctl->status == MN_OK) { im.ctl->status = code; im.ctl->fail_line = line; }
}
#define mn_fail(im, code) mn_fail_at((im), (code), __LINE__)

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

// ------------------------------------------------------------------------------------------------
// queue tree
MN_D int mn_root_of(float mp) {
  uint32_t b = mn_f2u(mp);
  if (b < MN_ROOT_LO_BITS) return 0;
  if (b >= MN_ROOT_HI_BITS) return MN_NROOTS - 1;
  return (int)((b - MN_ROOT_LO_BITS) >> MN_ROOT_SHIFT) + 1;
}
// 6-bit digit `level` (1-based, below the root) of the 80-bit pop-order key [~mpbits:32][lo:24][hi:24].
// Regular roots fix the top 20 key bits, so their digits start at bit 20; the two open-ended roots
// start at bit 0.  Smaller digit = pops first.
MN_D int mn_digit(int root, int level, float mp, int lo, int hi) {
  int start = (root == 0 || root == MN_NROOTS - 1) ? 0 : 20;
  int pos = start + 6 * (level - 1);  // bit offset from the top of the 80-bit key
  unsigned long long hi64 = ((unsigned long long)(~mn_f2u(mp)) << 32) | ((unsigned long long)(uint32_t)lo << 8) |
                            ((unsigned long long)(uint32_t)hi >> 16);
  unsigned long long lo16 = (unsigned long long)((uint32_t)hi & 0xFFFFu);
  int d = 0;
  for (int b = 0; b < 6; b++) {
    int p = pos + b;
    int bit;
    if (p < 64) bit = (int)((hi64 >> (63 - p)) & 1ull);
    else if (p < 80) bit = (int)((lo16 >> (79 - p)) & 1ull);
    else bit = 0;
    d = (d << 1) | bit;
  }
  return d;
}
MN_D int mn_max_level(int root) { return (root == 0 || root == MN_NROOTS - 1) ? 14 : 10; }

MN_D int mn_qc_alloc(const MnImage& im) {  // called only from phases that never free
  int t = MN_ATOMIC_SUB(&im.ctl->qc_free_top, 1);
  if (t > 0) return im.qc_free[t - 1];
  MN_ATOMIC_ADD(&im.ctl->qc_free_top, 1);
  int c = MN_ATOMIC_ADD(&im.ctl->qc_bump, 1);
  if (c >= im.qc_cap) { mn_fail(im, MN_ERR_Q_POOL); return -1; }
  return c;
}
MN_D void mn_qc_free(const MnImage& im, int c) {  // called only from phases that never allocate
  int t = MN_ATOMIC_ADD(&im.ctl->qc_free_top, 1);
  im.qc_free[t] = c;
}

// conflict-table helpers are also used as a small node -> group hash by distribute()
MN_D int mn_ct_slot(MnSm& sm, int obj);
MN_D int mn_ct_find(const MnSm& sm, int obj);

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
      sm.ct_w[s] = g; sm.ds_node[g] = sm.ct_obj[s]; sm.ds_cnt[g] = 0;
    }
  }
  MN_SYNC();
  MN_FOR(i, n) {
    int s = mn_ct_find(sm, sm.sb_node[i]);
    int g = s >= 0 ? sm.ct_w[s] : 0;
    int local = MN_ATOMIC_ADD(&sm.ds_cnt[g], 1);
    sm.ds_el[i] = (g << 16) | local;
  }
  MN_SYNC();
  const int ngroups = sm.ds_ngroups;
  MN_FOR(g, ngroups) {
    int node = sm.ds_node[g], cnt = sm.ds_cnt[g];
    int tail = im.tn_tail[node];
    int fill = tail >= 0 ? im.qc_cnt[tail] : MN_QCH;
    int room = MN_QCH - fill;
    int nnew = cnt > room ? (cnt - room + MN_QCH - 1) / MN_QCH : 0;
    int base = nnew ? MN_ATOMIC_ADD(&sm.ds_ndir, nnew) : 0;
    sm.ds_tail[g] = tail;
    sm.ds_fd[g] = fill | (base << 8);
    if (tail >= 0) im.qc_cnt[tail] = cnt > room ? MN_QCH : fill + cnt;
    int prev = tail, left = cnt - room;
    for (int j = 0; j < nnew; j++) {
      int c = mn_qc_alloc(im);
      if (base + j < MN_SB + 128) sm.ds_dir[base + j] = c;
      if (c < 0) break;
      im.qc_next[c] = -1;
      im.qc_cnt[c] = left > MN_QCH ? MN_QCH : left;
      left -= MN_QCH;
      if (prev >= 0) im.qc_next[prev] = c; else im.tn_head[node] = c;
      prev = c;
    }
    if (nnew) im.tn_tail[node] = prev;
    im.tn_cnt[node] += cnt;
  }
  MN_SYNC();
  MN_FOR(i, n) {
    int g = sm.ds_el[i] >> 16, local = sm.ds_el[i] & 0xffff;
    int fill = sm.ds_fd[g] & 0xff, base = sm.ds_fd[g] >> 8;
    int pos = fill + local, chunk, slot;
    if (pos < MN_QCH) { chunk = sm.ds_tail[g]; slot = pos; }
    else { int q = pos - MN_QCH; chunk = sm.ds_dir[base + q / MN_QCH]; slot = q % MN_QCH; }
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
  while (im.tn_child[node] >= 0) {
    MN_ATOMIC_ADD(&im.tn_cnt[node], 1);
    level++;
    node = im.tn_child[node] + mn_digit(root, level, mp, lo, hi);
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
    im.ctl->tree_entries += total;
    sm.nins = 0;
    sm.st_flushes++;
  }
  MN_SYNC();
}

// Split leaf `node` (at depth `level` under `root`): its entries move to 64 children by the next
// digit.  Entries are validated on the way (dead ones are dropped: garbage collection).
MN_D void mn_split_leaf(const MnImage& im, MnSm& sm, int root, int node, int level) {
  if (MN_T0) {
    int base = MN_ATOMIC_ADD(&im.ctl->tn_bump, MN_TREE_FANOUT);
    if (base + MN_TREE_FANOUT > im.tn_cap) { mn_fail(im, MN_ERR_TREE_POOL); base = -1; }
    sm.tmp0 = base;
    sm.tmp1 = im.tn_head[node];
    sm.tmp2 = 0;  // surviving entries
    sm.st_splits++;
  }
  MN_SYNC();
  int base = sm.tmp0;
  if (base < 0) return;
  MN_FOR(i, MN_TREE_FANOUT) {
    im.tn_head[base + i] = -1; im.tn_tail[base + i] = -1; im.tn_cnt[base + i] = 0; im.tn_child[base + i] = -1;
  }
  MN_SYNC();
  int chunk = sm.tmp1;
  int guard = 0;
  while (chunk >= 0) {
    // stage up to MN_SB entries (MN_SB / MN_QCH chunks)
    MN_SYNC();
    if (MN_T0) {
      int n = 0, c = chunk, nch = 0;
      while (c >= 0 && nch < MN_SB / MN_QCH) {
        sm.cw_chunk[nch++] = c;
        n += im.qc_cnt[c];
        c = im.qc_next[c];
      }
      sm.tmp1 = c;
      sm.tmp3 = nch;
      sm.npr = 0;
    }
    MN_SYNC();
    int nch = sm.tmp3;
    MN_FOR(i, nch * MN_QCH) {
      int c = sm.cw_chunk[i / MN_QCH], s = i % MN_QCH;
      if (s < im.qc_cnt[c]) {
        uint4 e = im.q_ent[(size_t)c * MN_QCH + s];
        int rec = (int)e.y;
        int2 lh = im.rec_lh[rec];
        float4 v = im.rec_val[rec];
        if (lh.x == (int)e.z && lh.y == (int)e.w && mn_f2u(v.w) == e.x) {
          int p = MN_ATOMIC_ADD(&sm.npr, 1);
          sm.sb_mp[p] = v.w; sm.sb_lo[p] = lh.x; sm.sb_hi[p] = lh.y; sm.sb_rec[p] = rec;
          sm.sb_node[p] = base + mn_digit(root, level + 1, v.w, lh.x, lh.y);
        }
      }
    }
    MN_SYNC();
    int n = sm.npr;
    mn_distribute(im, sm, n);
    MN_FOR(i, nch) mn_qc_free(im, sm.cw_chunk[i]);
    if (MN_T0) sm.tmp2 += n;
    MN_SYNC();
    chunk = sm.tmp1;
    if (++guard > (1 << 24)) { mn_fail(im, MN_ERR_LIMIT); break; }
  }
  MN_SYNC();
  if (MN_T0) {
    // fix the counts on the path: the node now holds only the survivors
    int removed = im.tn_cnt[node] - sm.tmp2;
    for (int i = 0; i < sm.path_n; i++) im.tn_cnt[sm.path[i]] -= removed;
    im.tn_cnt[node] = sm.tmp2;
    im.ctl->tree_entries -= removed;
    im.tn_head[node] = -1;
    im.tn_tail[node] = -1;
    im.tn_child[node] = base;
  }
  MN_SYNC();
}

// Find the first non-empty leaf in pop order; split it while it is too large.  Leaves the path
// (ancestors, root first) in sm.path and returns the leaf, -1 when the tree is empty, or -2 when the
// leaf needs a split that the caller did not allow.
MN_D int mn_top_leaf(const MnImage& im, MnSm& sm, int* root_out, bool allow_split) {
  for (int guard = 0; guard < 64; guard++) {
    MN_SYNC();
    if (MN_T0) {
      // highest non-empty root = first in pop order
      int root = -1;
      const int nsum = ((MN_NROOTS + 31) / 32 + 31) / 32;
      for (int s = nsum - 1; s >= 0 && root < 0; s--) {
        uint32_t sv = sm.root_sum[s];
        while (sv && root < 0) {
          int wb = 31 - MN_CLZ(sv);
          int w = s * 32 + wb;
          uint32_t bits = sm.root_bits[w];
          while (bits && root < 0) {
            int b = 31 - MN_CLZ(bits);
            int r = w * 32 + b;
            if (im.tn_cnt[r] > 0) root = r;
            else { bits &= ~(1u << b); sm.root_bits[w] = bits; }
          }
          if (root < 0) { sv &= ~(1u << wb); sm.root_sum[s] = sv; }
        }
      }
      sm.tmp0 = root;
      sm.path_n = 0;
      int node = root, level = 0;
      if (root >= 0) {
        while (im.tn_child[node] >= 0) {
          sm.path[sm.path_n++] = node;
          int cb = im.tn_child[node], found = -1;
          for (int d = 0; d < MN_TREE_FANOUT; d++)
            if (im.tn_cnt[cb + d] > 0) { found = cb + d; break; }
          if (found < 0) { mn_fail(im, MN_ERR_INTERNAL); break; }
          node = found;
          level++;
        }
      }
      sm.tmp1 = node;
      sm.tmp2 = level;
    }
    MN_SYNC();
    int root = sm.tmp0, node = sm.tmp1, level = sm.tmp2;
    if (root < 0) return -1;
    *root_out = root;
    if (im.tn_cnt[node] <= MN_LEAFCAP || level >= mn_max_level(root)) return node;
    if (!allow_split) return -2;  // the caller's staging buffers are in use
    MN_TOC(MN_CY_REFILL);
    mn_split_leaf(im, sm, root, node, level);
    MN_TOC(MN_CY_SPLIT);
    if (im.ctl->status != MN_OK) return -1;
  }
  mn_fail(im, MN_ERR_LIMIT);
  return -1;
}

// decode the i-th sorted initial entry
MN_D void mn_decode_init(const MnImage& im, const MnMergeArgs& A, uint64_t key, float* mp, int* lo, int* hi, int* rec) {
  uint32_t ord = (uint32_t)(key & ((1ull << MN_ORD_BITS) - 1));
  *mp = mn_u2f(~(uint32_t)(key >> MN_ORD_BITS));
  int l = (int)(ord / (uint32_t)A.K), rank = (int)(ord % (uint32_t)A.K);
  int k = A.off.k_of_rank[rank];
  int d = A.off.delta[k];
  int h = l + (d > 0 ? d : -d);
  *lo = l; *hi = h;
  int p = d > 0 ? l : h;
  *rec = p * A.K + k;
}

// exclusive prefix sum of n 0/1 flags (all threads call it); returns the total
MN_D int mn_exclusive_scan(MnSm& sm, const int* flags, int* out, int n) {
  const int nt = MN_NT, t = MN_TID;
  const int per = (n + nt - 1) / nt;
  const int b = t * per, e = b + per < n ? b + per : n;
  int ssum = 0;
  for (int i = b; i < e; i++) ssum += flags[i];
  if (t < MN_SB) sm.ds_cnt[t] = ssum;
  MN_SYNC();
  if (MN_T0) {
    int acc = 0;
    for (int k = 0; k < nt && k < MN_SB; k++) { int v = sm.ds_cnt[k]; sm.ds_cnt[k] = acc; acc += v; }
    sm.tmp3 = acc;
  }
  MN_SYNC();
  int acc = t < MN_SB ? sm.ds_cnt[t] : 0;
  for (int i = b; i < e; i++) { out[i] = acc; acc += flags[i]; }
  MN_SYNC();
  return sm.tmp3;
}

// Refill the (empty) hot buffer.  Afterwards: hot holds every entry popping before-or-at `bound`.
MN_D void mn_refill(const MnImage& im, MnSm& sm, const MnMergeArgs& A) {
  if (MN_T0) sm.st_refills++;
  for (int guard = 0; guard < (1 << 20); guard++) {
    MN_SYNC();
    mn_flush_ins(im, sm);  // (also on retries: a previous pass may have pushed leaf entries back)
    if (im.ctl->status != MN_OK) return;
    // ---- load + validate successive top leaves (each pops entirely before the next) until a
    //      useful number of entries is staged; their chunks are recycled ----
    int nleaf = 0;
    if (MN_T0) sm.npr = 0;
    MN_SYNC();
    for (int lguard = 0; lguard < 4096 && nleaf < MN_REFILL_TARGET; lguard++) {
      int root = 0;
      int leaf = mn_top_leaf(im, sm, &root, nleaf == 0);  // splitting reuses the staging buffers
      if (nleaf == 0) { MN_SYNC(); if (MN_T0) sm.npr = 0; MN_SYNC(); }  // (a split used npr)
      if (leaf < 0) break;
      if (im.ctl->status != MN_OK) return;
      if (nleaf + im.tn_cnt[leaf] > MN_SB - MN_REFILL_STATIC_MIN) {  // keep room for initial entries
        if (nleaf == 0) mn_fail(im, MN_ERR_INTERNAL);  // an unsplittable leaf larger than the buffer
        break;
      }
      if (MN_T0) sm.tmp1 = im.tn_head[leaf];
      MN_SYNC();
      int chunk = sm.tmp1;
      while (chunk >= 0) {
        MN_SYNC();
        if (MN_T0) {
          int c = chunk, nch = 0;
          while (c >= 0 && nch < MN_CW) { sm.cw_chunk[nch++] = c; c = im.qc_next[c]; }
          sm.tmp1 = c; sm.tmp3 = nch;
        }
        MN_SYNC();
        int nch = sm.tmp3;
        MN_FOR(i, nch * MN_QCH) {
          int c = sm.cw_chunk[i / MN_QCH], s = i % MN_QCH;
          if (s < im.qc_cnt[c]) {
            uint4 e = im.q_ent[(size_t)c * MN_QCH + s];
            int rec = (int)e.y;
            int2 lh = im.rec_lh[rec];
            float4 v = im.rec_val[rec];
            if (lh.x == (int)e.z && lh.y == (int)e.w && mn_f2u(v.w) == e.x) {
              int p = MN_ATOMIC_ADD(&sm.npr, 1);
              if (p < MN_SB) { sm.sb_mp[p] = v.w; sm.sb_lo[p] = lh.x; sm.sb_hi[p] = lh.y; sm.sb_rec[p] = rec; }
            }
          }
        }
        MN_SYNC();
        MN_FOR(i, nch) mn_qc_free(im, sm.cw_chunk[i]);
        MN_SYNC();
        chunk = sm.tmp1;
      }
      MN_SYNC();
      if (MN_T0) {
        int removed = im.tn_cnt[leaf];
        for (int i = 0; i < sm.path_n; i++) im.tn_cnt[sm.path[i]] -= removed;
        im.tn_cnt[leaf] = 0;
        im.tn_head[leaf] = -1;
        im.tn_tail[leaf] = -1;
        im.ctl->tree_entries -= removed;
        if (sm.npr > MN_SB) { mn_fail(im, MN_ERR_INTERNAL); sm.npr = MN_SB; }
      }
      MN_SYNC();
      nleaf = sm.npr;
    }
    // ---- the leaf's last entry in pop order bounds what may be taken from the sorted initial
    //      entries: everything else in the tree pops after it ----
    if (MN_T0) {
      int w = -1;
      for (int i = 0; i < nleaf; i++)
        if (w < 0 || mn_before(sm.sb_mp[w], sm.sb_lo[w], sm.sb_hi[w], sm.sb_mp[i], sm.sb_lo[i], sm.sb_hi[i])) w = i;
      sm.tmp0 = w;
      sm.tmp2 = 0;
    }
    MN_SYNC();
    const int w = sm.tmp0;
    const float lmp = w >= 0 ? sm.sb_mp[w] : 0.f; const int llo = w >= 0 ? sm.sb_lo[w] : 0, lhi = w >= 0 ? sm.sb_hi[w] : 0;
    const int sc = im.ctl->static_cursor, ninit = im.ctl->n_init;
    const int room = MN_HC - nleaf;
    int navail = ninit - sc; if (navail > room) navail = room;
    // count the initial entries (a prefix, they are sorted) that pop before the leaf's last entry
    MN_FOR(i, navail) {
      float mp; int lo, hi, rec;
      mn_decode_init(im, A, im.init_keys[sc + i], &mp, &lo, &hi, &rec);
      bool take = (w < 0) || mn_before(mp, lo, hi, lmp, llo, lhi);
      if (take) MN_ATOMIC_ADD(&sm.tmp2, 1);
    }
    MN_SYNC();
    const int ntake = sm.tmp2;
    // does an untaken initial entry still pop before the leaf's last entry?
    bool more_before = false;
    if (w >= 0 && ntake == navail && sc + navail < ninit) {
      float mp; int lo, hi, rec;
      mn_decode_init(im, A, im.init_keys[sc + navail], &mp, &lo, &hi, &rec);
      more_before = mn_before(mp, lo, hi, lmp, llo, lhi);
    }
    MN_FOR(i, ntake) {
      float mp; int lo, hi, rec;
      mn_decode_init(im, A, im.init_keys[sc + i], &mp, &lo, &hi, &rec);
      int2 lh = im.rec_lh[rec];
      float4 v = im.rec_val[rec];
      bool valid = (lh.x == lo && lh.y == hi && v.w == mp);
      int p = nleaf + i;
      sm.sb_mp[p] = valid ? mp : MN_NEG_INF; sm.sb_lo[p] = lo; sm.sb_hi[p] = hi; sm.sb_rec[p] = rec;
    }
    MN_SYNC();
    // the bound
    if (MN_T0) {
      if (w < 0) {  // tree empty: the last taken initial entry bounds the rest of the array
        if (ntake > 0) {
          float mp; int lo, hi, rec;
          mn_decode_init(im, A, im.init_keys[sc + ntake - 1], &mp, &lo, &hi, &rec);
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
        mn_decode_init(im, A, im.init_keys[sc + ntake - 1], &mp, &lo, &hi, &rec);
        sm.b_mp = mp; sm.b_lo = lo; sm.b_hi = hi;
        sm.cold_empty = 0;
      }
      im.ctl->static_cursor = sc + ntake;
    }
    MN_SYNC();
    if (more_before) {
      // push the leaf entries that pop after the bound back to the insert buffer
      MN_FOR(i, nleaf) {
        if (mn_before(sm.b_mp, sm.b_lo, sm.b_hi, sm.sb_mp[i], sm.sb_lo[i], sm.sb_hi[i])) {
          int p = MN_ATOMIC_ADD(&sm.nins, 1);
          sm.ins_mp[p] = sm.sb_mp[i]; sm.ins_lo[p] = sm.sb_lo[i]; sm.ins_hi[p] = sm.sb_hi[i]; sm.ins_rec[p] = sm.sb_rec[i];
          sm.sb_mp[i] = MN_NEG_INF;
        }
      }
      MN_SYNC();
    }
#ifdef MN_EMUL_TRACE
    fprintf(stderr, "refill: nleaf %d ntake %d more_before %d nins %d cold_empty %d bound %.9g %d %d sc %d ninit %d tree %d\n", nleaf, ntake, (int)more_before, sm.nins, sm.cold_empty, sm.b_mp, sm.b_lo, sm.b_hi, im.ctl->static_cursor, ninit, im.ctl->tree_entries);
#endif
    const int n = nleaf + ntake;
    if (n == 0) {
      if (MN_T0) sm.nhot = 0;
      MN_SYNC();
      return;  // nothing left anywhere (cold_empty set above)
    }
    if (nleaf <= MN_RANK_MAX) {
      // ---- fast path, no barrier-heavy sort: rank the (few) leaf entries by brute force, compact the
      //      already sorted initial entries with a scan, then merge the two runs by binary search.
      //      Duplicates of one record stay adjacent and are dropped when they are popped. ----
      if (MN_T0) sm.tmp0 = 0;
      MN_SYNC();
      MN_FOR(i, nleaf) {
        int rk = -1;
        if (sm.sb_mp[i] > MN_NEG_INF) {
          rk = 0;
          for (int q = 0; q < nleaf; q++) {
            if (q == i || !(sm.sb_mp[q] > MN_NEG_INF)) continue;
            bool qb = mn_before(sm.sb_mp[q], sm.sb_lo[q], sm.sb_hi[q], sm.sb_mp[i], sm.sb_lo[i], sm.sb_hi[i]);
            bool ib = mn_before(sm.sb_mp[i], sm.sb_lo[i], sm.sb_hi[i], sm.sb_mp[q], sm.sb_lo[q], sm.sb_hi[q]);
            if (qb || (!ib && q < i)) rk++;
          }
          MN_ATOMIC_ADD(&sm.tmp0, 1);
        }
        sm.ne_pos[i] = rk;
      }
      MN_FOR(i, ntake) sm.sb_node[i] = sm.sb_mp[nleaf + i] > MN_NEG_INF ? 1 : 0;
      MN_SYNC();
      const int nl = sm.tmp0;
      MN_FOR(i, nleaf) {
        int p = sm.ne_pos[i];
        if (p >= 0) { sm.ne_mp[p] = sm.sb_mp[i]; sm.ne_lo[p] = sm.sb_lo[i]; sm.ne_hi[p] = sm.sb_hi[i]; sm.ne_rec[p] = sm.sb_rec[i]; }
      }
      const int ns = mn_exclusive_scan(sm, sm.sb_node, sm.ds_el, ntake);
      const int cur = sm.hsel, dst = sm.hsel ^ 1;
      MN_FOR(i, ntake) {
        if (sm.sb_node[i]) {
          int p = sm.ds_el[i], q = nleaf + i;
          sm.hot_mp[cur][p] = sm.sb_mp[q]; sm.hot_lo[cur][p] = sm.sb_lo[q]; sm.hot_hi[cur][p] = sm.sb_hi[q]; sm.hot_rec[cur][p] = sm.sb_rec[q];
        }
      }
      MN_SYNC();
      MN_FOR(i, nl) {  // leaf entry i -> i + #initial entries popping before-or-equal it
        float mp = sm.ne_mp[i]; int lo = sm.ne_lo[i], hi = sm.ne_hi[i];
        int a = 0, bnd = ns;
        while (a < bnd) { int mid = (a + bnd) >> 1; if (mn_before(mp, lo, hi, sm.hot_mp[cur][mid], sm.hot_lo[cur][mid], sm.hot_hi[cur][mid])) bnd = mid; else a = mid + 1; }
        int p = i + a;
        sm.hot_mp[dst][p] = mp; sm.hot_lo[dst][p] = lo; sm.hot_hi[dst][p] = hi; sm.hot_rec[dst][p] = sm.ne_rec[i];
      }
      MN_FOR(j, ns) {  // initial entry j -> j + #leaf entries popping strictly before it
        float mp = sm.hot_mp[cur][j]; int lo = sm.hot_lo[cur][j], hi = sm.hot_hi[cur][j];
        int a = 0, bnd = nl;
        while (a < bnd) { int mid = (a + bnd) >> 1; if (mn_before(sm.ne_mp[mid], sm.ne_lo[mid], sm.ne_hi[mid], mp, lo, hi)) a = mid + 1; else bnd = mid; }
        int p = j + a;
        sm.hot_mp[dst][p] = mp; sm.hot_lo[dst][p] = lo; sm.hot_hi[dst][p] = hi; sm.hot_rec[dst][p] = sm.hot_rec[cur][j];
      }
      MN_SYNC();
      if (MN_T0) { sm.hsel = dst; sm.nhot = nl + ns; }
      MN_SYNC();
    } else {
      const int n2 = mn_pow2_ge(n);
      MN_FOR(i, n2 - n) { sm.sb_mp[n + i] = MN_NEG_INF; sm.sb_lo[n + i] = INT_MAX; sm.sb_hi[n + i] = INT_MAX; sm.sb_rec[n + i] = -1; }
      MN_SYNC();
      mn_sort_sb(sm, n2);
      // valid entries first (pads / invalid sort last); they are a prefix after the sort
      if (MN_T0) sm.tmp0 = 0;
      MN_SYNC();
      MN_FOR(i, n) {
        if (sm.sb_mp[i] > MN_NEG_INF) {
          MN_ATOMIC_ADD(&sm.tmp0, 1);
          HOT_MP(i) = sm.sb_mp[i]; HOT_LO(i) = sm.sb_lo[i]; HOT_HI(i) = sm.sb_hi[i]; HOT_REC(i) = sm.sb_rec[i];
        }
      }
      MN_SYNC();
      if (MN_T0) sm.nhot = sm.tmp0;
      MN_SYNC();
    }
    if (sm.nhot > 0 || sm.cold_empty) return;
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
  MN_ATOMIC_AND(&im.live_mask[p], ~(1u << k));
  MN_ATOMIC_AND(&im.live_mask[p + A.off.delta[k]], ~(1u << (16 + k)));
}

// queue a created entry (cc:564,697,705): hot-bound entries are staged in ne_*, colder ones go to ins
MN_D void mn_push_entry(MnSm& sm, float mp, int lo, int hi, int rec) {
  bool cold = !sm.cold_empty && mn_before(sm.b_mp, sm.b_lo, sm.b_hi, mp, lo, hi);
  if (cold) {
    int p = MN_ATOMIC_ADD(&sm.nins, 1);
    sm.ins_mp[p] = mp; sm.ins_lo[p] = lo; sm.ins_hi[p] = hi; sm.ins_rec[p] = rec;
  } else {
    int p = MN_ATOMIC_ADD(&sm.nne, 1);
    sm.ne_mp[p] = mp; sm.ne_lo[p] = lo; sm.ne_hi[p] = hi; sm.ne_rec[p] = rec;
  }
}

// ---- plan the pairs [p0, p1) (record t of candidate j's absorbed object), cc:650-707 -----------
MN_D void mn_plan_pairs(const MnImage& im, MnSm& sm, const MnMergeArgs& A, const float* c_clp, int p0, int p1) {
  MN_FOR(ii, p1 - p0) {
    int i = p0 + ii;
    int j = sm.pr_cand[i], t = sm.pr_t[i];
    int a = sm.c_surv[j], b = sm.c_abs[j];
    int2 lh = im.rec_lh[t];
    float4 v = im.rec_val[t];
    int x = lh.x == b ? lh.y : lh.x;
    if (lh.x != b && lh.y != b) mn_fail(im, MN_ERR_INTERNAL);  // cc:665-668
    if (x == a) mn_fail(im, MN_ERR_INTERNAL);                    // cc:670-673
    int nlo = a < x ? a : x, nhi = a < x ? x : a;
    int u = mn_hash_find(im, nlo, nhi);  // cc:685-686
    float oml = v.x, same = v.y, diff = v.z;
    if (u >= 0) {  // cc:690-692: that += this
      float4 uv = im.rec_val[u];
      oml = MN_FADD(uv.x, v.x); diff = MN_FADD(uv.z, v.z); same = MN_FADD(uv.y, v.y);
    }
    uint32_t xnc = im.obj_nc[x];
    int nx = mn_nc_npix(xnc), cx = mn_nc_cls(xnc);
    const float* clpa = c_clp + (size_t)j * A.C;
    const float* clpx = im.clp + (size_t)x * A.C;
    float mp;
    if (a < x) mp = mn_priority(oml, A.omf, A.mlb, A.C, sm.c_na[j], sm.c_merged[j], clpa, nx, cx, clpx, nullptr);
    else mp = mn_priority(oml, A.omf, A.mlb, A.C, nx, cx, clpx, sm.c_na[j], sm.c_merged[j], clpa, nullptr);
    sm.pr_x[i] = x; sm.pr_u[i] = u; sm.pr_oml[i] = oml; sm.pr_same[i] = same; sm.pr_diff[i] = diff;
    sm.pr_mp[i] = mp; sm.pr_lo[i] = nlo; sm.pr_hi[i] = nhi;
    if (mp >= 0.0f) MN_ATOMIC_MAX(&sm.c_maxnew[j], mn_f2u(mp) + 1u);
  }
}

// ---- commit the pairs [p0, p1) of accepted candidates -------------------------------------------
MN_D void mn_commit_pairs(const MnImage& im, MnSm& sm, const MnMergeArgs& A, int p0, int p1) {
  MN_FOR(ii, p1 - p0) {
    int i = p0 + ii;
    int j = sm.pr_cand[i];
    if (!sm.c_accept[j]) continue;
    int t = sm.pr_t[i], u = sm.pr_u[i], b = sm.c_abs[j], x = sm.pr_x[i];
    int olo = b < x ? b : x, ohi = b < x ? x : b;
    mn_hash_erase(im, olo, ohi, t);  // cc:680
    float mp = sm.pr_mp[i];
    if (u >= 0) {
      im.rec_val[u] = make_float4(sm.pr_oml[i], sm.pr_same[i], sm.pr_diff[i], mp);  // cc:690-695
      im.rec_lh[t] = make_int2(-1, -1);                                               // cc:694
      mn_clear_live(im, A, t);
      if (mp >= 0.0f) mn_push_entry(sm, mp, sm.pr_lo[i], sm.pr_hi[i], u);             // cc:696-698
    } else {
      im.rec_lh[t] = make_int2(sm.pr_lo[i], sm.pr_hi[i]);                             // cc:659-664,677
      im.rec_val[t] = make_float4(sm.pr_oml[i], sm.pr_same[i], sm.pr_diff[i], mp);    // cc:703
      mn_hash_insert(im, sm.pr_lo[i], sm.pr_hi[i], t);                                // cc:700-702
      if (mp >= 0.0f) mn_push_entry(sm, mp, sm.pr_lo[i], sm.pr_hi[i], t);             // cc:704-706
    }
  }
}

// pixel-list chunk from the per-round cache (filled by thread 0 before the commit phase)
MN_D int mn_plc_take(const MnImage& im, MnSm& sm) {
  int i = MN_ATOMIC_ADD(&sm.plcache_used, 1);
  if (i >= sm.plcache_n) { mn_fail(im, MN_ERR_PL_POOL); return -1; }
  return sm.plcache[i];
}
MN_D void mn_plc_free(const MnImage& im, int c) {
  int t = MN_ATOMIC_ADD(&im.ctl->plc_free_top, 1);
  im.plc_free[t] = c;
}
MN_D void mn_plc_cache_fill(const MnImage& im, MnSm& sm) {  // thread 0, between phases
  int keep = 0;
  for (int i = sm.plcache_used; i < sm.plcache_n; i++) sm.plcache[keep++] = sm.plcache[i];
  while (keep < MN_PLCACHE) {
    int c;
    if (im.ctl->plc_free_top > 0) c = im.plc_free[--im.ctl->plc_free_top];
    else if (im.ctl->plc_bump < im.plc_cap) c = im.ctl->plc_bump++;
    else break;
    sm.plcache[keep++] = c;
  }
  sm.plcache_n = keep;
  sm.plcache_used = 0;
}

// append pixel `pix` to object a's pixel list
MN_D void mn_pl_append(const MnImage& im, MnSm& sm, int a, int pix) {
  int tail = im.pl_tail[a];
  if (tail < 0 || im.plc_cnt[tail] >= MN_PLC) {
    int c = mn_plc_take(im, sm);
    if (c < 0) return;
    im.plc_next[c] = -1;
    im.plc_cnt[c] = 0;
    if (tail >= 0) im.plc_next[tail] = c; else im.pl_head[a] = c;
    im.pl_tail[a] = c;
    tail = c;
  }
  im.plc_pix[(size_t)tail * MN_PLC + im.plc_cnt[tail]] = pix;
  im.plc_cnt[tail]++;
}

// cc:635-647 for candidate j: object-level part of Merge (one thread)
MN_D void mn_commit_merge_object(const MnImage& im, MnSm& sm, const MnMergeArgs& A, int j) {
  int a = sm.c_surv[j], b = sm.c_abs[j], r = sm.c_rec[j];
  int nb = mn_nc_npix(im.obj_nc[b]);
  im.obj_nc[a] = mn_pack_nc(sm.c_na[j], sm.c_merged[j]);                                  // cc:635-639
  im.obj_same[a] = MN_FADD(im.obj_same[a], MN_FADD(sm.c_rsame[j], im.obj_same[b]));       // cc:641-642
  im.parent[b] = a;                                                                       // cc:724-725
  mn_hash_erase(im, sm.c_lo[j], sm.c_hi[j], r);                                           // cc:645-647
  im.rec_lh[r] = make_int2(-1, -1);                                                       // cc:726
  mn_clear_live(im, A, r);
  // pixel-set union (cc:636-639): b's root pixel and its chunks join a's list
  int bh = im.pl_head[b];
  int at = im.pl_tail[a];
  int room = at >= 0 ? MN_PLC - im.plc_cnt[at] : 0;
  if (bh < 0) {
    mn_pl_append(im, sm, a, b);
  } else if (nb <= room || (nb <= MN_PLC && nb <= 8)) {
    // small object: copy its pixels, recycle its chunks
    mn_pl_append(im, sm, a, b);
    for (int c = bh; c >= 0;) {
      int n = im.plc_cnt[c];
      for (int s = 0; s < n; s++) mn_pl_append(im, sm, a, im.plc_pix[(size_t)c * MN_PLC + s]);
      int nx = im.plc_next[c];
      mn_plc_free(im, c);
      c = nx;
    }
  } else {
    mn_pl_append(im, sm, b, b);  // b's own root pixel goes to the end of b's list
    if (at >= 0) im.plc_next[at] = im.pl_head[b]; else im.pl_head[a] = im.pl_head[b];
    im.pl_tail[a] = im.pl_tail[b];
  }
  im.pl_head[b] = -1;
  im.pl_tail[b] = -1;
}

// ------------------------------------------------------------------------------------------------
// Merge the new hot-bound entries (ne_*) into hot after dropping the first `cut` hot entries.
// Output goes to the other hot buffer; what does not fit spills to the insert buffer and the bound
// moves up to the last kept entry.
MN_D void mn_hot_update(const MnImage& im, MnSm& sm, int cut) {
  MN_SYNC();
  const int m = sm.nne;
  const int nh = sm.nhot - cut;
  if (m == 0 && cut == 0) return;
  if (m > 0) {
    if (m <= MN_RANK_MAX) {  // rank by brute force
      MN_FOR(i, m) {
        int rk = 0;
        for (int q = 0; q < m; q++) {
          if (q == i) continue;
          bool qb = mn_before(sm.ne_mp[q], sm.ne_lo[q], sm.ne_hi[q], sm.ne_mp[i], sm.ne_lo[i], sm.ne_hi[i]);
          bool ib = mn_before(sm.ne_mp[i], sm.ne_lo[i], sm.ne_hi[i], sm.ne_mp[q], sm.ne_lo[q], sm.ne_hi[q]);
          if (qb || (!ib && q < i)) rk++;
        }
        sm.ne_pos[i] = rk;
      }
      MN_SYNC();
      MN_FOR(i, m) { int p = sm.ne_pos[i]; sm.sb_mp[p] = sm.ne_mp[i]; sm.sb_lo[p] = sm.ne_lo[i]; sm.sb_hi[p] = sm.ne_hi[i]; sm.sb_rec[p] = sm.ne_rec[i]; }
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
  MN_SYNC();
}

// ------------------------------------------------------------------------------------------------
// round phases

// Phase 1: validate + classify candidate j (cc:554-561).  One thread per candidate.
MN_D void mn_classify(const MnImage& im, MnSm& sm, const MnMergeArgs& A, int j) {
  float mp = HOT_MP(j); int lo = HOT_LO(j), hi = HOT_HI(j), rec = HOT_REC(j);
  sm.c_rec[j] = rec; sm.c_key[j] = mp; sm.c_lo[j] = lo; sm.c_hi[j] = hi;
  sm.c_kind[j] = 0; sm.c_npairs[j] = 0; sm.c_pfill[j] = 0; sm.c_maxnew[j] = 0; sm.c_conflict[j] = 0;
  sm.c_npix[j] = 0; sm.c_accept[j] = 0;
  int2 lh = im.rec_lh[rec];
  float4 v = im.rec_val[rec];
  bool valid = (lh.x == lo && lh.y == hi && v.w == mp);
  if (j > 0 && HOT_REC(j - 1) == rec && HOT_MP(j - 1) == mp && HOT_LO(j - 1) == lo && HOT_HI(j - 1) == hi) valid = false;
  if (!valid) return;
  uint32_t nc1 = im.obj_nc[lo], nc2 = im.obj_nc[hi];
  int n1 = mn_nc_npix(nc1), n2 = mn_nc_npix(nc2), cl1 = mn_nc_cls(nc1), cl2 = mn_nc_cls(nc2);
  int merged;
  float nmp = mn_priority(v.x, A.omf, A.mlb, A.C, n1, cl1, im.clp + (size_t)lo * A.C, n2, cl2,
                          im.clp + (size_t)hi * A.C, &merged);  // cc:560
  sm.c_newmp[j] = nmp;
  sm.c_merged[j] = merged;
  if (nmp == mp) {  // cc:561-562 -> Merge; cc:612-616: the larger object survives, lower id on ties
    sm.c_kind[j] = 2;
    int a = lo, b = hi;
    if (n1 < n2) { a = hi; b = lo; }
    sm.c_surv[j] = a; sm.c_abs[j] = b; sm.c_na[j] = n1 + n2;
    sm.c_npix[j] = (a == lo) ? n2 : n1;
    sm.c_rsame[j] = v.y;
  } else {  // cc:563-565
    sm.c_kind[j] = 1;
    if (nmp >= 0.0f) sm.c_maxnew[j] = mn_f2u(nmp) + 1u;
  }
}

// expand the pixel-list chunks [c0..] of candidate j's absorbed object into pw (root pixel first)
MN_D void mn_expand_pixels(const MnImage& im, MnSm& sm, int ncw) {
  MN_FOR(i, ncw * MN_PLC) {
    int w = i / MN_PLC, s = i - w * MN_PLC;
    int c = sm.cw_chunk[w];
    if (s < im.plc_cnt[c]) {
      int p = MN_ATOMIC_ADD(&sm.npw, 1);
      if (p < MN_PW) { sm.pw_cand[p] = sm.cw_cand[w]; sm.pw_pix[p] = im.plc_pix[(size_t)c * MN_PLC + s]; }
    }
  }
}

// live records of pixel p other than the merging record itself
MN_D uint32_t mn_live_bits(const MnImage& im, const MnMergeArgs& A, int p, int skip_rec) {
  uint32_t m = im.live_mask[p];
  uint32_t out = m;
  while (m) {
    int bit = 31 - MN_CLZ(m);
    m &= ~(1u << bit);
    if (mn_rec_of_bit(A, p, bit) == skip_rec) out &= ~(1u << bit);
  }
  return out;
}
MN_D int mn_popc(uint32_t x) { int c = 0; while (x) { x &= x - 1; c++; } return c; }

// write the pair list entries of pixel work items [w0, w1)
MN_D void mn_fill_pairs(const MnImage& im, MnSm& sm, const MnMergeArgs& A, int w0, int w1) {
  MN_FOR(ii, w1 - w0) {
    int i = w0 + ii;
    int j = sm.pw_cand[i], p = sm.pw_pix[i];
    uint32_t m = mn_live_bits(im, A, p, sm.c_rec[j]);
    while (m) {
      int bit = 31 - MN_CLZ(m);
      m &= ~(1u << bit);
      int slot = sm.c_pbase[j] + MN_ATOMIC_ADD(&sm.c_pfill[j], 1);
      if (slot < MN_WL) { sm.pr_cand[slot] = j; sm.pr_t[slot] = mn_rec_of_bit(A, p, bit); }
    }
  }
}

// post-merge class vector of candidate j's survivor (cc:640: this += other)
MN_D void mn_stage_clp(const MnImage& im, MnSm& sm, const MnMergeArgs& A, float* c_clp, int j0, int j1) {
  MN_FOR(i, (j1 - j0) * A.C) {
    int j = j0 + i / A.C, c = i % A.C;
    if (sm.c_kind[j] == 2)
      c_clp[(size_t)j * A.C + c] = MN_FADD(im.clp[(size_t)sm.c_surv[j] * A.C + c], im.clp[(size_t)sm.c_abs[j] * A.C + c]);
  }
}

// Solo mode: the first valid candidate f is a merge whose absorbed object does not fit the work
// lists.  It is the next event of the sequential order whatever else is queued, so it is planned
// and committed in slices, alone.
MN_D void mn_solo_merge(const MnImage& im, MnSm& sm, const MnMergeArgs& A, float* c_clp, int f) {
  // candidate f becomes candidate 0 of a one-member round
  MN_SYNC();
  if (MN_T0) {
    sm.c_rec[0] = sm.c_rec[f]; sm.c_key[0] = sm.c_key[f]; sm.c_lo[0] = sm.c_lo[f]; sm.c_hi[0] = sm.c_hi[f];
    sm.c_kind[0] = 2; sm.c_newmp[0] = sm.c_newmp[f]; sm.c_merged[0] = sm.c_merged[f];
    sm.c_surv[0] = sm.c_surv[f]; sm.c_abs[0] = sm.c_abs[f]; sm.c_na[0] = sm.c_na[f];
    sm.c_rsame[0] = sm.c_rsame[f]; sm.c_npix[0] = sm.c_npix[f]; sm.c_accept[0] = 1;
    sm.c_maxnew[0] = 0; sm.c_pbase[0] = 0;
    sm.st_solo++; sm.st_events++; sm.st_merges++;
    sm.nne = 0;
  }
  MN_SYNC();
  mn_hot_update(im, sm, f + 1);  // drop the consumed prefix (invalid entries and f itself)
  mn_stage_clp(im, sm, A, c_clp, 0, 1);
  MN_SYNC();
  const int b = sm.c_abs[0];
  int chunk = im.pl_head[b];
  bool first = true;
  for (int guard = 0; guard < (1 << 26); guard++) {
    MN_SYNC();
    // ---- next slice of pixels: up to MN_PW / MN_PLC - 1 chunks (+ the root pixel once) ----
    if (MN_T0) {
      sm.npw = 0; sm.ncw = 0;
      if (first) { sm.pw_cand[0] = 0; sm.pw_pix[0] = b; sm.npw = 1; }
      int c = chunk;
      while (c >= 0 && sm.ncw < MN_PW / MN_PLC - 1) { sm.cw_cand[sm.ncw] = 0; sm.cw_chunk[sm.ncw] = c; sm.ncw++; c = im.plc_next[c]; }
      sm.tmp1 = c;
    }
    MN_SYNC();
    chunk = sm.tmp1;
    first = false;
    mn_expand_pixels(im, sm, sm.ncw);
    MN_SYNC();
    const int npw = sm.npw;
    if (npw == 0) break;
    // ---- sub-slices of at most MN_WL pairs ----
    int w0 = 0;
    while (w0 < npw) {
      MN_SYNC();
      if (MN_T0) {
        int tot = 0, w = w0;
        while (w < npw) {
          int c = mn_popc(mn_live_bits(im, A, sm.pw_pix[w], sm.c_rec[0]));
          if (tot + c > MN_WL) break;
          tot += c; w++;
        }
        sm.tmp2 = w; sm.tmp3 = tot; sm.c_pfill[0] = 0;
      }
      MN_SYNC();
      const int w1 = sm.tmp2, npr = sm.tmp3;
      mn_fill_pairs(im, sm, A, w0, w1);
      MN_SYNC();
      mn_plan_pairs(im, sm, A, c_clp, 0, npr);
      MN_SYNC();
      mn_commit_pairs(im, sm, A, 0, npr);
      MN_SYNC();
      if (MN_T0) sm.st_pairs += npr;
      mn_hot_update(im, sm, 0);
      if (sm.nins > MN_IC - MN_NE - 64) mn_flush_ins(im, sm);
      if (im.ctl->status != MN_OK) return;
      w0 = w1;
    }
    if (chunk < 0) break;
  }
  MN_SYNC();
  if (MN_T0) {
    mn_plc_cache_fill(im, sm);
    mn_commit_merge_object(im, sm, A, 0);
  }
  MN_FOR(c, A.C) im.clp[(size_t)sm.c_surv[0] * A.C + c] = c_clp[c];
  MN_SYNC();
}

// The scheduler for one image.  c_clp: MN_H * C floats of shared memory.
MN_D void mn_merge_image(const MnImage& im, MnSm& sm, const MnMergeArgs& A, float* c_clp) {
  // ---- init ----
  if (MN_T0) {
    sm.hsel = 0; sm.nhot = 0; sm.nins = 0; sm.nne = 0; sm.cold_empty = 0; sm.plcache_n = 0; sm.plcache_used = 0;
    sm.b_mp = 0; sm.b_lo = 0; sm.b_hi = 0; sm.path_n = 0;
    sm.st_rounds = sm.st_events = sm.st_merges = sm.st_restores = sm.st_invalid = sm.st_solo = 0;
    sm.st_refills = sm.st_flushes = sm.st_splits = sm.st_pairs = sm.st_cut_conf = sm.st_cut_casc = sm.st_cut_cap = 0;
    for (int i = 0; i < MN_NCYC; i++) sm.cyc[i] = 0;
    sm.cyc_t0 = 0;
    // number of real (non-sentinel) initial entries: first index whose key is the sentinel
    long long E = (long long)A.N * A.K;
    long long a = 0, b = E;
    while (a < b) { long long mid = (a + b) >> 1; if (im.init_keys[mid] == ~0ull) b = mid; else a = mid + 1; }
    im.ctl->n_init = (int)a;
    im.ctl->static_cursor = 0;
  }
  MN_FOR(i, (int)((MN_NROOTS + 31) / 32)) sm.root_bits[i] = 0;
  MN_FOR(i, (int)(((MN_NROOTS + 31) / 32 + 31) / 32)) sm.root_sum[i] = 0;
  MN_SYNC();

  for (long long round = 0;; round++) {
    MN_SYNC();
    if (im.ctl->status != MN_OK) break;
    if (A.max_rounds > 0 && round >= A.max_rounds) { if (MN_T0) mn_fail(im, MN_ERR_LIMIT); break; }
    MN_TIC();
    if (sm.nins > MN_IC - MN_NE - 64) { mn_flush_ins(im, sm); MN_TOC(MN_CY_FLUSH); }
    if (sm.nhot == 0) {
      mn_refill(im, sm, A);
      MN_TOC(MN_CY_REFILL);
      if (im.ctl->status != MN_OK) break;
      if (sm.nhot == 0) break;  // queue empty: cc:542
    }
    const int ncand0 = sm.nhot < MN_H ? sm.nhot : MN_H;
    // ---- phase 1: classify ----
    MN_FOR(i, MN_CT) { sm.ct_obj[i] = -1; sm.ct_w[i] = INT_MAX; sm.ct_r[i] = INT_MAX; }
    MN_FOR(j, ncand0) mn_classify(im, sm, A, j);
    MN_SYNC();
    // ---- phase 2: capacity cut by pixels; chunk lists of the absorbed objects ----
    if (MN_T0) {
      int tot = 0, n = 0, f = -1, solo = 0;
      for (int j = 0; j < ncand0; j++) {
        if (sm.c_kind[j] != 0 && f < 0) f = j;
        if (sm.c_kind[j] == 2) {
          if (tot + sm.c_npix[j] > MN_PW) { if (j == f) solo = 1; break; }
          tot += sm.c_npix[j];
        }
        n = j + 1;
      }
      sm.ncand = n; sm.solo = solo; sm.tmp0 = f; sm.ncw = 0; sm.npw = 0; sm.npr = 0;
    }
    MN_SYNC();
    if (sm.solo) { mn_solo_merge(im, sm, A, c_clp, sm.tmp0); if (MN_T0) sm.st_rounds++; MN_TOC(MN_CY_SOLO); continue; }
    MN_FOR(j, sm.ncand) {
      if (sm.c_kind[j] == 2) {
        int p = MN_ATOMIC_ADD(&sm.npw, 1);
        sm.pw_cand[p] = j; sm.pw_pix[p] = sm.c_abs[j];
        for (int c = im.pl_head[sm.c_abs[j]]; c >= 0; c = im.plc_next[c]) {
          int w = MN_ATOMIC_ADD(&sm.ncw, 1);
          if (w < MN_CW) { sm.cw_cand[w] = j; sm.cw_chunk[w] = c; }
        }
      }
    }
    MN_SYNC();
    if (sm.ncw > MN_CW) {
      // sparsely filled chunk chains overflowed the chunk work list: shrink the round to its first
      // valid member (a merge goes through the sliced solo path, which has no such limit)
      const int f = sm.tmp0;
      MN_SYNC();
      if (sm.c_kind[f] == 2) { mn_solo_merge(im, sm, A, c_clp, f); if (MN_T0) sm.st_rounds++; continue; }
      if (MN_T0) { sm.ncand = f + 1; sm.ncw = 0; sm.npw = 0; sm.st_cut_cap++; }
      MN_SYNC();
    }
    mn_expand_pixels(im, sm, sm.ncw);
    MN_SYNC();
    // ---- phase 3: count pairs, capacity cut by pairs ----
    MN_FOR(i, sm.npw) {
      int j = sm.pw_cand[i];
      int c = mn_popc(mn_live_bits(im, A, sm.pw_pix[i], sm.c_rec[j]));
      if (c) MN_ATOMIC_ADD(&sm.c_npairs[j], c);
    }
    MN_SYNC();
    if (MN_T0) {
      int tot = 0, n = 0, f = sm.tmp0, solo = 0;
      for (int j = 0; j < sm.ncand; j++) {
        if (sm.c_kind[j] == 2) {
          if (tot + sm.c_npairs[j] > MN_WL) { if (j == f) solo = 1; break; }
          sm.c_pbase[j] = tot;
          tot += sm.c_npairs[j];
        }
        n = j + 1;
      }
      if (n < sm.ncand) sm.st_cut_cap++;
      sm.ncand = n; sm.npr = tot; sm.solo = solo;
    }
    MN_SYNC();
    if (sm.solo) { mn_solo_merge(im, sm, A, c_clp, sm.tmp0); if (MN_T0) sm.st_rounds++; MN_TOC(MN_CY_SOLO); continue; }
    const int ncand = sm.ncand, npr = sm.npr;
    // ---- phase 4: pair lists + staged class vectors ----
    {
      // only pixels of candidates inside the cut contribute
      MN_FOR(ii, sm.npw) {
        int j = sm.pw_cand[ii];
        if (j >= ncand) continue;
        int p = sm.pw_pix[ii];
        uint32_t m = mn_live_bits(im, A, p, sm.c_rec[j]);
        while (m) {
          int bit = 31 - MN_CLZ(m);
          m &= ~(1u << bit);
          int slot = sm.c_pbase[j] + MN_ATOMIC_ADD(&sm.c_pfill[j], 1);
          if (slot < MN_WL) { sm.pr_cand[slot] = j; sm.pr_t[slot] = mn_rec_of_bit(A, p, bit); }
        }
      }
    }
    mn_stage_clp(im, sm, A, c_clp, 0, ncand);
    MN_SYNC();
    MN_TOC(MN_CY_SELECT);
    // ---- phase 5: plan ----
    mn_plan_pairs(im, sm, A, c_clp, 0, npr);
    MN_SYNC();
    MN_TOC(MN_CY_PLAN);
    // ---- phase 6: footprints into the conflict table ----
    MN_FOR(j, ncand) {
      if (sm.c_kind[j] == 0) continue;
      int s1 = mn_ct_slot(sm, sm.c_lo[j]), s2 = mn_ct_slot(sm, sm.c_hi[j]);
      if (s1 < 0 || s2 < 0) { sm.c_conflict[j] = 1; continue; }
      if (sm.c_kind[j] == 2) { MN_ATOMIC_MIN(&sm.ct_w[s1], j); MN_ATOMIC_MIN(&sm.ct_w[s2], j); }
      else { MN_ATOMIC_MIN(&sm.ct_r[s1], j); MN_ATOMIC_MIN(&sm.ct_r[s2], j); }
    }
    MN_FOR(i, npr) {
      int s = mn_ct_slot(sm, sm.pr_x[i]);
      if (s < 0) sm.c_conflict[sm.pr_cand[i]] = 1; else MN_ATOMIC_MIN(&sm.ct_r[s], sm.pr_cand[i]);
    }
    MN_SYNC();
    MN_FOR(j, ncand) {
      if (sm.c_kind[j] == 0) continue;
      int s1 = mn_ct_find(sm, sm.c_lo[j]), s2 = mn_ct_find(sm, sm.c_hi[j]);
      bool cf = false;
      if (s1 >= 0) cf = cf || sm.ct_w[s1] < j || (sm.c_kind[j] == 2 && sm.ct_r[s1] < j);
      if (s2 >= 0) cf = cf || sm.ct_w[s2] < j || (sm.c_kind[j] == 2 && sm.ct_r[s2] < j);
      if (cf) sm.c_conflict[j] = 1;
    }
    MN_FOR(i, npr) {
      int s = mn_ct_find(sm, sm.pr_x[i]);
      if (s >= 0 && sm.ct_w[s] < sm.pr_cand[i]) sm.c_conflict[sm.pr_cand[i]] = 1;
    }
    MN_SYNC();
    // ---- phase 7: accept the longest provably sequential prefix ----
    if (MN_T0) {
      uint32_t runmax = 0;  // bits+1 of the largest priority created by an accepted member
      int cut = ncand, nacc = 0;
      for (int j = 0; j < ncand; j++) {
        if (sm.c_kind[j] == 0) { sm.st_invalid++; continue; }
        if (sm.c_conflict[j] && nacc > 0) { cut = j; sm.st_cut_conf++; break; }
        if (runmax != 0 && runmax - 1u >= mn_f2u(sm.c_key[j]) && nacc > 0) { cut = j; sm.st_cut_casc++; break; }
        sm.c_accept[j] = 1;
        nacc++;
        if (sm.c_maxnew[j] > runmax) runmax = sm.c_maxnew[j];
        if (sm.c_kind[j] == 2) sm.st_merges++; else sm.st_restores++;
      }
      // invalid entries counted past the cut were not consumed
      sm.cutpos = cut; sm.nacc = nacc;
      sm.st_events += nacc; sm.st_rounds++; sm.st_pairs += npr;
      mn_plc_cache_fill(im, sm);
    }
    MN_SYNC();
#ifdef MN_EMUL_TRACE
    fprintf(stderr, "round: ncand %d nacc %d cut %d nhot %d nins %d npr %d cold_empty %d key0 %.9g\n", ncand, sm.nacc, sm.cutpos, sm.nhot, sm.nins, npr, sm.cold_empty, sm.c_key[0]);
#endif
    MN_TOC(MN_CY_ACCEPT);
    // ---- phase 8: commit ----
    MN_FOR(j, ncand) {
      if (!sm.c_accept[j]) continue;
      if (sm.c_kind[j] == 1) {  // cc:563-565
        int rec = sm.c_rec[j];
        float4 v = im.rec_val[rec];
        v.w = sm.c_newmp[j];
        im.rec_val[rec] = v;
        if (v.w >= 0.0f) mn_push_entry(sm, v.w, sm.c_lo[j], sm.c_hi[j], rec);
      } else {
        mn_commit_merge_object(im, sm, A, j);
      }
    }
    MN_FOR(i, ncand * A.C) {
      int j = i / A.C, c = i % A.C;
      if (sm.c_accept[j] && sm.c_kind[j] == 2) im.clp[(size_t)sm.c_surv[j] * A.C + c] = c_clp[(size_t)j * A.C + c];
    }
    mn_commit_pairs(im, sm, A, 0, npr);
    MN_SYNC();
    MN_TOC(MN_CY_COMMIT);
    // ---- phase 9: queue maintenance ----
    mn_hot_update(im, sm, sm.cutpos);
    MN_TOC(MN_CY_HOT);
  }
  MN_SYNC();
  if (MN_T0) {
    MnCtl* c = im.ctl;
    c->rounds = sm.st_rounds; c->events = sm.st_events; c->merges = sm.st_merges; c->restores = sm.st_restores;
    c->invalid_pops = sm.st_invalid; c->solo_events = sm.st_solo; c->refills = sm.st_refills;
    c->flushes = sm.st_flushes; c->splits = sm.st_splits; c->pairs = sm.st_pairs;
    c->cuts_conflict = sm.st_cut_conf; c->cuts_cascade = sm.st_cut_casc; c->cuts_capacity = sm.st_cut_cap;
    for (int i = 0; i < MN_NCYC; i++) c->cyc[i] = sm.cyc[i];
  }
  MN_SYNC();
}
