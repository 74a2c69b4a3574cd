// mn_exact.cuh -- TIE-EXACT replay of the reference's C++ segmenter (utils/csegment/segment.cc, `cc:LINE`; Mode A).
//
// The hot path (mn_merge.cuh) reproduces the reference's sequential merge order wherever that order is decided by the
// priorities, and breaks ties among EQUAL priorities by a fixed rule of its own (mn_common.h: mn_tie).  The reference
// has no such rule: PriorityCompare sees the priority only (segment.h:270-275), so among equals it pops whatever
// libstdc++'s binary heap has at its front, which depends on the order of every earlier push -- and Merge() pushes in
// the iteration order of the absorbed object's std::unordered_map (cc:650-652).  On inputs whose final partition
// depends on the tie order (block-quantized maps) the two therefore disagree.  This file closes that
// gap for callers who need the reference's exact result on such inputs: ONE thread replays the reference's loop with
// libstdc++'s heap and hash-table orders restated literally (mn_stl_order.h), including the `objects` map whose
// iteration order numbers the output labels (cc:503-515) -- so mask and object_class equal the reference's RAW
// arrays, not only up to relabelling.  Strictly sequential by nature; for small and medium images.
// The arithmetic (class log-probs, sameness / differentness logs, the same_different_bias rewrite) comes from the
// same edge pass as the hot path (mn_edge.cuh); this file starts from its outputs.
//
// Written for both the device and the host (tests/emul/emul_exact.cpp builds it for the CPU suite, which compares
// it with the unmodified reference; test infrastructure only).
#pragma once
#include <stdint.h>

#include "mn_common.h"
#include "mn_stl_order.h"

#define MNX_FLT_MIN 1.17549435e-38f  // numeric_limits<float>::min(): the "practically deleted" priority (cc:694)

struct MnExact {
  int C, K, H, W, N;
  long long E;
  int off_r[MN_MAX_K], off_c[MN_MAX_K];
  float omf, mlb;
  // edge pass outputs
  float* clp;             // [N][C] class_logprobs of the single-pixel objects (cc:5-21); accumulated in place (cc:640)
  const float* rec_same;  // [E] logf(s)                     record slot = pixel * K + offset index
  const float* rec_diff;  // [E] (float)log(1.0 - s)         (cc:32-36)
  // objects (segment.h:86-137)
  int* npix; int* cls; int* pix_next; int* pix_tail;
  MnStlTab* tab;          // [N + 1]: adjacency_list of every object; [N] = ObjectSegmenter::objects (segment.h:331)
  int* ob_next;           // [N] nodes of `objects` (key = object id)
  // records (segment.h:175-232)
  int* r_o1; int* r_o2;   // obj1 / obj2 (ids; -1 = NULL, cc:727)
  float* r_oml; float* r_mp; int* r_merged;
  int* nd_next;           // [2E] two hash-table nodes per record: one in each endpoint's adjacency_list
  unsigned long long* nd_key;  // [2E] the key the node was inserted under (cached_hash at that time)
  MnStlHeap heap;         // segmenter_queue (segment.h:335-336)
  MnStlArena arena;
  // outputs (cc:491-517)
  int* out_mask; int* out_cls; int* out_n;
  int* status;            // 0 ok, 1 queue full, 2 bucket arena full, 3 inconsistent adjacency list (the reference exit(1)s), 4 duplicate pixel pair
  long long* stats;       // pops, merges, pushes, bucket-arena collections
};

struct MnxRecNodes {
  int* nx; unsigned long long* k;
  MN_HD unsigned long long key(int n) const { return k[n]; }
  MN_HD int next(int n) const { return nx[n]; }
  MN_HD void set_next(int n, int v) const { nx[n] = v; }
};
struct MnxObjNodes {
  int* nx;
  MN_HD unsigned long long key(int n) const { return (unsigned long long)n; }
  MN_HD int next(int n) const { return nx[n]; }
  MN_HD void set_next(int n, int v) const { nx[n] = v; }
};

// AdjacencyRecordHasher (segment.h:237-242) on the sorted ids (cc:49-56)
MN_HD unsigned long long mnx_hash(int lo, int hi) { return (unsigned long long)lo * 1619ull + (unsigned long long)hi * 3203ull; }

// UpdateMergePriority (cc:107-150)
MN_HD void mnx_update_priority(const MnExact& m, int rec) {
  const int a = m.r_o1[rec], b = m.r_o2[rec];
  int merged;
  m.r_mp[rec] = mn_priority(m.r_oml[rec], m.omf, m.mlb, m.C, m.npix[a], m.cls[a], m.clp + (size_t)a * m.C, m.npix[b],
                            m.cls[b], m.clp + (size_t)b * m.C, &merged);
  m.r_merged[rec] = merged;
}
MN_HD void mnx_push(MnExact& m, int rec) {  // segmenter_queue.push(make_pair(priority, arec))
  if (!mns_heap_push(m.heap, m.r_mp[rec], rec)) *m.status = 1;
  m.stats[2]++;
}

// the constructor (cc:196-232)
MN_HD void mnx_init(MnExact& m) {
  const int N = m.N, C = m.C, K = m.K, W = m.W, H = m.H;
  const MnxRecNodes rn{m.nd_next, m.nd_key};
  const MnxObjNodes on{m.ob_next};
  for (int i = 0; i <= N; i++) mns_tab_init(m.tab[i]);
  for (int p = 0; p < N; p++) {  // cc:196-207, Object ctor cc:5-21 (first maximum)
    const float* c = m.clp + (size_t)p * C;
    int best = 0;
    for (int i = 1; i < C; i++) if (c[i] > c[best]) best = i;
    m.cls[p] = best; m.npix[p] = 1; m.pix_next[p] = -1; m.pix_tail[p] = p;
    mns_insert(m.tab[N], m.arena, on, (unsigned long long)p, p);  // objects[obj_id] = obj
  }
  for (long long r = 0; r < m.E; r++) { m.r_o1[r] = -1; m.r_o2[r] = -1; }
  for (int row = 0; row < H && *m.status == 0; row++) {
    for (int col = 0; col < W; col++) {
      const int p = row * W + col;
      for (int k = 0; k < K; k++) {
        const int r2 = row + m.off_r[k], c2 = col + m.off_c[k];
        if (r2 < 0 || r2 >= H || c2 < 0 || c2 >= W) continue;
        const int q = r2 * W + c2;
        const int rec = p * K + k;
        const int lo = p < q ? p : q, hi = p < q ? q : p;
        const unsigned long long h = mnx_hash(lo, hi);
        if (p == q || mns_find(m.tab[p], m.arena, rn, h) >= 0) { *m.status = 4; return; }  // (an offset list holding o and -o: undefined in the reference)
        m.r_o1[rec] = lo; m.r_o2[rec] = hi;
        m.r_oml[rec] = MN_FSUB(m.rec_same[rec], m.rec_diff[rec]);  // cc:36
        mnx_update_priority(m, rec);
        // obj1 = the source pixel's object, obj2 = the target's (cc:212-224): node 2 rec in p's list, 2 rec + 1 in q's
        m.nd_key[2 * (size_t)rec] = h; m.nd_key[2 * (size_t)rec + 1] = h;
        mns_insert(m.tab[p], m.arena, rn, h, 2 * rec);
        mns_insert(m.tab[q], m.arena, rn, h, 2 * rec + 1);
        if (m.r_mp[rec] >= 0) mnx_push(m, rec);  // cc:225-227
      }
    }
    if (*m.arena.overflow) { *m.status = 2; return; }
  }
}

// Merge (cc:602-727)
MN_HD void mnx_merge(MnExact& m, int arec) {
  const MnxRecNodes rn{m.nd_next, m.nd_key};
  const MnxObjNodes on{m.ob_next};
  int o1 = m.r_o1[arec], o2 = m.r_o2[arec];
  if (o1 < 0 || o2 < 0 || o1 == o2) return;
  if (m.npix[o1] < m.npix[o2]) { const int t = o1; o1 = o2; o2 = t; }  // cc:612-616
  m.cls[o1] = m.r_merged[arec];  // cc:635
  m.pix_next[m.pix_tail[o1]] = o2; m.pix_tail[o1] = m.pix_tail[o2];  // cc:636-639
  m.npix[o1] += m.npix[o2];
  for (int c = 0; c < m.C; c++) m.clp[(size_t)o1 * m.C + c] = MN_FADD(m.clp[(size_t)o1 * m.C + c], m.clp[(size_t)o2 * m.C + c]);  // cc:640
  const unsigned long long ah = mnx_hash(m.r_o1[arec], m.r_o2[arec]);
  mns_erase(m.tab[o1], m.arena, rn, ah);  // cc:645-647
  mns_erase(m.tab[o2], m.arena, rn, ah);
  for (int node = m.tab[o2].first; node != MNS_NULL;) {  // cc:650-652: the absorbed object's list in container order
    const int next_node = m.nd_next[node];  // (the list of o2 is not touched below; its node is reused for o1's list)
    const int t = node >> 1;
    int o3;
    if (m.r_o1[t] == o2) o3 = m.r_o2[t];
    else if (m.r_o2[t] == o2) o3 = m.r_o1[t];
    else { *m.status = 3; return; }  // cc:664-667
    if (o3 == o1) { *m.status = 3; return; }  // cc:669-672 "cyclic merging"
    const unsigned long long old_h = m.nd_key[node];
    const int lo = o1 < o3 ? o1 : o3, hi = o1 < o3 ? o3 : o1;  // cc:658-663,676 SortAndUpdateHash
    const unsigned long long new_h = mnx_hash(lo, hi);
    m.r_o1[t] = lo; m.r_o2[t] = hi;
    const int node3 = mns_erase(m.tab[o3], m.arena, rn, old_h);  // cc:680
    if (node3 < 0) { *m.status = 3; return; }
    const int that_node = mns_find(m.tab[o1], m.arena, rn, new_h);  // cc:684
    if (that_node >= 0) {
      const int that = that_node >> 1;
      m.r_oml[that] = MN_FADD(m.r_oml[that], m.r_oml[t]);  // cc:689 (the sameness / differentness sums only feed a printed total)
      m.r_mp[t] = MNX_FLT_MIN;  // cc:694
      mnx_update_priority(m, that);
      if (m.r_mp[that] >= 0) mnx_push(m, that);
    } else {
      m.nd_key[node] = new_h; m.nd_key[node3] = new_h;
      mns_insert(m.tab[o1], m.arena, rn, new_h, node);   // cc:700
      mns_insert(m.tab[o3], m.arena, rn, new_h, node3);  // cc:701
      mnx_update_priority(m, t);
      if (m.r_mp[t] >= 0) mnx_push(m, t);
    }
    node = next_node;
  }
  mns_erase(m.tab[m.N], m.arena, on, (unsigned long long)o2);  // cc:724 objects.erase(obj2->GetId())
  mns_tab_drop(m.tab[o2]);  // cc:725 delete obj2 (its bucket array is reclaimed by the next collection)
  m.r_o2[arec] = -1;  // cc:727 arec->SetObj2(NULL)
  m.stats[1]++;
}

// RunSegmentation (cc:539-573) + OutputMask (cc:491-517)
MN_HD void mnx_run(MnExact& m) {
  mnx_init(m);
  while (m.heap.n > 0 && *m.status == 0) {
    float mp; int rec;
    mns_heap_pop(m.heap, &mp, &rec);  // top() + pop()
    m.stats[0]++;
    if (mp != m.r_mp[rec]) continue;                      // cc:555
    if (m.r_o1[rec] < 0 || m.r_o2[rec] < 0) continue;     // cc:558
    mnx_update_priority(m, rec);                          // cc:561
    if (m.r_mp[rec] == mp) mnx_merge(m, rec);             // cc:562
    else if (m.r_mp[rec] >= 0) mnx_push(m, rec);          // cc:564
    if (*m.arena.overflow) *m.status = 2;
  }
  if (*m.status) return;
  for (int p = 0; p < m.N; p++) { m.out_mask[p] = 0; m.out_cls[p] = -1; }
  int k = 1;
  for (int o = m.tab[m.N].first; o != MNS_NULL; o = m.ob_next[o]) {  // `objects` in container order
    if (m.cls[o] == 0) continue;
    m.out_cls[k - 1] = m.cls[o];
    for (int p = o; p >= 0; p = m.pix_next[p]) m.out_mask[p] = k;
    k++;
  }
  *m.out_n = k - 1;
}
