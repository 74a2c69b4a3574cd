// mn_modeb.cuh -- "Mode B": the semantics of the reference's pure-Python segmenter (utils/segmenter.py, `py:LINE`),
// the class the COCO recipe calls (egs/coco/local/segment.py:155-164).  It differs from the C++ port (Mode A, the
// hot path of this library) in the priority formula (den = n1 * n2, bias inside the division: py:189-193), the accept
// rule (>=, py:470), the dead-record marker (-100000.0, py:562), float64 class accumulators (py:51), the
// post-pass prune(200) (py:351-375) and in what breaks ties: Python's heapq over tuples (-priority, record), whose
// order among equal priorities is decided by the records' CURRENT priorities (py:217-218) and by the sift order of
// heapq itself, and the insertion order of the adjacency dicts (py:534: `for this_arec in obj2.adjacency_list`).
// None of that is a total order that parallel rounds could reproduce, so this mode is strictly sequential: ONE
// thread of one CTA per image replays heapq and the dict orders exactly (arrays in global memory).  It is the
// secondary, small-image mode (the Python reference itself is practical to ~128 x 256), not the B200 hot path.
//
// Arithmetic types follow NumPy 2 (NEP 50), which is what the reference computes with in this image:
//   * np.log of the float32 maps is float32; `1.0 - same_prob` stays float32           (py:135-136)
//   * record sums (sameness, differentness, obj_merge_logprob) accumulate in float32     (py:557-559)
//   * class_logprobs accumulate in float64                                              (py:51-54, 524)
//   * priority: float32 arithmetic when the classes are equal (class_delta_logprob is the Python float 0.0 and
//     Python scalars are weak), float64 when they differ (class_delta_logprob is a float64)   (py:179-193)
// The logarithms themselves are NOT evaluated here: the host side hands in np.log(...) arrays computed by the
// caller's own NumPy (mergenet_b200/segmenter.py), so that they carry exactly the bits the reference would use.
//
// Written for both the device and the host (tests/emul builds it for the CPU suite; test infrastructure only).
#pragma once
#include <stdint.h>

#include "mn_common.h"

struct MnModeB {
  int C, K, H, W, N;
  long long E;
  const float* logc;    // [C][N] np.log(class_probs)
  const float* lsame;   // [K][N] np.log(sameness)
  const float* ldiff;   // [K][N] np.log(1.0 - sameness)
  int off_r[MN_MAX_K], off_c[MN_MAX_K];
  float omf32, mlb32;   // the Python floats rounded to float32 (weak-scalar promotion)
  double omf, mlb;
  double prune_threshold;
  // objects (py:27-90)
  int* npix; int* cls; double* clp; float* osame; unsigned char* alive;
  int* adj_head; int* adj_tail;           // adjacency "dict" in insertion order: doubly linked through the records
  int* pix_next; int* pix_tail;           // pixel set as a linked list (head = the object's own pixel)
  // records (py:93-222)
  int* r_o1; int* r_o2;                   // obj1.id <= obj2.id (py:199-202)
  float* r_oml; float* r_same; float* r_diff;
  double* r_mp;                           // merge_priority (float32-valued or float64)
  int* r_link;                            // [E][2][3]: per endpoint slot (object id, prev record, next record)
  // (obj1.id, obj2.id) -> live record: what `this_arec in obj1.adjacency_list` looks up (py:555)
  unsigned long long* h_key; int* h_val; unsigned h_mask;
  // heapq (py:289,464-473)
  double* q_key; int* q_rec; long long q_n, q_cap;
  // outputs
  long long* out_mask; int* out_cls; int* out_n;
  // status / statistics
  int* status;                            // 0 ok, 1 heap overflow, 2 hash overflow, 3 prune found no background object
  long long* stats;                       // pops, merges, pushes, pruned
};

#define MNB_DEAD (-100000.0)
#define MNB_EMPTY 0xFFFFFFFFFFFFFFFFull
#define MNB_TOMB 0xFFFFFFFFFFFFFFFEull

MN_HD unsigned long long mnb_pair(int a, int b) { return ((unsigned long long)(uint32_t)a << 32) | (uint32_t)b; }
MN_HD unsigned mnb_hash(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return (unsigned)k;
}
MN_HD int mnb_find(const MnModeB& m, int a, int b) {
  const unsigned long long key = mnb_pair(a, b);
  for (unsigned i = mnb_hash(key) & m.h_mask, n = 0; n <= m.h_mask; i = (i + 1) & m.h_mask, n++) {
    const unsigned long long k = m.h_key[i];
    if (k == key) return m.h_val[i];
    if (k == MNB_EMPTY) return -1;
  }
  return -1;
}
MN_HD void mnb_insert(const MnModeB& m, int a, int b, int rec) {
  const unsigned long long key = mnb_pair(a, b);
  for (unsigned i = mnb_hash(key) & m.h_mask, n = 0; n <= m.h_mask; i = (i + 1) & m.h_mask, n++) {
    const unsigned long long k = m.h_key[i];
    if (k == MNB_EMPTY || k == MNB_TOMB) { m.h_key[i] = key; m.h_val[i] = rec; return; }
  }
  *m.status = 2;
}
MN_HD void mnb_erase(const MnModeB& m, int a, int b) {
  const unsigned long long key = mnb_pair(a, b);
  for (unsigned i = mnb_hash(key) & m.h_mask, n = 0; n <= m.h_mask; i = (i + 1) & m.h_mask, n++) {
    const unsigned long long k = m.h_key[i];
    if (k == key) { m.h_key[i] = MNB_TOMB; return; }
    if (k == MNB_EMPTY) return;
  }
}

// ---- adjacency dicts in insertion order -----------------------------------------------------------
MN_HD int mnb_slot(const MnModeB& m, int rec, int obj) { return m.r_link[(size_t)rec * 6 + 0] == obj ? 0 : 1; }
MN_HD void mnb_list_append(const MnModeB& m, int obj, int rec, int slot) {  // dict[rec] = rec for a new key
  int* L = m.r_link + (size_t)rec * 6 + slot * 3;
  L[0] = obj; L[1] = m.adj_tail[obj]; L[2] = -1;
  if (m.adj_tail[obj] >= 0) m.r_link[(size_t)m.adj_tail[obj] * 6 + mnb_slot(m, m.adj_tail[obj], obj) * 3 + 2] = rec;
  else m.adj_head[obj] = rec;
  m.adj_tail[obj] = rec;
}
MN_HD void mnb_list_remove(const MnModeB& m, int obj, int rec) {  // del dict[rec]
  const int* L = m.r_link + (size_t)rec * 6 + mnb_slot(m, rec, obj) * 3;
  const int prev = L[1], next = L[2];
  if (prev >= 0) m.r_link[(size_t)prev * 6 + mnb_slot(m, prev, obj) * 3 + 2] = next; else m.adj_head[obj] = next;
  if (next >= 0) m.r_link[(size_t)next * 6 + mnb_slot(m, next, obj) * 3 + 1] = prev; else m.adj_tail[obj] = prev;
}

// ---- priority (py:179-193) --------------------------------------------------------------------------
MN_HD void mnb_update_priority(const MnModeB& m, int rec, int* merged_out) {
  const int a = m.r_o1[rec], b = m.r_o2[rec];
  const float t32 = MN_FMUL(m.r_oml[rec], m.omf32);  // np.float32 * python float -> float32
  const double den = (double)m.npix[a] * (double)m.npix[b];  // python int product (exact below 2^53)
  int merged;
  double mp;
  if (m.cls[a] == m.cls[b]) {  // class_delta_logprob = 0.0 (python float): everything stays float32
    merged = m.cls[a];
    float t = MN_FADD(t32, 0.0f);
    t = MN_FADD(t, m.mlb32);
    t = MN_FDIV(t, (float)den);
    mp = (double)t;
  } else {  // float64 class_delta_logprob promotes the sum
    const double* ca = m.clp + (size_t)a * m.C;
    const double* cb = m.clp + (size_t)b * m.C;
    double best = MN_DADD(ca[0], cb[0]);
    merged = 0;
    for (int c = 1; c < m.C; c++) {
      const double j = MN_DADD(ca[c], cb[c]);
      if (j > best) { best = j; merged = c; }  // np.argmax: first maximum
    }
    double cdl = MN_DADD(best, -ca[m.cls[a]]);
    cdl = MN_DADD(cdl, -cb[m.cls[b]]);
    double d = MN_DADD((double)t32, cdl);
    d = MN_DADD(d, m.mlb);
#if defined(__CUDA_ARCH__)
    mp = __ddiv_rn(d, den);
#else
    mp = d / den;
#endif
  }
  m.r_mp[rec] = mp;
  if (merged_out) *merged_out = merged;
}

// ---- heapq (CPython Lib/heapq.py _siftdown / _siftup; Modules/_heapqmodule.c is the same algorithm) --
// item = (-merge_priority at push time, record).  Tuple `<`: first differing element decides; equal keys fall to
// the records: equal (same id pair, py:206-207) -> not less; else record.__lt__ = CURRENT priorities (py:217-218).
MN_HD bool mnb_item_lt(const MnModeB& m, double ka, int ra, double kb, int rb) {
  if (ka != kb) return ka < kb;
  if (ra == rb) return false;
  if (m.r_o1[ra] == m.r_o1[rb] && m.r_o2[ra] == m.r_o2[rb]) return false;
  return m.r_mp[ra] < m.r_mp[rb];
}
MN_HD void mnb_siftdown(const MnModeB& m, long long startpos, long long pos) {
  const double nk = m.q_key[pos]; const int nr = m.q_rec[pos];
  while (pos > startpos) {
    const long long parent = (pos - 1) >> 1;
    const double pk = m.q_key[parent]; const int pr = m.q_rec[parent];
    if (mnb_item_lt(m, nk, nr, pk, pr)) { m.q_key[pos] = pk; m.q_rec[pos] = pr; pos = parent; continue; }
    break;
  }
  m.q_key[pos] = nk; m.q_rec[pos] = nr;
}
MN_HD void mnb_siftup(MnModeB& m, long long pos) {
  const long long endpos = m.q_n, startpos = pos;
  const double nk = m.q_key[pos]; const int nr = m.q_rec[pos];
  long long child = 2 * pos + 1;
  while (child < endpos) {
    const long long right = child + 1;
    if (right < endpos && !mnb_item_lt(m, m.q_key[child], m.q_rec[child], m.q_key[right], m.q_rec[right])) child = right;
    m.q_key[pos] = m.q_key[child]; m.q_rec[pos] = m.q_rec[child];
    pos = child;
    child = 2 * pos + 1;
  }
  m.q_key[pos] = nk; m.q_rec[pos] = nr;
  mnb_siftdown(m, startpos, pos);
}
MN_HD void mnb_heappush(MnModeB& m, double key, int rec) {
  if (m.q_n >= m.q_cap) { *m.status = 1; return; }
  m.q_key[m.q_n] = key; m.q_rec[m.q_n] = rec;
  m.q_n++;
  mnb_siftdown(m, 0, m.q_n - 1);
  m.stats[2]++;
}
MN_HD void mnb_heappop(MnModeB& m, double* key, int* rec) {
  m.q_n--;
  const double lk = m.q_key[m.q_n]; const int lr = m.q_rec[m.q_n];
  if (m.q_n > 0) {
    *key = m.q_key[0]; *rec = m.q_rec[0];
    m.q_key[0] = lk; m.q_rec[0] = lr;
    mnb_siftup(m, 0);
  } else {
    *key = lk; *rec = lr;
  }
}

// ---- init (py:262-289) ---------------------------------------------------------------------------------
MN_HD void mnb_init(MnModeB& m) {
  const int N = m.N, C = m.C, K = m.K, W = m.W, H = m.H;
  for (int p = 0; p < N; p++) {
    double best = 0; int bc = 0;
    for (int c = 0; c < C; c++) {
      const double v = MN_DADD(0.0, (double)m.logc[(size_t)c * N + p]);  // np.zeros (float64) += float32 log
      m.clp[(size_t)p * C + c] = v;
      if (c == 0 || v > best) { best = v; bc = c; }
    }
    m.cls[p] = bc; m.npix[p] = 1; m.osame[p] = 0.0f; m.alive[p] = 1;
    m.adj_head[p] = -1; m.adj_tail[p] = -1; m.pix_next[p] = -1; m.pix_tail[p] = p;
  }
  for (long long r = 0; r < m.E; r++) { m.r_o1[r] = -1; m.r_o2[r] = -1; m.r_mp[r] = MNB_DEAD; }
  for (int row = 0; row < H; row++) {
    for (int col = 0; col < W; col++) {
      const int p = row * W + col;
      for (int k = 0; k < K; k++) {
        const int r2 = row + m.off_r[k], c2 = col + m.off_c[k];
        if (r2 < 0 || r2 >= H || c2 < 0 || c2 >= W) continue;
        const int q = r2 * W + c2;
        const int rec = p * K + k;
        const int a = p < q ? p : q, b = p < q ? q : p;
        m.r_o1[rec] = a; m.r_o2[rec] = b;
        const float ls = m.lsame[(size_t)k * N + p], ld = m.ldiff[(size_t)k * N + p];
        m.r_diff[rec] = ld; m.r_same[rec] = ls; m.r_oml[rec] = MN_FSUB(ls, ld);  // py:137-139
        mnb_update_priority(m, rec, nullptr);
        // adjacency_records[arec] = arec; obj1.adjacency_list[arec]; obj2.adjacency_list[arec]  (obj1 = source pixel)
        mnb_insert(m, a, b, rec);
        m.r_link[(size_t)rec * 6 + 0] = p; m.r_link[(size_t)rec * 6 + 3] = q;
        mnb_list_append(m, p, rec, 0);
        mnb_list_append(m, q, rec, 1);
        if (m.r_mp[rec] >= 0) mnb_heappush(m, -m.r_mp[rec], rec);
      }
    }
  }
}

// ---- merge (py:485-578) ----------------------------------------------------------------------------------
MN_HD void mnb_merge(MnModeB& m, int arec, int merged_class) {
  int o1 = m.r_o1[arec], o2 = m.r_o2[arec];
  if (!m.alive[o1] || !m.alive[o2]) return;  // py:513
  if (o1 == o2) return;
  if (m.npix[o2] > m.npix[o1]) { const int t = o1; o1 = o2; o2 = t; }  // py:517
  m.cls[o1] = merged_class;  // py:522 (arec.merged_class of the update that preceded this call)
  // pixels (py:523): obj1's list then obj2's
  m.pix_next[m.pix_tail[o1]] = o2; m.pix_tail[o1] = m.pix_tail[o2];
  m.npix[o1] += m.npix[o2];
  for (int c = 0; c < m.C; c++) m.clp[(size_t)o1 * m.C + c] = MN_DADD(m.clp[(size_t)o1 * m.C + c], m.clp[(size_t)o2 * m.C + c]);
  m.osame[o1] = MN_FADD(m.osame[o1], MN_FADD(m.r_same[arec], m.osame[o2]));  // py:525
  // py:527-529
  mnb_erase(m, m.r_o1[arec], m.r_o2[arec]);
  mnb_list_remove(m, o1, arec);
  mnb_list_remove(m, o2, arec);
  for (int t = m.adj_head[o2]; t >= 0;) {  // py:530: obj2's dict in insertion order
    const int tnext = m.r_link[(size_t)t * 6 + mnb_slot(m, t, o2) * 3 + 2];
    const int o3 = m.r_o1[t] == o2 ? m.r_o2[t] : m.r_o1[t];
    mnb_list_remove(m, o3, t);               // py:535
    mnb_erase(m, m.r_o1[t], m.r_o2[t]);      // py:536
    // py:537-541: re-point obj2 -> obj1, sort the ids
    const int na = o1 < o3 ? o1 : o3, nb = o1 < o3 ? o3 : o1;
    m.r_o1[t] = na; m.r_o2[t] = nb;
    const int that = mnb_find(m, na, nb);    // py:545: `this_arec in obj1.adjacency_list`
    if (that >= 0) {
      m.r_oml[that] = MN_FADD(m.r_oml[that], m.r_oml[t]);
      m.r_diff[that] = MN_FADD(m.r_diff[that], m.r_diff[t]);
      m.r_same[that] = MN_FADD(m.r_same[that], m.r_same[t]);
      m.r_mp[t] = MNB_DEAD;                  // py:551
      // (re-assigning an existing dict key keeps its position: py:552-553)
      mnb_update_priority(m, that, nullptr);
      if (m.r_mp[that] >= 0) mnb_heappush(m, -m.r_mp[that], that);
    } else {
      // the slot of the record that pointed at obj2 now points at obj1; appended to both dicts (py:558-560)
      const int s2 = mnb_slot(m, t, o2);
      mnb_list_append(m, o1, t, s2);
      mnb_list_append(m, o3, t, 1 - s2);
      mnb_insert(m, na, nb, t);
      mnb_update_priority(m, t, nullptr);
      if (m.r_mp[t] >= 0) mnb_heappush(m, -m.r_mp[t], t);
    }
    t = tnext;
  }
  m.alive[o2] = 0;  // py:578
  m.stats[1]++;
}

// ---- run (py:432-483), prune (py:351-375), output_mask (py:377-389) ------------------------------------------
MN_HD void mnb_run(MnModeB& m) {
  mnb_init(m);
  while (m.q_n > 0 && *m.status == 0) {
    double key; int rec;
    mnb_heappop(m, &key, &rec);
    m.stats[0]++;
    const double mp = -key;
    if (mp != m.r_mp[rec]) continue;                 // py:466
    int merged;
    mnb_update_priority(m, rec, &merged);            // py:468
    if (m.r_mp[rec] >= mp) mnb_merge(m, rec, merged);  // py:469-470
    else if (m.r_mp[rec] >= 0) mnb_heappush(m, -m.r_mp[rec], rec);
  }
  if (*m.status) return;
  // prune: the biggest class-0 object (first one among equals, ascending id = dict order)
  int bg = -1, bgn = 0;
  for (int o = 0; o < m.N; o++)
    if (m.alive[o] && m.cls[o] == 0 && m.npix[o] > bgn) { bg = o; bgn = m.npix[o]; }
  for (int o = 0; o < m.N; o++) {
    if (!m.alive[o]) continue;
    const double score = MN_DADD(m.clp[(size_t)o * m.C + m.cls[o]], -m.clp[(size_t)o * m.C + 0]);
    if (score < m.prune_threshold) {
      if (bg < 0) { *m.status = 3; return; }   // `background_obj` unbound: the reference raises UnboundLocalError
      if (o != bg) {
        m.pix_next[m.pix_tail[bg]] = o; m.pix_tail[bg] = m.pix_tail[o]; m.npix[bg] += m.npix[o];
        m.alive[o] = 0;
        m.stats[3]++;
      }
    }
  }
  // output_mask: labels in ascending surviving id, class-0 objects skipped
  for (int p = 0; p < m.N; p++) m.out_mask[p] = 0;
  int k = 1;
  for (int o = 0; o < m.N; o++) {
    if (!m.alive[o] || m.cls[o] == 0) continue;
    m.out_cls[k - 1] = m.cls[o];
    for (int p = o; p >= 0; p = m.pix_next[p]) m.out_mask[p] = k;
    k++;
  }
  *m.out_n = k - 1;
}
