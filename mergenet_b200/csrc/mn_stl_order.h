// mn_stl_order.h -- the two libstdc++ containers whose INTERNAL ORDER the reference's results depend on, restated
// on flat integer arrays for one sequential thread (device or host):
//
//   * std::unordered_map<size_t, T*>  (segment.h:136,331: Object::adjacency_list, ObjectSegmenter::objects).
//     Merge() walks the absorbed object's adjacency_list in the container's iteration order (cc:650-652) and
//     pushes the touched records in that order; OutputMask() numbers the instances in the iteration order of
//     `objects` (cc:503-515).  libstdc++'s table is one singly linked node list threaded through all buckets, a
//     bucket pointing at the node BEFORE its first one; a node that lands in an empty bucket goes to the FRONT of
//     the whole list, otherwise to the front of its bucket's run; the table grows through the primes of
//     mn_stl_primes.h (first 13 buckets, then at least twice as many) and re-threads the list on every growth.
//     Keys are size_t and std::hash<size_t> is the identity, so bucket = key % bucket_count.
//   * std::priority_queue<pair<float, AdjacencyRecord*>, vector<...>, PriorityCompare> (segment.h:270-275,335):
//     PriorityCompare looks at the priority only, so which of several EQUAL priorities pops first is whatever
//     std::push_heap / std::pop_heap leave at the front of the vector.
//
// Nothing here is taken from the reference tree: it is the behaviour of GCC 13's libstdc++ (the toolchain the
// unmodified reference is compiled with in this image), written from scratch and pinned against the real containers
// by tests/emul/stl_order_check.cpp (random operation sequences: same iteration order, same pop order).
// Used by the tie-exact replay only (mn_exact.cuh); the B200 hot path (mn_merge.cuh) uses its own fixed tie order.
#pragma once
#include "mn_common.h"

#define MNS_NULL (-1)      // nullptr
#define MNS_BB (-2)        // &_M_before_begin (the list head seen as a node)
#define MNS_NOTFOUND (-3)

struct MnStlTab {          // one std::unordered_map<size_t, T*>
  int first;               // _M_before_begin._M_nxt
  int single;              // _M_single_bucket (the bucket array of a table that never held anything)
  unsigned nbkt;           // _M_bucket_count
  unsigned cnt;            // _M_element_count
  unsigned long long next_resize;  // _Prime_rehash_policy::_M_next_resize (max_load_factor = 1)
  long long boff;          // where the bucket array lives in the arena (nbkt > 1)
};

// Bucket arrays of every table: two half-spaces, bump allocated in the current one.  A table abandons its array when it
// grows and when its owner dies; when the current half is full the live arrays (tables with nbkt > 1) are copied, packed,
// to the other half -- bucket words hold node ids, not addresses, so they move freely.  The live arrays stay below
// ~2.3 x (2 nodes per record) + 13 per object words (measured peak: ~3 words per record slot); a half of 5 E + 32 N
// words always fits.
struct MnStlArena {
  int* bk;                 // 2 * half words
  long long half;
  long long* bump;         // next free word
  long long* base;         // start of the current half (0 or half)
  MnStlTab* tabs;          // every table that may own an array, for the collection
  int ntabs;
  const unsigned* primes;  // MN_STL_PRIMES
  int* overflow;           // set to 1 when the live arrays do not fit a half or the prime table is exhausted
  long long* collections;
};

MN_HD void mns_tab_init(MnStlTab& t) {
  t.first = MNS_NULL; t.single = MNS_NULL; t.nbkt = 1; t.cnt = 0; t.next_resize = 0; t.boff = 0;
}

// _Prime_rehash_policy::_M_next_bkt: smallest listed prime >= n (a short fixed answer below 14)
MN_HD unsigned mns_next_bkt(const MnStlArena& A, unsigned long long n, unsigned long long* next_resize) {
  if (n < 14) {
    if (n == 0) return 1;
    const unsigned f = n <= 2 ? 2u : n == 3 ? 3u : n <= 5 ? 5u : n <= 7 ? 7u : n <= 11 ? 11u : 13u;
    *next_resize = f;
    return f;
  }
  int lo = 6, hi = 256;  // lower_bound over primes[6, 256)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((unsigned long long)A.primes[mid] < n) lo = mid + 1; else hi = mid;
  }
  if (lo >= 256) { *A.overflow = 1; lo = 255; }
  *next_resize = A.primes[lo];
  return A.primes[lo];
}
// _Prime_rehash_policy::_M_need_rehash for ONE insertion: the new bucket count, or 0
MN_HD unsigned mns_need_rehash(const MnStlArena& A, MnStlTab& t) {
  const unsigned long long want = (unsigned long long)t.cnt + 1;
  if (want <= t.next_resize) return 0;
  unsigned long long min_bkts = want;
  if (t.next_resize == 0 && min_bkts < 11) min_bkts = 11;  // a table that allocates for the first time starts at 11 (-> 13)
  if (min_bkts >= t.nbkt) {
    unsigned long long n = min_bkts + 1, g = 2ull * t.nbkt;
    return mns_next_bkt(A, n > g ? n : g, &t.next_resize);
  }
  t.next_resize = t.nbkt;
  return 0;
}

MN_HD int* mns_buckets(MnStlTab& t, const MnStlArena& A) { return t.nbkt == 1 ? &t.single : A.bk + t.boff; }

// NP (node policy): unsigned long long key(int node); int next(int node); void set_next(int node, int v)
template <class NP> MN_HD int mns_next_of(const MnStlTab& t, const NP& np, int prev) { return prev == MNS_BB ? t.first : np.next(prev); }
template <class NP> MN_HD void mns_set_next(MnStlTab& t, const NP& np, int prev, int v) {
  if (prev == MNS_BB) t.first = v; else np.set_next(prev, v);
}

// _M_find_before_node: the node before the one holding `key` in bucket bkt
template <class NP> MN_HD int mns_find_before(MnStlTab& t, const MnStlArena& A, const NP& np, unsigned bkt, unsigned long long key) {
  const int* B = mns_buckets(t, A);
  int prev = B[bkt];
  if (prev == MNS_NULL) return MNS_NOTFOUND;
  for (int p = mns_next_of(t, np, prev);; p = np.next(p)) {
    if (np.key(p) == key) return prev;
    const int nx = np.next(p);
    if (nx == MNS_NULL || (unsigned)(np.key(nx) % t.nbkt) != bkt) break;
    prev = p;
  }
  return MNS_NOTFOUND;
}
template <class NP> MN_HD int mns_find(MnStlTab& t, const MnStlArena& A, const NP& np, unsigned long long key) {
  const int prev = mns_find_before(t, A, np, (unsigned)(key % t.nbkt), key);
  return prev == MNS_NOTFOUND ? -1 : mns_next_of(t, np, prev);
}

// n bucket words in the current half; collects into the other half when it is full
MN_HD long long mns_alloc(const MnStlArena& A, unsigned n) {
  if (*A.bump + (long long)n > *A.base + A.half) {
    const long long to = *A.base == 0 ? A.half : 0;
    long long at = to;
    for (int i = 0; i < A.ntabs; i++) {
      MnStlTab& t = A.tabs[i];
      if (t.nbkt <= 1) continue;
      for (unsigned j = 0; j < t.nbkt; j++) A.bk[at + j] = A.bk[t.boff + j];
      t.boff = at;
      at += t.nbkt;
    }
    *A.base = to;
    *A.bump = at;
    (*A.collections)++;
    if (at + (long long)n > to + A.half) { *A.overflow = 1; return -1; }
  }
  const long long off = *A.bump;
  *A.bump = off + n;
  return off;
}

// _M_rehash_aux (unique keys): re-thread the list over n buckets
template <class NP> MN_HD void mns_rehash(MnStlTab& t, const MnStlArena& A, const NP& np, unsigned n) {
  t.nbkt = 1;  // the old bucket array is garbage from here on (only the node list is walked): a collection skips it
  const long long off = mns_alloc(A, n);
  if (off < 0) return;
  int* NB = A.bk + off;
  for (unsigned i = 0; i < n; i++) NB[i] = MNS_NULL;
  int p = t.first;
  t.first = MNS_NULL;
  unsigned bbegin_bkt = 0;
  while (p != MNS_NULL) {
    const int next = np.next(p);
    const unsigned bkt = (unsigned)(np.key(p) % n);
    if (NB[bkt] == MNS_NULL) {
      np.set_next(p, t.first);
      t.first = p;
      NB[bkt] = MNS_BB;
      if (np.next(p) != MNS_NULL) NB[bbegin_bkt] = p;
      bbegin_bkt = bkt;
    } else {
      np.set_next(p, mns_next_of(t, np, NB[bkt]));
      mns_set_next(t, np, NB[bkt], p);
    }
    p = next;
  }
  t.boff = off;
  t.nbkt = n;
}

// operator[] / insert of a key that is NOT in the table (_M_insert_unique_node + _M_insert_bucket_begin);
// np.key(node) must already answer `key`
template <class NP> MN_HD void mns_insert(MnStlTab& t, const MnStlArena& A, const NP& np, unsigned long long key, int node) {
  const unsigned grow = mns_need_rehash(A, t);
  if (grow) mns_rehash(t, A, np, grow);
  int* B = mns_buckets(t, A);
  const unsigned bkt = (unsigned)(key % t.nbkt);
  if (B[bkt] != MNS_NULL) {
    np.set_next(node, mns_next_of(t, np, B[bkt]));
    mns_set_next(t, np, B[bkt], node);
  } else {
    np.set_next(node, t.first);
    t.first = node;
    if (np.next(node) != MNS_NULL) B[(unsigned)(np.key(np.next(node)) % t.nbkt)] = node;
    B[bkt] = MNS_BB;
  }
  t.cnt++;
}

// the table's owner is gone: its bucket array is garbage
MN_HD void mns_tab_drop(MnStlTab& t) { t.nbkt = 1; t.first = MNS_NULL; t.single = MNS_NULL; t.cnt = 0; }

// erase(key): the node that held it, or -1 (_M_erase + _M_remove_bucket_begin)
template <class NP> MN_HD int mns_erase(MnStlTab& t, const MnStlArena& A, const NP& np, unsigned long long key) {
  const unsigned bkt = (unsigned)(key % t.nbkt);
  const int prev = mns_find_before(t, A, np, bkt, key);
  if (prev == MNS_NOTFOUND) return -1;
  int* B = mns_buckets(t, A);
  const int n = mns_next_of(t, np, prev);
  const int nx = np.next(n);
  if (prev == B[bkt]) {  // n opened its bucket
    const unsigned nb = nx != MNS_NULL ? (unsigned)(np.key(nx) % t.nbkt) : 0u;
    if (nx == MNS_NULL || nb != bkt) {  // ... and was alone in it
      if (nx != MNS_NULL) B[nb] = B[bkt];
      if (B[bkt] == MNS_BB) t.first = nx;
      B[bkt] = MNS_NULL;
    }
  } else if (nx != MNS_NULL) {
    const unsigned nb = (unsigned)(np.key(nx) % t.nbkt);
    if (nb != bkt) B[nb] = prev;
  }
  mns_set_next(t, np, prev, nx);
  t.cnt--;
  return n;
}

// ---- std::priority_queue over (priority, record): std::push_heap / std::pop_heap with a comparator that sees the
// priority only ------------------------------------------------------------------------------------------------
struct MnStlHeap {
  float* key;
  int* rec;
  long long n, cap;
};
MN_HD void mns_heap_sift_up(MnStlHeap& h, long long hole, long long top, float vk, int vr) {  // std::__push_heap
  long long parent = (hole - 1) / 2;
  while (hole > top && h.key[parent] < vk) {
    h.key[hole] = h.key[parent]; h.rec[hole] = h.rec[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  h.key[hole] = vk; h.rec[hole] = vr;
}
MN_HD bool mns_heap_push(MnStlHeap& h, float k, int r) {
  if (h.n >= h.cap) return false;
  h.n++;
  mns_heap_sift_up(h, h.n - 1, 0, k, r);
  return true;
}
// top() then pop(): std::pop_heap moves the last element's value down from the root (always towards the larger
// child, the LEFT one among equals) to a leaf, then sifts it back up
MN_HD void mns_heap_pop(MnStlHeap& h, float* k, int* r) {
  *k = h.key[0]; *r = h.rec[0];
  h.n--;
  const long long len = h.n;
  if (len <= 0) return;
  const float vk = h.key[len]; const int vr = h.rec[len];
  long long hole = 0, child = 0;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (h.key[child] < h.key[child - 1]) child--;
    h.key[hole] = h.key[child]; h.rec[hole] = h.rec[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    h.key[hole] = h.key[child - 1]; h.rec[hole] = h.rec[child - 1];
    hole = child - 1;
  }
  mns_heap_sift_up(h, hole, 0, vk, vr);
}
