/*
 * mergenet_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked or imported by the product path).
 *
 * Plain-C, single-threaded CPU restatement of MergeNet's "Mode A" (csegment) greedy merge
 * segmenter.  It follows, function by function, the reference sources
 *     /root/reference/utils/csegment/segment.cc   (cited below as cc:LINE)
 *     /root/reference/utils/csegment/segment.h    (cited below as h:LINE)
 * but restates them on flat arrays:
 *   - the lazy std::priority_queue (h:335, cc:551-565) is kept as a lazy binary heap whose entries
 *     carry a snapshot (mp, lo, hi, rec); an entry is consumed only if it still equals the record's
 *     stored (mp, lo, hi) -- observationally the same "indexed map of stored priorities" as the
 *     reference heap (duplicates of a valid key are no-ops there: cc:554-559);
 *   - ties between equal priorities, which the reference leaves to libstdc++ heap layout and
 *     unordered_map iteration order, are broken deterministically: (mp desc, then a fixed scattering bijection of (lo, hi));
 *   - per-object unordered_map adjacency lists (h:136) become intrusive doubly-linked lists plus
 *     one global (lo,hi)->record hash table.
 * PARITY PIN: the reference ships no golden vectors for this path ("parity unpinned" by its own
 * tests); this file is pinned instead against the reference itself, compiled unmodified into
 * oracle/_ref/libsegment_ref.so by oracle/Makefile, on the inputs in tests/ (masks identical after
 * canonical relabel, identical per-instance classes).
 *
 * Arithmetic: every float op is individually rounded fp32 (build with -ffp-contract=off, no
 * -ffast-math), logf / log(double) / expf are the host libm's, exactly as the reference uses them.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  long long pops;        /* heap pops, valid or not                        (cc:551-553) */
  long long valid_pops;  /* pops whose entry matched the stored priority   (cc:554-559) */
  long long merges;      /* Merge() calls                                  (cc:561-562) */
  long long repushes;    /* pop -> recompute -> changed -> re-push         (cc:563-564) */
  long long pushes;      /* all heap pushes incl. initial                  (cc:226,697,705) */
  long long adj_visits;  /* records visited in Merge's loop                (cc:650-715) */
  long long folds;       /* this_arec folded into that_arec                (cc:685-698) */
  long long init_records;
  long long init_pushes;
  long long sum_abs_npix; /* sum over merges of the absorbed object's pixel count */
  long long max_abs_npix;
  long long merges_abs_gt32, merges_abs_gt1024;
  int final_objects;     /* all surviving objects, incl. class 0 */
  int final_instances;   /* surviving objects with class != 0 */
} mno_stats;

/* ------------------------------------------------------------------------------------------ */
/* libm restatement used by the CUDA kernels; kept here so the CPU suite can pin it to libm.   */
/* Follows the public glibc / ARM optimized-routines logf: all in double.  */
static const double MNO_LOGF_INVC[16] = {
    0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010bp+0,  0x1.3c995b0b80385p+0,
    0x1.30d190c8864a5p+0, 0x1.25e227b0b8eap+0,  0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0,
    0x1.0953f419900a7p+0, 0x1p+0,               0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aap-1,
    0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1};
static const double MNO_LOGF_LOGC[16] = {
    -0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3,
    -0x1.6574f0ac07758p-3, -0x1.1aa2bc79c81p-3,   -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4,
    -0x1.252f438e10c1ep-5, 0x0p+0,                0x1.aa5aa5df25984p-5,  0x1.c5e53aa362eb4p-4,
    0x1.526e57720db08p-3,  0x1.bc2860d22477p-3,   0x1.1058bc8a07ee1p-2,  0x1.4043057b6ee09p-2};

float mno_logf_recipe(float x) {
  uint32_t ix;
  memcpy(&ix, &x, 4);
  uint32_t tmp = ix - 0x3f330000u;
  int i = (tmp >> 19) & 15;
  int k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & 0xff800000u);
  float zf;
  memcpy(&zf, &iz, 4);
  double z = (double)zf;
  double r = z * MNO_LOGF_INVC[i] - 1.0;
  double y0 = MNO_LOGF_LOGC[i] + (double)k * 0x1.62e42fefa39efp-1;
  double r2 = r * r;
  /* A[0]=-0x1.00ea348b88334p-2, A[1]=0x1.5575b0be00b6ap-2, A[2]=-0x1.ffffef20a4123p-2
     y = A[1]*r + A[2]; y = A[0]*r2 + y; y = y*r2 + (y0 + r) */
  double y = 0x1.5575b0be00b6ap-2 * r + -0x1.ffffef20a4123p-2;
  y = -0x1.00ea348b88334p-2 * r2 + y;
  y = y * r2 + (y0 + r);
  return (float)y;
}

/* Count inputs in [lo_bits, hi_bits] (float bit patterns, inclusive) where the recipe differs
 * from the host libm logf.  Exhaustive range of the clipped domain: [0x34000000, 0x3f7ffffe]. */
long long mno_logf_recipe_mismatches(uint32_t lo_bits, uint32_t hi_bits, uint32_t stride) {
  long long bad = 0;
  if (stride == 0) stride = 1;
  for (uint64_t b = lo_bits; b <= hi_bits; b += stride) {
    uint32_t bb = (uint32_t)b;
    float x;
    memcpy(&x, &bb, 4);
    float a = logf(x), c = mno_logf_recipe(x);
    if (memcmp(&a, &c, 4) != 0) bad++;
  }
  return bad;
}

/* Host libm tables for the device exhaustive tests: out[i] = logf(bits lo+i) or
 * (float)log(1.0 - (double)x) (cc:34-35). */
void mno_host_logf_table(uint32_t lo_bits, uint32_t n, float* out) {
  for (uint32_t i = 0; i < n; i++) {
    uint32_t b = lo_bits + i;
    float x;
    memcpy(&x, &b, 4);
    out[i] = logf(x);
  }
}
void mno_host_log1m_table(uint32_t lo_bits, uint32_t n, float* out) {
  for (uint32_t i = 0; i < n; i++) {
    uint32_t b = lo_bits + i;
    float x;
    memcpy(&x, &b, 4);
    out[i] = (float)log(1.0 - (double)x);
  }
}
/* glibc / ARM optimized-routines expf (|x| < 88), the restatement the CUDA kernel uses; the x86-64
 * libm runs its FMA build, hence the three explicit fma() steps.  Pinned to expf() by the CPU suite. */
static const uint64_t MNO_EXP2F_TAB[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
float mno_expf_recipe(float x) {
  const double InvLn2N = 0x1.71547652b82fep+0 * 32.0, SHIFT = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / 32768.0, C1 = 0x1.ebfce50fac4f3p-3 / 1024.0, C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
  double z = InvLn2N * (double)x;
  double kd = z + SHIFT;
  uint64_t ki;
  memcpy(&ki, &kd, 8);
  kd -= SHIFT;
  double r = z - kd;
  uint64_t t = MNO_EXP2F_TAB[ki & 31] + (ki << 47);
  double s;
  memcpy(&s, &t, 8);
  z = fma(C0, r, C1);
  double r2 = r * r;
  double y = fma(C2, r, 1.0);
  y = fma(z, r2, y);
  y = y * s;
  return (float)y;
}
/* mismatches of the recipe against the host expf over +-[2^-20 .. 2^6.3] sampled with `stride` */
long long mno_expf_recipe_mismatches(uint32_t stride) {
  long long bad = 0;
  for (uint32_t b = 0x35800000u; b < 0x42a00000u; b += stride) {
    float x;
    memcpy(&x, &b, 4);
    for (int sg = 0; sg < 2; sg++) {
      float v = sg ? -x : x;
      float a = expf(v), c = mno_expf_recipe(v);
      if (memcmp(&a, &c, 4) != 0) bad++;
    }
  }
  return bad;
}
/* same_different_bias transform, cc:183-195 */
static float mno_bias_sameness(float s, float sdb) {
  float logit = (float)((double)logf(s) - log(1.0 - (double)s) + (double)sdb);
  return (float)(1.0 / (1.0 + (double)expf(-logit)));
}
void mno_host_bias_table(uint32_t lo_bits, uint32_t n, float sdb, float* out) {
  for (uint32_t i = 0; i < n; i++) {
    uint32_t b = lo_bits + i;
    float x;
    memcpy(&x, &b, 4);
    out[i] = mno_bias_sameness(x, sdb);
  }
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
  float mp;
  int lo, hi, rec;
} heap_ent;

/* a pops before b?  (mp desc, tie(lo, hi) asc), tie = (u, D) lexicographic with D = hi - lo and
 * u = (bitrev24(lo) + 0x9E3779 * D) mod 2^24 -- a bijection of the pair, hence a total order.  The
 * reference leaves ties to libstdc++ heap layout / unordered_map order; this fixed rule scatters equal
 * priorities over the image and over the records of one object (the CUDA scheduler uses the same one,
 * mn_common.h). */
static inline uint64_t mno_tie(int lo, int hi) {
  uint32_t v = (uint32_t)lo, r = 0;
  for (int i = 0; i < 24; i++) { r = (r << 1) | (v & 1u); v >>= 1; }
  uint32_t D = (uint32_t)(hi - lo);
  uint32_t u = (r + 0x9E3779u * D) & 0xFFFFFFu;
  return ((uint64_t)u << 24) | D;
}
static inline int ent_before(const heap_ent* a, const heap_ent* b) {
  if (a->mp != b->mp) return a->mp > b->mp;
  return mno_tie(a->lo, a->hi) < mno_tie(b->lo, b->hi);
}

typedef struct {
  int H, W, C, K, N;
  long long E; /* N*K record slots, rec = pixel*K + k */
  float omf, mlb;
  /* objects */
  int* npix;
  int* cls;
  float* same_obj;
  float* clp; /* N*C */
  char* oalive;
  int* pix_next; /* per-object pixel chains */
  int* pix_tail;
  int* adj_head; /* head record of the object's adjacency list, -1 if empty */
  /* records */
  int* lo;
  int* hi;
  float* oml;
  float* same;
  float* diff;
  float* mp;
  char* state; /* 0 = never existed, 1 = live, 2 = folded into another (cc:694), 3 = merged (cc:726) */
  int* lnext[2]; /* side 0: list of lo, side 1: list of hi */
  int* lprev[2];
  /* (lo,hi) -> rec hash, open addressing, backward-shift delete */
  uint64_t hcap;
  int* hslot;
  /* lazy heap */
  heap_ent* heap;
  long long hn, hcapacity;
  mno_stats st;
  /* ---- schedule analysis (mno_analyze_rounds): greedy split of the sequential event order
   * into rounds of mutually independent events that all existed at round start ---- */
  int an_on;
  long long an_cap, an_rounds, an_round_events, an_events;
  long long an_cut_conflict, an_cut_cascade, an_cut_cap;
  int an_cur;          /* current round id */
  int* an_lastw;       /* per object: round id of last write */
  int* an_lastr;       /* per object: round id of last read */
  int* an_stored_round; /* per record: round id in which its stored priority was last written */
  long long an_hist[16]; /* round-size histogram, bucket = floor(log2(size)) */
} seg_t;

static inline uint64_t hkey(int lo, int hi) { return ((uint64_t)(uint32_t)lo << 32) | (uint32_t)hi; }
static inline uint64_t hmix(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}
static int hash_find(seg_t* s, int lo, int hi) {
  uint64_t m = s->hcap - 1, i = hmix(hkey(lo, hi)) & m;
  for (;;) {
    int r = s->hslot[i];
    if (r < 0) return -1;
    if (s->lo[r] == lo && s->hi[r] == hi) return r;
    i = (i + 1) & m;
  }
}
static void hash_insert(seg_t* s, int r) {
  uint64_t m = s->hcap - 1, i = hmix(hkey(s->lo[r], s->hi[r])) & m;
  while (s->hslot[i] >= 0) i = (i + 1) & m;
  s->hslot[i] = r;
}
/* delete the slot holding record r (its lo/hi must still be the keyed values) */
static void hash_erase(seg_t* s, int r) {
  uint64_t m = s->hcap - 1, i = hmix(hkey(s->lo[r], s->hi[r])) & m;
  while (s->hslot[i] != r) {
    if (s->hslot[i] < 0) return;
    i = (i + 1) & m;
  }
  uint64_t j = i;
  for (;;) {
    j = (j + 1) & m;
    int q = s->hslot[j];
    if (q < 0) break;
    uint64_t h = hmix(hkey(s->lo[q], s->hi[q])) & m;
    /* can q move to i?  yes iff its home h is cyclically not in (i, j] */
    int in_range = (i <= j) ? (h > i && h <= j) : (h > i || h <= j);
    if (!in_range) {
      s->hslot[i] = q;
      i = j;
    }
  }
  s->hslot[i] = -1;
}

static void heap_push(seg_t* s, float mp, int lo, int hi, int rec) {
  if (s->hn == s->hcapacity) {
    s->hcapacity = s->hcapacity * 2 + 1024;
    s->heap = (heap_ent*)realloc(s->heap, sizeof(heap_ent) * (size_t)s->hcapacity);
  }
  heap_ent e = {mp, lo, hi, rec};
  long long i = s->hn++;
  while (i > 0) {
    long long p = (i - 1) >> 1;
    if (!ent_before(&e, &s->heap[p])) break;
    s->heap[i] = s->heap[p];
    i = p;
  }
  s->heap[i] = e;
  s->st.pushes++;
}
static heap_ent heap_pop(seg_t* s) {
  heap_ent top = s->heap[0];
  heap_ent e = s->heap[--s->hn];
  long long i = 0, n = s->hn;
  for (;;) {
    long long c = 2 * i + 1;
    if (c >= n) break;
    if (c + 1 < n && ent_before(&s->heap[c + 1], &s->heap[c])) c++;
    if (!ent_before(&s->heap[c], &e)) break;
    s->heap[i] = s->heap[c];
    i = c;
  }
  if (n > 0) s->heap[i] = e;
  return top;
}

/* adjacency list helpers */
static inline int side_of(seg_t* s, int r, int o) { return s->lo[r] == o ? 0 : 1; }
static void list_add(seg_t* s, int o, int r) {
  int sd = side_of(s, r, o);
  int h = s->adj_head[o];
  s->lnext[sd][r] = h;
  s->lprev[sd][r] = -1;
  if (h >= 0) s->lprev[side_of(s, h, o)][h] = r;
  s->adj_head[o] = r;
}
static void list_del(seg_t* s, int o, int r) {
  int sd = side_of(s, r, o);
  int nx = s->lnext[sd][r], pv = s->lprev[sd][r];
  if (pv >= 0)
    s->lnext[side_of(s, pv, o)][pv] = nx;
  else
    s->adj_head[o] = nx;
  if (nx >= 0) s->lprev[side_of(s, nx, o)][nx] = pv;
}

/* cc:107-150: ComputeClassDeltaLogprob + UpdateMergePriority.  o1 = lower id. */
static float compute_priority(seg_t* s, int r, int* merged_class) {
  int o1 = s->lo[r], o2 = s->hi[r];
  float cdl;
  int merged;
  if (s->cls[o1] == s->cls[o2]) { /* cc:108,120-121 */
    cdl = 0.0f;
    merged = s->cls[o1];
  } else { /* cc:123-140 */
    const float* c1 = s->clp + (size_t)o1 * s->C;
    const float* c2 = s->clp + (size_t)o2 * s->C;
    float best = c1[0] + c2[0];
    merged = 0;
    for (int c = 1; c < s->C; c++) {
      float j = c1[c] + c2[c];
      if (j > best) { /* max_element: first maximum */
        best = j;
        merged = c;
      }
    }
    cdl = best - c1[s->cls[o1]];
    cdl = cdl - c2[s->cls[o2]];
  }
  if (merged_class) *merged_class = merged;
  size_t den = (size_t)s->npix[o1] + (size_t)s->npix[o2]; /* cc:147 */
  float num = s->oml[r] * s->omf;                         /* cc:148 */
  num = num + cdl;
  float mp = num / (float)den;
  mp = mp + s->mlb; /* cc:149 */
  return mp;
}

static void seg_free(seg_t* s) {
  free(s->npix); free(s->cls); free(s->same_obj); free(s->clp); free(s->oalive);
  free(s->pix_next); free(s->pix_tail); free(s->adj_head);
  free(s->lo); free(s->hi); free(s->oml); free(s->same); free(s->diff); free(s->mp); free(s->state);
  free(s->lnext[0]); free(s->lnext[1]); free(s->lprev[0]); free(s->lprev[1]);
  free(s->hslot); free(s->heap);
}

/* cc:153-232: constructor.  Clipping is the caller's job (c_segment.pyx:53-55). */
static int seg_init(seg_t* s, float* class_pred, int class_dim, float* adj_pred, int offset_dim,
                    int W, int H, int num_classes, const int* offsets, float sdb, float omf,
                    float mlb, int build_lists) {
  memset(s, 0, sizeof(*s));
  (void)class_dim;
  s->H = H; s->W = W; s->C = num_classes; s->K = offset_dim; s->N = H * W;
  s->E = (long long)s->N * s->K;
  s->omf = omf; s->mlb = mlb;
  int N = s->N, C = s->C, K = s->K;
  size_t E = (size_t)s->E;
  s->npix = (int*)malloc(sizeof(int) * N);
  s->cls = (int*)malloc(sizeof(int) * N);
  s->same_obj = (float*)calloc(N, sizeof(float));
  s->clp = (float*)malloc(sizeof(float) * (size_t)N * C);
  s->oalive = (char*)malloc(N);
  s->pix_next = (int*)malloc(sizeof(int) * N);
  s->pix_tail = (int*)malloc(sizeof(int) * N);
  s->adj_head = (int*)malloc(sizeof(int) * N);
  s->lo = (int*)malloc(sizeof(int) * E);
  s->hi = (int*)malloc(sizeof(int) * E);
  s->oml = (float*)malloc(sizeof(float) * E);
  s->same = (float*)malloc(sizeof(float) * E);
  s->diff = (float*)malloc(sizeof(float) * E);
  s->mp = (float*)malloc(sizeof(float) * E);
  s->state = (char*)calloc(E, 1);
  if (build_lists) {
    for (int i = 0; i < 2; i++) {
      s->lnext[i] = (int*)malloc(sizeof(int) * E);
      s->lprev[i] = (int*)malloc(sizeof(int) * E);
    }
    s->hcap = 1;
    while (s->hcap < 2 * E + 16) s->hcap <<= 1;
    s->hslot = (int*)malloc(sizeof(int) * s->hcap);
    memset(s->hslot, 0xff, sizeof(int) * s->hcap);
  }
  /* cc:183-195: optional same/different bias, IN PLACE on the caller's buffer */
  if (sdb != 0) {
    for (size_t i = 0; i < E; i++) adj_pred[i] = mno_bias_sameness(adj_pred[i], sdb);
  }
  /* cc:196-207 + Object ctor cc:5-21 */
  for (int p = 0; p < N; p++) {
    float best = 0;
    int bc = 0;
    for (int c = 0; c < C; c++) {
      float v = 0.0f;
      v += logf(class_pred[(size_t)c * N + p]); /* h:295-297 */
      s->clp[(size_t)p * C + c] = v;
      if (c == 0 || v > best) {
        best = v;
        bc = c;
      }
    }
    s->cls[p] = bc;
    s->npix[p] = 1;
    s->oalive[p] = 1;
    s->pix_next[p] = -1;
    s->pix_tail[p] = p;
    s->adj_head[p] = -1;
  }
  /* cc:209-231 + AdjacencyRecord ctor cc:24-46 */
  for (int row = 0; row < H; row++) {
    for (int col = 0; col < W; col++) {
      int p = row * W + col;
      for (int k = 0; k < K; k++) {
        int r2 = row + offsets[2 * k], c2 = col + offsets[2 * k + 1];
        if (r2 < 0 || r2 >= H || c2 < 0 || c2 >= W) continue;
        int q = r2 * W + c2;
        int r = p * K + k;
        float sp = adj_pred[(size_t)k * N + p];   /* h:301-303: read at the SOURCE pixel */
        s->diff[r] = (float)log(1.0 - (double)sp); /* cc:34 */
        s->same[r] = logf(sp);                     /* cc:35 */
        s->oml[r] = s->same[r] - s->diff[r];       /* cc:36 */
        s->lo[r] = p < q ? p : q;                  /* cc:49-56 */
        s->hi[r] = p < q ? q : p;
        s->state[r] = 1;
        s->st.init_records++;
        s->mp[r] = compute_priority(s, r, NULL); /* cc:45 */
        if (build_lists) {
          list_add(s, s->lo[r], r);
          list_add(s, s->hi[r], r);
          hash_insert(s, r);
          if (s->mp[r] >= 0) { /* cc:225-227 */
            heap_push(s, s->mp[r], s->lo[r], s->hi[r], r);
            s->st.init_pushes++;
          }
        }
      }
    }
  }
  return 0;
}

static void an_close_round(seg_t* s) {
  if (s->an_round_events > 0) {
    int b = 0;
    long long v = s->an_round_events;
    while (v > 1 && b < 15) { v >>= 1; b++; }
    s->an_hist[b]++;
    s->an_rounds++;
  }
  s->an_round_events = 0;
  s->an_cur++;
}
/* called before an event on record r executes; is_merge: o_abs is the absorbed object */
static void an_event(seg_t* s, int r, int is_merge, int o_surv, int o_abs) {
  int cut = 0;
  if (s->an_stored_round[r] == s->an_cur) { cut = 1; s->an_cut_cascade++; }
  else if (s->an_round_events >= s->an_cap) { cut = 1; s->an_cut_cap++; }
  else {
    int a = s->lo[r], b = s->hi[r], c = 0;
    if (s->an_lastw[a] == s->an_cur || s->an_lastw[b] == s->an_cur) c = 1;
    if (!c && is_merge) {
      if (s->an_lastr[a] == s->an_cur || s->an_lastr[b] == s->an_cur) c = 1;
      for (int t = s->adj_head[o_abs]; t >= 0 && !c; t = s->lnext[side_of(s, t, o_abs)][t]) {
        int o3 = s->lo[t] == o_abs ? s->hi[t] : s->lo[t];
        if (o3 != o_surv && s->an_lastw[o3] == s->an_cur) c = 1;
      }
    }
    if (c) { cut = 1; s->an_cut_conflict++; }
  }
  if (cut) an_close_round(s);
  s->an_round_events++;
  s->an_events++;
  int a = s->lo[r], b = s->hi[r];
  if (is_merge) {
    s->an_lastw[a] = s->an_cur; s->an_lastw[b] = s->an_cur;
    for (int t = s->adj_head[o_abs]; t >= 0; t = s->lnext[side_of(s, t, o_abs)][t]) {
      int o3 = s->lo[t] == o_abs ? s->hi[t] : s->lo[t];
      s->an_lastr[o3] = s->an_cur;
      s->an_stored_round[t] = s->an_cur; /* t (or the record it folds into) is re-stored */
      if (o3 != o_surv) {
        int nlo = o_surv < o3 ? o_surv : o3, nhi = o_surv < o3 ? o3 : o_surv;
        int u = hash_find(s, nlo, nhi);
        if (u >= 0) s->an_stored_round[u] = s->an_cur;
      }
    }
  } else {
    s->an_lastr[a] = s->an_cur; s->an_lastr[b] = s->an_cur;
    s->an_stored_round[r] = s->an_cur;
  }
}

/* cc:602-727 */
static void seg_merge(seg_t* s, int rec, int merged_class, int* merge_log, long long log_cap) {
  int o1 = s->lo[rec], o2 = s->hi[rec];
  if (s->npix[o1] < s->npix[o2]) { /* cc:612-616: lower id survives on equal sizes */
    int t = o1; o1 = o2; o2 = t;
  }
  if (merge_log && s->st.merges < log_cap) {
    merge_log[2 * s->st.merges] = o1;
    merge_log[2 * s->st.merges + 1] = o2;
  }
  s->st.merges++;
  s->st.sum_abs_npix += s->npix[o2];
  if (s->npix[o2] > s->st.max_abs_npix) s->st.max_abs_npix = s->npix[o2];
  if (s->npix[o2] > 32) s->st.merges_abs_gt32++;
  if (s->npix[o2] > 1024) s->st.merges_abs_gt1024++;
  s->cls[o1] = merged_class;                 /* cc:635 */
  s->pix_next[s->pix_tail[o1]] = o2;         /* cc:636-639 (pixel-set union) */
  s->pix_tail[o1] = s->pix_tail[o2];
  s->npix[o1] += s->npix[o2];
  float* c1 = s->clp + (size_t)o1 * s->C;
  const float* c2 = s->clp + (size_t)o2 * s->C;
  for (int c = 0; c < s->C; c++) c1[c] += c2[c]; /* cc:640, h:95-106 */
  s->same_obj[o1] += (s->same[rec] + s->same_obj[o2]); /* cc:641-642 */
  /* cc:645-647 */
  hash_erase(s, rec);
  list_del(s, o1, rec);
  list_del(s, o2, rec);
  s->state[rec] = 3;
  /* cc:650-715 */
  int t = s->adj_head[o2];
  while (t >= 0) {
    int sd2 = side_of(s, t, o2);
    int nxt = s->lnext[sd2][t];
    int o3 = sd2 == 0 ? s->hi[t] : s->lo[t];
    s->st.adj_visits++;
    hash_erase(s, t);       /* cc:680 (old key) */
    list_del(s, o3, t);     /* cc:681 */
    int nlo = o1 < o3 ? o1 : o3, nhi = o1 < o3 ? o3 : o1; /* cc:659-664,677 */
    int u = hash_find(s, nlo, nhi);                       /* cc:685-686 */
    if (u >= 0) {
      s->oml[u] += s->oml[t];   /* cc:690 */
      s->diff[u] += s->diff[t]; /* cc:691 */
      s->same[u] += s->same[t]; /* cc:692 */
      s->state[t] = 2;          /* cc:694 */
      s->mp[t] = FLT_MIN;
      s->lo[t] = nlo; s->hi[t] = nhi;
      s->st.folds++;
      s->mp[u] = compute_priority(s, u, NULL); /* cc:695 */
      if (s->mp[u] >= 0) heap_push(s, s->mp[u], s->lo[u], s->hi[u], u); /* cc:696-698 */
    } else {
      s->lo[t] = nlo; s->hi[t] = nhi;
      list_add(s, o1, t); /* cc:700-702 */
      list_add(s, o3, t);
      hash_insert(s, t);
      s->mp[t] = compute_priority(s, t, NULL); /* cc:703 */
      if (s->mp[t] >= 0) heap_push(s, s->mp[t], s->lo[t], s->hi[t], t); /* cc:704-706 */
    }
    t = nxt;
  }
  s->adj_head[o2] = -1;
  s->oalive[o2] = 0; /* cc:724-726 */
}

/* cc:539-573 */
static void seg_run(seg_t* s, int* merge_log, long long log_cap) {
  while (s->hn > 0) {
    heap_ent e = heap_pop(s);
    s->st.pops++;
    int r = e.rec;
    /* cc:554-559 restated on the indexed-map view: the entry must equal the stored state */
    if (s->state[r] != 1 || e.mp != s->mp[r] || e.lo != s->lo[r] || e.hi != s->hi[r]) continue;
    s->st.valid_pops++;
    int merged;
    float mp = compute_priority(s, r, &merged); /* cc:560 */
    s->mp[r] = mp;
    if (s->an_on) {
      int is_m = (mp == e.mp);
      int a = s->lo[r], b = s->hi[r];
      int surv = s->npix[a] < s->npix[b] ? b : a;
      an_event(s, r, is_m, surv, surv == a ? b : a);
    }
    if (mp == e.mp) { /* cc:561-562 */
      seg_merge(s, r, merged, merge_log, log_cap);
    } else if (mp >= 0) { /* cc:563-565 */
      heap_push(s, mp, s->lo[r], s->hi[r], r);
      s->st.repushes++;
    }
  }
}

/* cc:491-517; labels in ascending surviving-object id (the reference's order is unordered_map
 * iteration order, i.e. arbitrary; compare after canonical relabel). */
static void seg_output(seg_t* s, int* output, int* object_class) {
  int N = s->N;
  for (int i = 0; i < N; i++) output[i] = 0;
  for (int i = 0; i < N; i++) object_class[i] = -1;
  int k = 1;
  s->st.final_objects = 0;
  for (int o = 0; o < N; o++) {
    if (!s->oalive[o]) continue;
    s->st.final_objects++;
    if (s->cls[o] == 0) continue; /* cc:506-508 */
    object_class[k - 1] = s->cls[o];
    for (int p = o; p >= 0; p = s->pix_next[p]) output[p] = k;
    k++;
  }
  s->st.final_instances = k - 1;
}

/* Same contract as the reference C ABI (cc:742-752) plus optional stats / merge log.
 * Returns 0. */
int mno_run_segmentation(float* class_pred, int class_dim, float* adj_pred, int offset_dim,
                         int img_width, int img_height, int num_classes, const int* offset_list,
                         int* output, int* object_class, float same_different_bias,
                         float object_merge_factor, float merge_logprob_bias, mno_stats* stats,
                         int* merge_log, long long merge_log_cap) {
  seg_t s;
  seg_init(&s, class_pred, class_dim, adj_pred, offset_dim, img_width, img_height, num_classes,
           offset_list, same_different_bias, object_merge_factor, merge_logprob_bias, 1);
  seg_run(&s, merge_log, merge_log_cap);
  seg_output(&s, output, object_class);
  if (stats) *stats = s.st;
  seg_free(&s);
  return 0;
}

/* Dump of the constructor's results (cc:153-232) for edge-pass parity tests.
 *   clp[N*C] (pixel-major), cls[N]; per record slot r = pixel*K + k: same, diff, oml, mp (0 when
 *   the offset leaves the image) and valid[r]. */
/* cc:272-287 (ComputeTotalLogprob): sum over surviving objects of their class log-prob, plus
 * object_merge_factor * (sum over surviving records of differentness + sum over objects of the
 * sameness inside them).  The per-object / per-record sums are the fp32 accumulators the merges
 * maintained; they are added up here in double, in index order (the reference prints a float sum in
 * hash-map order).  totals = {class term, object sameness term, record differentness term, total}. */
int mno_run_segmentation_totals(float* class_pred, int class_dim, float* adj_pred, int offset_dim,
                                int img_width, int img_height, int num_classes,
                                const int* offset_list, int* output, int* object_class,
                                float same_different_bias, float object_merge_factor,
                                float merge_logprob_bias, double* totals) {
  seg_t s;
  seg_init(&s, class_pred, class_dim, adj_pred, offset_dim, img_width, img_height, num_classes,
           offset_list, same_different_bias, object_merge_factor, merge_logprob_bias, 1);
  seg_run(&s, NULL, 0);
  seg_output(&s, output, object_class);
  double tc = 0.0, ts = 0.0, td = 0.0;
  for (int o = 0; o < s.N; o++) {
    if (!s.oalive[o]) continue;
    tc += (double)s.clp[(size_t)o * s.C + s.cls[o]];
    ts += (double)s.same_obj[o];
  }
  for (long long r = 0; r < s.E; r++)
    if (s.state[r] == 1) td += (double)s.diff[r];
  totals[0] = tc; totals[1] = ts; totals[2] = td;
  totals[3] = tc + (double)object_merge_factor * (td + ts);
  seg_free(&s);
  return 0;
}

int mno_init_dump(float* class_pred, int class_dim, float* adj_pred, int offset_dim, int img_width,
                  int img_height, int num_classes, const int* offset_list,
                  float same_different_bias, float object_merge_factor, float merge_logprob_bias,
                  float* clp, int* cls, float* same, float* diff, float* oml, float* mp,
                  unsigned char* valid) {
  seg_t s;
  seg_init(&s, class_pred, class_dim, adj_pred, offset_dim, img_width, img_height, num_classes,
           offset_list, same_different_bias, object_merge_factor, merge_logprob_bias, 0);
  memcpy(clp, s.clp, sizeof(float) * (size_t)s.N * s.C);
  memcpy(cls, s.cls, sizeof(int) * s.N);
  for (long long r = 0; r < s.E; r++) {
    int v = s.state[r] == 1;
    valid[r] = (unsigned char)v;
    same[r] = v ? s.same[r] : 0.0f;
    diff[r] = v ? s.diff[r] : 0.0f;
    oml[r] = v ? s.oml[r] : 0.0f;
    mp[r] = v ? s.mp[r] : 0.0f;
  }
  seg_free(&s);
  return 0;
}

/* Schedule analysis (design aid, not part of the parity oracle): how many rounds does the
 * sequential event order split into when a round may hold at most `cap` events that (a) all
 * had their stored priority before the round began and (b) are pairwise independent
 * (no object written by one is read or written by another)?
 * out[0]=events out[1]=rounds out[2]=cuts by conflict out[3]=cuts by cascade out[4]=cuts by cap
 * out[5..20]=histogram of round sizes by floor(log2). */
int mno_analyze_rounds(float* class_pred, int class_dim, float* adj_pred, int offset_dim,
                       int img_width, int img_height, int num_classes, const int* offset_list,
                       float same_different_bias, float object_merge_factor,
                       float merge_logprob_bias, long long cap, long long* out) {
  seg_t s;
  seg_init(&s, class_pred, class_dim, adj_pred, offset_dim, img_width, img_height, num_classes,
           offset_list, same_different_bias, object_merge_factor, merge_logprob_bias, 1);
  s.an_on = 1;
  s.an_cap = cap;
  s.an_cur = 1;
  s.an_lastw = (int*)calloc(s.N, sizeof(int));
  s.an_lastr = (int*)calloc(s.N, sizeof(int));
  s.an_stored_round = (int*)calloc((size_t)s.E, sizeof(int));
  seg_run(&s, NULL, 0);
  an_close_round(&s);
  out[0] = s.an_events; out[1] = s.an_rounds; out[2] = s.an_cut_conflict;
  out[3] = s.an_cut_cascade; out[4] = s.an_cut_cap;
  for (int i = 0; i < 16; i++) out[5 + i] = s.an_hist[i];
  free(s.an_lastw); free(s.an_lastr); free(s.an_stored_round);
  seg_free(&s);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Round-schedule model (design validation for the CUDA scheduler; still test infrastructure).
 *
 * A round takes the next `cap` valid queue entries in pop order, PLANS every one of them against
 * the round-start state only (what a GPU worker could read before any commit), accepts the
 * longest prefix whose members are pairwise independent and none of which is overtaken by a
 * priority created by an earlier member, then COMMITS the accepted events one by one with the
 * sequential code above and checks that what the sequential code stored equals the plan.
 * A mismatch means the independence rule is wrong.  Result must equal mno_run_segmentation. */
typedef struct { int t, u, x; float oml, same, diff, mp; int lo, hi; } plan_rec;
typedef struct {
  heap_ent e;
  int kind; /* 1 = re-store, 2 = merge */
  float newmp; int merged, surv, absd;
  int nrec, roff;
  int has_new; heap_ent maxnew;
} plan_cand;

static float prio_explicit(const seg_t* s, float oml, int n1, int cls1, const float* c1, int n2,
                           int cls2, const float* c2, int* merged_out) {
  float cdl; int merged;
  if (cls1 == cls2) { cdl = 0.0f; merged = cls1; }
  else {
    float best = c1[0] + c2[0]; merged = 0;
    for (int c = 1; c < s->C; c++) { float j = c1[c] + c2[c]; if (j > best) { best = j; merged = c; } }
    cdl = best - c1[cls1]; cdl = cdl - c2[cls2];
  }
  if (merged_out) *merged_out = merged;
  size_t den = (size_t)n1 + (size_t)n2;
  float num = oml * s->omf; num = num + cdl;
  float mp = num / (float)den; mp = mp + s->mlb;
  return mp;
}

long long mno_model_mismatches;

int mno_run_rounds_model(float* class_pred, int class_dim, float* adj_pred, int offset_dim,
                         int img_width, int img_height, int num_classes, const int* offset_list,
                         int* output, int* object_class, float same_different_bias,
                         float object_merge_factor, float merge_logprob_bias, int cap,
                         long long* out_rounds, long long* out_events) {
  seg_t S; seg_t* s = &S;
  seg_init(s, class_pred, class_dim, adj_pred, offset_dim, img_width, img_height, num_classes,
           offset_list, same_different_bias, object_merge_factor, merge_logprob_bias, 1);
  int C = s->C;
  plan_cand* cand = (plan_cand*)malloc(sizeof(plan_cand) * cap);
  size_t prcap = 1 << 16; plan_rec* pr = (plan_rec*)malloc(sizeof(plan_rec) * prcap);
  float* tmpclp = (float*)malloc(sizeof(float) * C);
  int* wstamp = (int*)calloc(s->N, sizeof(int)); /* round id in which an accepted event writes o */
  int* rstamp = (int*)calloc(s->N, sizeof(int));
  long long rounds = 0, events = 0; int rid = 0;
  mno_model_mismatches = 0;
  while (s->hn > 0) {
    /* ---- select: next `cap` valid entries in pop order ---- */
    int nc = 0;
    while (nc < cap && s->hn > 0) {
      heap_ent e = heap_pop(s); s->st.pops++;
      int r = e.rec;
      if (s->state[r] != 1 || e.mp != s->mp[r] || e.lo != s->lo[r] || e.hi != s->hi[r]) continue;
      int dup = 0; for (int i = 0; i < nc; i++) if (cand[i].e.rec == r) dup = 1;
      if (dup) continue;
      cand[nc++].e = e;
    }
    if (nc == 0) break;
    rid++; rounds++;
    /* ---- plan (reads round-start state only) ---- */
    size_t npr = 0;
    for (int i = 0; i < nc; i++) {
      plan_cand* c = &cand[i]; int r = c->e.rec;
      c->has_new = 0; c->nrec = 0; c->roff = (int)npr;
      c->newmp = compute_priority(s, r, &c->merged);
      if (c->newmp != c->e.mp) {
        c->kind = 1;
        if (c->newmp >= 0) { c->has_new = 1; c->maxnew.mp = c->newmp; c->maxnew.lo = s->lo[r]; c->maxnew.hi = s->hi[r]; c->maxnew.rec = r; }
        continue;
      }
      c->kind = 2;
      int a = s->lo[r], b = s->hi[r];
      if (s->npix[a] < s->npix[b]) { int t = a; a = b; b = t; }
      c->surv = a; c->absd = b;
      int na = s->npix[a] + s->npix[b];
      for (int k = 0; k < C; k++) tmpclp[k] = s->clp[(size_t)a * C + k] + s->clp[(size_t)b * C + k];
      for (int t = s->adj_head[b]; t >= 0; t = s->lnext[side_of(s, t, b)][t]) {
        if (t == r) continue;
        int x = s->lo[t] == b ? s->hi[t] : s->lo[t];
        int nlo = a < x ? a : x, nhi = a < x ? x : a;
        int u = hash_find(s, nlo, nhi);
        if (npr == prcap) { prcap *= 2; pr = (plan_rec*)realloc(pr, sizeof(plan_rec) * prcap); }
        plan_rec* q = &pr[npr++]; c->nrec++;
        q->t = t; q->u = u; q->x = x; q->lo = nlo; q->hi = nhi;
        if (u >= 0) { q->oml = s->oml[u] + s->oml[t]; q->diff = s->diff[u] + s->diff[t]; q->same = s->same[u] + s->same[t]; }
        else { q->oml = s->oml[t]; q->diff = s->diff[t]; q->same = s->same[t]; }
        const float* cx = s->clp + (size_t)x * C;
        if (a < x) q->mp = prio_explicit(s, q->oml, na, c->merged, tmpclp, s->npix[x], s->cls[x], cx, NULL);
        else q->mp = prio_explicit(s, q->oml, s->npix[x], s->cls[x], cx, na, c->merged, tmpclp, NULL);
        if (q->mp >= 0) {
          heap_ent ne = {q->mp, nlo, nhi, u >= 0 ? u : t};
          if (!c->has_new || ent_before(&ne, &c->maxnew)) { c->has_new = 1; c->maxnew = ne; }
        }
      }
    }
    /* ---- accept the longest valid prefix ---- */
    int nacc = 0; int have_max = 0; heap_ent runmax;
    for (int i = 0; i < nc; i++) {
      plan_cand* c = &cand[i]; int ok = 1;
      if (have_max && ent_before(&runmax, &c->e)) ok = 0; /* a created entry pops first */
      int a = s->lo[c->e.rec], b = s->hi[c->e.rec];
      if (ok && (wstamp[a] == rid || wstamp[b] == rid)) ok = 0;
      if (ok && c->kind == 2) {
        if (rstamp[a] == rid || rstamp[b] == rid) ok = 0;
        for (int k = 0; ok && k < c->nrec; k++) if (wstamp[pr[c->roff + k].x] == rid) ok = 0;
      }
      if (!ok) break;
      nacc++;
      if (c->kind == 2) {
        wstamp[a] = rid; wstamp[b] = rid;
        for (int k = 0; k < c->nrec; k++) rstamp[pr[c->roff + k].x] = rid;
      } else { rstamp[a] = rid; rstamp[b] = rid; }
      if (c->has_new && (!have_max || ent_before(&c->maxnew, &runmax))) { have_max = 1; runmax = c->maxnew; }
    }
    /* ---- push back the rest untouched ---- */
    for (int i = nacc; i < nc; i++) { heap_push(s, cand[i].e.mp, cand[i].e.lo, cand[i].e.hi, cand[i].e.rec); s->st.pushes--; }
    /* ---- commit with the sequential code, check against the plan ---- */
    for (int i = 0; i < nacc; i++) {
      plan_cand* c = &cand[i]; int r = c->e.rec; events++;
      s->st.valid_pops++;
      int merged; float mp = compute_priority(s, r, &merged);
      if (mp != c->newmp || merged != c->merged) mno_model_mismatches++;
      s->mp[r] = mp;
      if (c->kind == 2) {
        if (mp != c->e.mp) mno_model_mismatches++;
        seg_merge(s, r, merged, NULL, 0);
        for (int k = 0; k < c->nrec; k++) {
          plan_rec* q = &pr[c->roff + k]; int w = q->u >= 0 ? q->u : q->t;
          if (s->state[w] != 1 || s->mp[w] != q->mp || s->oml[w] != q->oml || s->same[w] != q->same ||
              s->diff[w] != q->diff || s->lo[w] != q->lo || s->hi[w] != q->hi) mno_model_mismatches++;
          if (q->u >= 0 && s->state[q->t] != 2) mno_model_mismatches++;
        }
      } else {
        if (mp == c->e.mp) mno_model_mismatches++;
        if (mp >= 0) { heap_push(s, mp, s->lo[r], s->hi[r], r); s->st.repushes++; }
      }
    }
  }
  seg_output(s, output, object_class);
  if (out_rounds) *out_rounds = rounds;
  if (out_events) *out_events = events;
  free(cand); free(pr); free(tmpclp); free(wstamp); free(rstamp);
  seg_free(s);
  return (int)(mno_model_mismatches != 0);
}

/* ================================================================================================
 * The step after the path: checkers for mergenet_b200/csrc/mn_post.cuh.
 *
 * mno_resize_nearest: cv2.resize(mask, (ow, oh), interpolation=cv2.INTER_NEAREST) as OpenCV's resizeNN
 * computes it (egs/cityscape/local/segment.py:147-149): sx = min(cvFloor(x * (1 / (ow / (double)w))), w-1).
 * Pinned against cv2 itself by tests/test_post_oracle.py (cv2 is in this image).
 *
 * mno_coco_rle: maskUtils.encode(np.asfortranarray(mask == i)) for i = 1..n
 * (egs/cityscape/local/segment.py:165-186).  pycocotools is NOT in /root/reference nor in this image and
 * the reference pins no version: restated from the published cocoapi common/maskApi.c (rleEncode,
 * rleToString, rleFrString, rleDecode).  PARITY UNPINNED against pycocotools itself; pinned only by the
 * encode -> decode round trip (mno_coco_rle_decode) and hand-checked strings in the tests.
 * ================================================================================================ */
void mno_resize_nearest(const int* in, int h, int w, int* out, int oh, int ow) {
  const double ifx = 1.0 / ((double)ow / (double)w), ify = 1.0 / ((double)oh / (double)h);
  for (int y = 0; y < oh; y++) {
    int sy = (int)floor(y * ify);
    if (sy > h - 1) sy = h - 1;
    for (int x = 0; x < ow; x++) {
      int sx = (int)floor(x * ifx);
      if (sx > w - 1) sx = w - 1;
      out[(size_t)y * ow + x] = in[(size_t)sy * w + sx];
    }
  }
}

/* rleToString (maskApi.c): returns the number of characters written (no NUL) */
static long long mno_rle_to_string(const unsigned* cnts, long long m, unsigned char* s, long long cap, long long p0) {
  long long p = p0;
  for (long long i = 0; i < m; i++) {
    long long x = (long long)cnts[i];
    if (i > 2) x -= (long long)cnts[i - 2];
    int more = 1;
    while (more) {
      int c = (int)(x & 0x1f);
      x >>= 5;
      more = (c & 0x10) ? x != -1 : x != 0;
      if (more) c |= 0x20;
      c += 48;
      if (p < cap) s[p] = (unsigned char)c;
      p++;
    }
  }
  return p - p0;
}

/* mask: int32 [h][w] labels 0..n.  offsets: n + 1 entries.  Returns the total number of bytes needed. */
long long mno_coco_rle(const int* mask, int h, int w, int n, unsigned char* counts, long long cap, long long* offsets) {
  const long long a = (long long)h * w;
  unsigned* cnts = (unsigned*)malloc(sizeof(unsigned) * (size_t)(a + 1));
  long long p = 0;
  for (int v = 1; v <= n; v++) {
    /* rleEncode over the column-major (Fortran) binary mask (mask == v) */
    long long k = 0;
    unsigned c = 0;
    int prev = 0;
    for (long long j = 0; j < a; j++) {
      const int col = (int)(j / h), row = (int)(j % h);
      const int t = mask[(size_t)row * w + col] == v;
      if (t != prev) { cnts[k++] = c; c = 0; prev = t; }
      c++;
    }
    cnts[k++] = c;
    offsets[v - 1] = p;
    p += mno_rle_to_string(cnts, k, counts, cap, p);
  }
  offsets[n] = p;
  free(cnts);
  return p;
}

/* rleFrString + rleDecode: paints instance v's pixels (value v) into a zeroed int32 [h][w] mask */
int mno_coco_rle_decode(const unsigned char* s, long long len, int h, int w, int v, int* mask) {
  const long long a = (long long)h * w;
  unsigned* cnts = (unsigned*)malloc(sizeof(unsigned) * (size_t)(len + 1));
  long long m = 0, p = 0;
  while (p < len) {
    long long x = 0;
    int k = 0, more = 1;
    while (more) {
      if (p >= len) { free(cnts); return -1; }
      const int c = (int)s[p] - 48;
      x |= (long long)(c & 0x1f) << (5 * k);
      more = c & 0x20;
      p++; k++;
      if (!more && (c & 0x10)) x |= -1ll << (5 * k);
    }
    if (m > 2) x += (long long)cnts[m - 2];
    cnts[m++] = (unsigned)x;
  }
  long long j = 0;
  int val = 0;
  for (long long i = 0; i < m; i++) {
    for (unsigned q = 0; q < cnts[i]; q++, j++) {
      if (j >= a) { free(cnts); return -2; }
      if (val) mask[(size_t)(j % h) * w + (size_t)(j / h)] = v;
    }
    val = !val;
  }
  free(cnts);
  return j == a ? 0 : -3;
}
