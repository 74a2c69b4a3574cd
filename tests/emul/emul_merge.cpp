// tests/emul/emul_merge.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the device scheduler source (mergenet_b200/csrc/mn_merge.cuh) for the HOST and runs it
// with one logical thread, so the CPU test-suite can unit-test the scheduler's logic (queue tree,
// plan/commit rounds, solo merges) against the oracle without a GPU.  The record construction
// below is a plain host loop using libm (it is not the CUDA edge pass); only the scheduler is the
// product's source.  Nothing in mergenet_b200/ can reach this file.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../mergenet_b200/csrc/mn_merge.cuh"

template <typename T> static T* zalloc(size_t n) { return (T*)calloc(n ? n : 1, sizeof(T)); }

extern "C" int emul_run_segmentation(const float* class_pred, int class_dim, const float* adj_pred,
                                     int offset_dim, int W, int H, int num_classes,
                                     const int* offsets, int* output, int* object_class, float sdb,
                                     float omf, float mlb, long long* stats /* 16 */) {
  (void)class_dim; (void)sdb;
  const int C = num_classes, K = offset_dim, N = H * W;
  const size_t E = (size_t)N * K;
  MnImage im;
  memset(&im, 0, sizeof(im));
  im.clp = zalloc<float>((size_t)N * C);
  im.cls = zalloc<int>(N);
  im.obj = zalloc<uint4>(N);
  im.parent = zalloc<int>(N);
  // pool capacities: the library's own (mn_layout.h: mn_workspace_caps), so that a pool that is too small fails here.
  // Two deliberate exceptions: EMUL_HASH_PERMILLE (slots per 1000 records; a tight table exercises the overflow area),
  // and the stress builds with small tree leaves (-DMN_LEAFCAP < 512), which split far more often than the product.
  const MnCaps caps = mn_workspace_caps((size_t)N, E);
  im.pix_cap = caps.pix_cap;
  im.pix_pool = zalloc<int>(im.pix_cap);
  im.rec = zalloc<uint4>(E);
  im.hash_nbuckets = caps.hash_nbuckets;
  if (getenv("EMUL_HASH_PERMILLE")) im.hash_nbuckets = (uint32_t)(E * atoll(getenv("EMUL_HASH_PERMILLE")) / 1000 / 8 + 64);
  im.hash = zalloc<uint32_t>((size_t)im.hash_nbuckets * 8);
  im.hash_ovf_cap = caps.hash_ovf_cap;
  im.hash_ovf = zalloc<uint32_t>(im.hash_ovf_cap);
  im.qc_low_n = caps.qc_low_n;
  im.qc_cap = caps.qc_cap;
  im.q_ent = zalloc<uint4>((size_t)im.qc_cap * MN_QCH + 2 * MN_QCH);
  im.init_keys = (uint64_t*)im.q_ent;  // one arena, as in the library
  im.qc_next = zalloc<int>(im.qc_cap);
  im.qc_free = zalloc<int>(im.qc_cap);
  im.tn_cap = caps.tn_cap;
#if MN_LEAFCAP < 512
  im.tn_cap = MN_NROOTS + MN_TREE_FANOUT * (int)(16384 + E / 128);
#endif
  im.tn = zalloc<int4>(im.tn_cap);
  im.tn_dir = zalloc<int>((size_t)im.tn_cap * 8);
  im.ctl = zalloc<MnCtl>(1);
  for (int i = 0; i < im.tn_cap; i++) im.tn[i] = make_int4(-1, -1, 0, -1);
  im.ctl->tn_bump = MN_NROOTS;

  MnMergeArgs A;
  memset(&A, 0, sizeof(A));
  A.C = C; A.K = K; A.N = N; A.W = W; A.omf = omf; A.mlb = mlb; A.max_rounds = 0; A.H = MN_H;
  A.off.K = K;
  std::vector<std::pair<int, int>> mag;
  for (int k = 0; k < K; k++) {
    A.off.delta[k] = offsets[2 * k] * W + offsets[2 * k + 1];
    mag.push_back(std::make_pair(abs(A.off.delta[k]), k));
  }
  std::sort(mag.begin(), mag.end());
  int rank_of_k[MN_MAX_K];
  for (int r = 0; r < K; r++) { A.off.k_of_rank[r] = mag[r].second; rank_of_k[mag[r].second] = r; }

  // objects (cc:196-207) and records (cc:209-231): host loops, libm
  for (int p = 0; p < N; p++) {
    float best = 0; int bc = 0;
    for (int c = 0; c < C; c++) {
      float v = 0.0f + logf(class_pred[(size_t)c * N + p]);
      im.clp[(size_t)p * C + c] = v;
      if (c == 0 || v > best) { best = v; bc = c; }
    }
    im.cls[p] = bc;
    im.obj[p] = make_uint4(mn_pack_nc(1, bc), 0u, 0xffffffffu, 0u);
    im.parent[p] = p;
  }
  for (int p = 0; p < N; p++) {
    int row = p / W, col = p % W;
    uint32_t m = 0;
    for (int k = 0; k < K; k++) {
      int r2 = row + offsets[2 * k], c2 = col + offsets[2 * k + 1];
      if (r2 >= 0 && r2 < H && c2 >= 0 && c2 < W) m |= 1u << k;
      r2 = row - offsets[2 * k]; c2 = col - offsets[2 * k + 1];
      if (r2 >= 0 && r2 < H && c2 >= 0 && c2 < W) m |= 1u << (16 + k);
    }
    im.obj[p].w = m;
    for (int k = 0; k < K; k++) {
      size_t r = (size_t)p * K + k;
      int r2 = row + offsets[2 * k], c2 = col + offsets[2 * k + 1];
      uint64_t key = ~0ull;
      if (r2 >= 0 && r2 < H && c2 >= 0 && c2 < W) {
        int q = r2 * W + c2, lo = p < q ? p : q, hi = p < q ? q : p;
        float s = adj_pred[(size_t)k * N + p];
        float diff = (float)log(1.0 - (double)s), same = logf(s), oml = same - diff;
        float mp = mn_priority(oml, omf, mlb, C, 1, im.cls[lo], im.clp + (size_t)lo * C, 1, im.cls[hi],
                               im.clp + (size_t)hi * C, nullptr);
        const uint32_t g = mp >= 0.0f ? MN_G_EXACT : MN_G_NONE;
        mn_store_rec(im, (int)r, make_uint4(mn_rec_pack_x(lo, MN_HS_NONE, g), (uint32_t)hi, mn_f2u(oml), mn_f2u(mp)));
        int hslot = mn_hash_insert(im, lo, hi, (int)r);
        MN_REC(im, r).x = mn_rec_pack_x(lo, mn_hs_of_slot(mn_hash_pos(im.hash_nbuckets, lo, hi), hslot), g);
        if (mp >= 0.0f) {
          uint32_t ord = (mn_tie_u(lo, hi) << 4) | (uint32_t)rank_of_k[k];
          key = ((uint64_t)(~mn_f2u(mp == 0.0f ? 0.0f : mp)) << MN_ORD_BITS) | ord;
        }
      } else {
        mn_store_rec(im, (int)r, make_uint4(MN_REC_DEAD, 0u, 0u, mn_f2u(-1.0f)));
      }
      im.init_keys[r] = key;
    }
  }
  std::sort(im.init_keys, im.init_keys + E);

  MnSm* sm = (MnSm*)calloc(1, sizeof(MnSm));
  float* c_clp = zalloc<float>((size_t)MN_H * 3 * C);
  mn_merge_image(im, *sm, A, c_clp);

  // labels (cc:491-517): ascending surviving id, class-0 objects -> 0
  std::vector<int> label(N, 0);
  for (int i = 0; i < N; i++) { output[i] = 0; object_class[i] = -1; }
  int k = 1;
  for (int o = 0; o < N; o++) {
    if (im.parent[o] != o) continue;
    int cls = mn_nc_cls(im.obj[o].x);
    if (cls == 0) continue;
    object_class[k - 1] = cls;
    label[o] = k++;
  }
  for (int p = 0; p < N; p++) {
    int r = p;
    while (im.parent[r] != r) r = im.parent[r];
    output[p] = label[r];
  }
  int status = im.ctl->status;
  if (getenv("EMUL_DEBUG")) {
    for (size_t r = 0; r < E; r++) if (mn_rec_key(im, (int)r).x >= 0 && mn_u2f(MN_REC(im, r).w) >= 0.0f)
      fprintf(stderr, "LEFTOVER rec %zu lo %d hi %d mp %.9g (bits %08x) guard %u root %d\n", r, mn_rec_key(im, (int)r).x, mn_rec_key(im, (int)r).y, mn_u2f(MN_REC(im, r).w), MN_REC(im, r).w, mn_rec_guard(MN_REC(im, r).x), mn_root_of(mn_u2f(MN_REC(im, r).w)));
    fprintf(stderr, "status %d fail_line %d hash_ovf_n %d peak_entries %d peak_chunks %d (E %zu) tn_bump %d qc_bump %d pix_bump %d\n", im.ctl->status, im.ctl->fail_line, im.ctl->hash_ovf_n, im.ctl->peak_entries, im.ctl->peak_chunks, E, im.ctl->tn_bump, im.ctl->qc_bump, im.ctl->pix_bump);
    fprintf(stderr, "tree_entries %d static_cursor %d n_init %d nins %d nhot %d cold_empty %d\n", im.ctl->tree_entries, im.ctl->static_cursor, im.ctl->n_init, sm->nins, sm->nhot, sm->cold_empty);
  }
  if (stats) {
    MnCtl* c = im.ctl;
    long long v[16] = {c->rounds, c->events, c->merges, c->restores, c->invalid_pops, c->solo_events,
                       c->refills, c->flushes, c->splits, c->pairs, c->cuts_conflict, c->cuts_cascade,
                       c->cuts_capacity, (long long)c->qc_bump, (long long)c->pix_gcs, (long long)c->tn_bump};
    memcpy(stats, v, sizeof(v));
  }
  // (the long sweep, tests/manual/soak_sweep.py, calls this tens of thousands of times per process)
  free(im.clp); free(im.cls); free(im.obj); free(im.parent); free(im.pix_pool); free(im.rec); free(im.hash);
  free(im.hash_ovf); free(im.q_ent); free(im.qc_next); free(im.qc_free); free(im.tn); free(im.tn_dir); free(im.ctl);
  free(sm); free(c_clp);
  return status;
}

// The pool capacities of the library (mn_layout.h: mn_workspace_caps) for a shape, so that the CPU suite can check their
// invariants directly: {pix_cap, qc_low_n, qc_cap, tn_cap, hash_nbuckets, hash_ovf_cap, MN_QCH}.
extern "C" void emul_workspace_caps(long long N, long long E, long long* out7) {
  const MnCaps c = mn_workspace_caps((size_t)N, (size_t)E);
  out7[0] = c.pix_cap; out7[1] = c.qc_low_n; out7[2] = c.qc_cap; out7[3] = c.tn_cap;
  out7[4] = c.hash_nbuckets; out7[5] = c.hash_ovf_cap; out7[6] = MN_QCH;
}

// The accept pass compares 64-bit pop keys instead of full (mp, tie) orders: whenever entry a pops
// before entry b, key(a) >= key(b) must hold (equal keys are treated as "pops first": the safe side).
// Returns the number of violations over n random pairs (many of them tied on mp, close in lo/hi).
extern "C" long long emul_pop_key_violations(long long n, unsigned seed) {
  unsigned long long s = seed * 6364136223846793005ull + 1442695040888963407ull;
  auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(s >> 33); };
  const float mps[6] = {0.0f, 0.03f, 0.030000001f, 1.5f, 2.4e-7f, 17.25f};
  long long bad = 0;
  for (long long i = 0; i < n; i++) {
    const int N = 1 << (8 + rnd() % 14);
    int alo = (int)(rnd() % N), ahi = alo + 1 + (int)(rnd() % 90000);
    int blo = (rnd() % 4 == 0) ? alo : (int)(rnd() % N), bhi = blo + 1 + (int)(rnd() % 90000);
    float amp = mps[rnd() % 6], bmp = (rnd() % 2) ? amp : mps[rnd() % 6];
    const bool ab = mn_before(amp, alo, ahi, bmp, blo, bhi), ba = mn_before(bmp, blo, bhi, amp, alo, ahi);
    const unsigned long long ka = mn_pop_key(amp, alo, ahi), kb = mn_pop_key(bmp, blo, bhi);
    if (ab && !(ka >= kb)) bad++;
    if (ba && !(kb >= ka)) bad++;
    if (amp > bmp && !(ka > kb)) bad++;  // a higher priority always wins strictly
  }
  return bad;
}
