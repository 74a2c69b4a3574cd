// tests/emul/emul_exact.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the tie-exact replay (mergenet_b200/csrc/mn_exact.cuh + mn_stl_order.h) for the HOST so that the CPU
// suite can compare its logic with the unmodified reference (oracle/_ref/libsegment_ref.so) on tie-dependent inputs
// without a GPU.  The edge quantities the device path takes from its edge pass are computed here with the host's
// libm exactly as the reference's constructor writes them (segment.cc:5-21,24-46,183-195).  Same argument meaning
// as the drop-in symbol c_run_segmentation; nothing in mergenet_b200/ can reach this file.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../mergenet_b200/csrc/mn_exact.cuh"
#include "../../mergenet_b200/csrc/mn_stl_primes.h"

static const unsigned kPrimes[MN_STL_NPRIMES] = {MN_STL_PRIMES};
template <typename T> static T* zalloc(size_t n) { return (T*)calloc(n ? n : 1, sizeof(T)); }

extern "C" int emul_exact_segment(float* class_pred, int C, float* adj_pred, int K, int W, int H, const int* offset_list,
                                  float sdb, float omf, float mlb, int* mask, int* object_class, int* n_instances,
                                  long long* stats8, long long arena_half_words) {
  const size_t N = (size_t)H * W, E = N * K;
  MnExact m;
  memset(&m, 0, sizeof(m));
  m.C = C; m.K = K; m.H = H; m.W = W; m.N = (int)N; m.E = (long long)E; m.omf = omf; m.mlb = mlb;
  for (int k = 0; k < K; k++) { m.off_r[k] = offset_list[2 * k]; m.off_c[k] = offset_list[2 * k + 1]; }
  if (sdb != 0.0f) {  // cc:183-195, in place on the caller's buffer
    for (size_t i = 0; i < E; i++) {
      const float s = adj_pred[i];
      const float logit = (float)((double)logf(s) - log(1.0 - (double)s) + (double)sdb);
      adj_pred[i] = (float)(1.0 / (1.0 + (double)expf(-logit)));
    }
  }
  float* clp = zalloc<float>(N * C);
  float* same = zalloc<float>(E);
  float* diff = zalloc<float>(E);
  for (size_t p = 0; p < N; p++)
    for (int c = 0; c < C; c++) clp[p * C + c] = 0.0f + logf(class_pred[(size_t)c * N + p]);
  for (size_t p = 0; p < N; p++)
    for (int k = 0; k < K; k++) {
      const float s = adj_pred[(size_t)k * N + p];
      same[p * K + k] = logf(s);
      diff[p * K + k] = (float)log(1.0 - (double)s);
    }
  m.clp = clp; m.rec_same = same; m.rec_diff = diff;
  m.npix = zalloc<int>(N); m.cls = zalloc<int>(N); m.pix_next = zalloc<int>(N); m.pix_tail = zalloc<int>(N);
  m.tab = zalloc<MnStlTab>(N + 1); m.ob_next = zalloc<int>(N);
  m.r_o1 = zalloc<int>(E); m.r_o2 = zalloc<int>(E); m.r_oml = zalloc<float>(E); m.r_mp = zalloc<float>(E);
  m.r_merged = zalloc<int>(E); m.nd_next = zalloc<int>(2 * E); m.nd_key = zalloc<unsigned long long>(2 * E);
  m.heap.cap = (long long)(8 * E + 1024); m.heap.n = 0;
  m.heap.key = zalloc<float>((size_t)m.heap.cap); m.heap.rec = zalloc<int>((size_t)m.heap.cap);
  long long bump = 0, base = 0;
  int overflow = 0, status = 0, n = 0;
  long long stats[8] = {0};
  m.arena.half = arena_half_words > 0 ? arena_half_words : (long long)(32 * N + 5 * E + 4096);
  m.arena.bk = zalloc<int>((size_t)(2 * m.arena.half)); m.arena.bump = &bump; m.arena.base = &base;
  m.arena.tabs = m.tab; m.arena.ntabs = (int)N + 1; m.arena.primes = kPrimes; m.arena.overflow = &overflow;
  m.arena.collections = &stats[3];
  m.out_mask = mask; m.out_cls = object_class; m.out_n = &n; m.status = &status; m.stats = stats;
  mnx_run(m);
  *n_instances = n;
  stats[4] = bump - base; stats[5] = m.arena.half;
  if (stats8) for (int i = 0; i < 8; i++) stats8[i] = stats[i];
  free(clp); free(same); free(diff); free(m.npix); free(m.cls); free(m.pix_next); free(m.pix_tail); free(m.tab);
  free(m.ob_next); free(m.r_o1); free(m.r_o2); free(m.r_oml); free(m.r_mp); free(m.r_merged); free(m.nd_next);
  free(m.nd_key); free(m.heap.key); free(m.heap.rec); free(m.arena.bk);
  return status;
}
