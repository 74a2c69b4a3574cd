// tests/emul/emul_modeb.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the Mode-B device source (mergenet_b200/csrc/mn_modeb.cuh) for the HOST so that the CPU suite can
// check its logic against the imported reference `utils/segmenter.py::ObjectSegmenter` (fixtures under
// tests/golden/modeb, made by tests/golden/make_golden_modeb.py) without a GPU.  Same argument meaning as the
// library's mn_modeb_segment_host; nothing in mergenet_b200/ can reach this file.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../mergenet_b200/csrc/mn_modeb.cuh"

template <typename T> static T* zalloc(size_t n) { return (T*)calloc(n ? n : 1, sizeof(T)); }

extern "C" int emul_modeb_segment(const float* logc, const float* lsame, const float* ldiff, int C, int K, int H, int W,
                                  const int* offset_list, double omf, double mlb, double prune_threshold,
                                  long long* mask, int* object_class, int* n_instances, long long* stats4) {
  const size_t N = (size_t)H * W, E = N * K;
  MnModeB m;
  memset(&m, 0, sizeof(m));
  m.C = C; m.K = K; m.H = H; m.W = W; m.N = (int)N; m.E = (long long)E;
  for (int k = 0; k < K; k++) { m.off_r[k] = offset_list[2 * k]; m.off_c[k] = offset_list[2 * k + 1]; }
  m.omf = omf; m.mlb = mlb; m.omf32 = (float)omf; m.mlb32 = (float)mlb; m.prune_threshold = prune_threshold;
  unsigned hm = 1;
  while ((size_t)hm < 2 * E + 16) hm <<= 1;
  m.h_mask = hm - 1;
  m.q_cap = (long long)(8 * E + 1024);
  m.logc = logc; m.lsame = lsame; m.ldiff = ldiff;
  m.npix = zalloc<int>(N); m.cls = zalloc<int>(N); m.clp = zalloc<double>(N * C); m.osame = zalloc<float>(N);
  m.alive = zalloc<unsigned char>(N); m.adj_head = zalloc<int>(N); m.adj_tail = zalloc<int>(N);
  m.pix_next = zalloc<int>(N); m.pix_tail = zalloc<int>(N);
  m.r_o1 = zalloc<int>(E); m.r_o2 = zalloc<int>(E); m.r_oml = zalloc<float>(E); m.r_same = zalloc<float>(E);
  m.r_diff = zalloc<float>(E); m.r_mp = zalloc<double>(E); m.r_link = zalloc<int>(E * 6);
  m.h_key = zalloc<unsigned long long>(hm); m.h_val = zalloc<int>(hm);
  memset(m.h_key, 0xFF, (size_t)hm * 8);
  m.q_key = zalloc<double>((size_t)m.q_cap); m.q_rec = zalloc<int>((size_t)m.q_cap); m.q_n = 0;
  int status = 0, n = 0;
  long long stats[8] = {0};
  m.out_mask = mask; m.out_cls = object_class; m.out_n = &n; m.status = &status; m.stats = stats;
  for (size_t i = 0; i < N; i++) object_class[i] = -1;
  mnb_run(m);
  *n_instances = n;
  if (stats4) for (int i = 0; i < 4; i++) stats4[i] = stats[i];
  free(m.npix); free(m.cls); free(m.clp); free(m.osame); free(m.alive); free(m.adj_head); free(m.adj_tail);
  free(m.pix_next); free(m.pix_tail); free(m.r_o1); free(m.r_o2); free(m.r_oml); free(m.r_same); free(m.r_diff);
  free(m.r_mp); free(m.r_link); free(m.h_key); free(m.h_val); free(m.q_key); free(m.q_rec);
  return status;
}
