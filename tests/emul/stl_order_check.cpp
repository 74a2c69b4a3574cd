// tests/emul/stl_order_check.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Pins mergenet_b200/csrc/mn_stl_order.h (the restated libstdc++ container orders the tie-exact replay relies on)
// against the REAL containers of this toolchain: random sequences of the operations the reference performs
// (operator[] of a new key, erase, find, full iteration; push / top+pop with many equal priorities) must give the
// same iteration order and the same pop order.  Prints "ok <checks>" and exits 0, or says where it diverged.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <queue>
#include <unordered_map>
#include <vector>

#include "../../mergenet_b200/csrc/mn_stl_order.h"
#include "../../mergenet_b200/csrc/mn_stl_primes.h"

static const unsigned kPrimes[MN_STL_NPRIMES] = {MN_STL_PRIMES};

struct Nodes {
  int* nx;
  unsigned long long* k;
  unsigned long long key(int n) const { return k[n]; }
  int next(int n) const { return nx[n]; }
  void set_next(int n, int v) const { nx[n] = v; }
};

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

struct Cmp {
  bool operator()(const std::pair<float, int>& a, const std::pair<float, int>& b) const { return a.first < b.first; }
};

int main() {
  long long checks = 0, total_collections = 0;
  // ---- hash tables ----------------------------------------------------------------------------------------
  for (int trial = 0; trial < 60; trial++) {
    const int T = 1 + (int)(rnd() % 6);          // tables sharing one arena
    const int ops = 2000 + (int)(rnd() % 60000);
    const int max_nodes = 200000;
    std::vector<int> nx(max_nodes, -1);
    std::vector<unsigned long long> keys(max_nodes, 0);
    std::vector<MnStlTab> tab(T);
    // a small half-space on odd trials, so that the semi-space collection runs many times
    const long long half = trial % 2 ? (long long)(2.4 * ops) + 200 * T : 4000000;
    std::vector<int> arena((size_t)(2 * half));
    long long bump = 0, base = 0, collections = 0;
    int overflow = 0;
    MnStlArena A{arena.data(), half, &bump, &base, tab.data(), T, kPrimes, &overflow, &collections};
    Nodes np{nx.data(), keys.data()};
    std::vector<std::unordered_map<size_t, int>> real(T);
    std::vector<std::vector<unsigned long long>> live(T);
    for (auto& t : tab) mns_tab_init(t);
    int next_node = 0;
    const int key_mode = trial % 3;               // 0: record hashes 1619 a + 3203 b, 1: consecutive ids, 2: random
    const int erase_pct = 10 + (int)(rnd() % 45);
    for (int op = 0; op < ops && next_node < max_nodes; op++) {
      const int ti = (int)(rnd() % T);
      const bool do_erase = !live[ti].empty() && (int)(rnd() % 100) < erase_pct;
      if (do_erase) {
        const size_t j = rnd() % live[ti].size();
        const unsigned long long k = live[ti][j];
        live[ti][j] = live[ti].back();
        live[ti].pop_back();
        const int node = mns_erase(tab[ti], A, np, k);
        if (node < 0 || real[ti].at(k) != node) { printf("erase mismatch trial %d op %d\n", trial, op); return 1; }
        real[ti].erase(k);
      } else {
        unsigned long long k;
        if (key_mode == 0) { const unsigned long long a = rnd() % 3000, b = a + 1 + rnd() % 3000; k = a * 1619 + b * 3203; }
        else if (key_mode == 1) k = (unsigned long long)live[ti].size() + (unsigned long long)op;
        else k = rnd() >> (rnd() % 40);
        const bool present = real[ti].count(k) != 0;
        const int f = mns_find(tab[ti], A, np, k);
        if (present != (f >= 0) || (present && real[ti][k] != f)) { printf("find mismatch trial %d op %d\n", trial, op); return 1; }
        if (present) continue;
        const int node = next_node++;
        keys[node] = k;
        mns_insert(tab[ti], A, np, k, node);
        real[ti][k] = node;
        live[ti].push_back(k);
      }
      if (overflow) { printf("arena overflow\n"); return 1; }
      total_collections += collections; collections = 0;
      if (op % 97 == 0 || op == ops - 1) {  // full iteration order, bucket count, size
        for (int t = 0; t < T; t++) {
          if (real[t].bucket_count() != tab[t].nbkt || real[t].size() != tab[t].cnt) {
            printf("bucket count / size mismatch trial %d op %d: %zu/%u %zu/%u\n", trial, op, real[t].bucket_count(), tab[t].nbkt, real[t].size(), tab[t].cnt);
            return 1;
          }
          int n = tab[t].first;
          for (auto it = real[t].begin(); it != real[t].end(); ++it, n = nx[n]) {
            if (n < 0 || it->second != n) { printf("iteration order mismatch trial %d op %d table %d\n", trial, op, t); return 1; }
            checks++;
          }
          if (n != MNS_NULL) { printf("list longer than the real table, trial %d\n", trial); return 1; }
        }
      }
    }
  }
  // ---- thousands of small tables growing and dying in a tight arena (the replay's situation): many collections ------
  for (int trial = 0; trial < 6; trial++) {
    const int T = 3000;
    const int max_nodes = 900000;
    std::vector<int> nx(max_nodes, -1);
    std::vector<unsigned long long> keys(max_nodes, 0);
    std::vector<MnStlTab> tab(T);
    const long long half = 32ll * T;
    std::vector<int> arena((size_t)(2 * half));
    long long bump = 0, base = 0, collections = 0;
    int overflow = 0;
    MnStlArena A{arena.data(), half, &bump, &base, tab.data(), T, kPrimes, &overflow, &collections};
    Nodes np{nx.data(), keys.data()};
    std::vector<std::unordered_map<size_t, int>> real(T);
    std::vector<char> dead(T, 0);
    for (auto& t : tab) mns_tab_init(t);
    int next_node = 0;
    for (int op = 0; op < 1000000 && next_node < max_nodes; op++) {
      const int ti = (int)(rnd() % T);
      const unsigned r = (unsigned)(rnd() % 1000);
      if (r < 20) {  // the owner dies (its array becomes garbage); the slot starts over as a fresh table
        mns_tab_drop(tab[ti]); mns_tab_init(tab[ti]); real[ti] = std::unordered_map<size_t, int>();
      } else if (r < 300 && !real[ti].empty()) {
        const unsigned long long k = real[ti].begin()->first;
        const int node = mns_erase(tab[ti], A, np, k);
        if (node != real[ti].begin()->second) { printf("erase mismatch (many tables)\n"); return 1; }
        real[ti].erase(k);
      } else if (real[ti].size() < 24) {
        const unsigned long long a = rnd() % 5000, b = a + 1 + rnd() % 5000, k = a * 1619 + b * 3203;
        if (real[ti].count(k)) continue;
        const int node = next_node++;
        keys[node] = k;
        mns_insert(tab[ti], A, np, k, node);
        real[ti][k] = node;
      }
      if (overflow) { printf("arena overflow (many tables)\n"); return 1; }
    }
    for (int t = 0; t < T; t++) {
      if (dead[t]) continue;
      if (real[t].bucket_count() != tab[t].nbkt || real[t].size() != tab[t].cnt) { printf("bucket count mismatch (many tables)\n"); return 1; }
      int n = tab[t].first;
      for (auto it = real[t].begin(); it != real[t].end(); ++it, n = nx[n]) {
        if (n < 0 || it->second != n) { printf("iteration order mismatch (many tables) %d\n", t); return 1; }
        checks++;
      }
      if (n != MNS_NULL) { printf("list too long (many tables)\n"); return 1; }
    }
    total_collections += collections;
  }
  // ---- priority queue with few distinct priorities ----------------------------------------------------------
  for (int trial = 0; trial < 40; trial++) {
    std::priority_queue<std::pair<float, int>, std::vector<std::pair<float, int>>, Cmp> real;
    const long long cap = 400000;
    std::vector<float> qk(cap);
    std::vector<int> qr(cap);
    MnStlHeap h{qk.data(), qr.data(), 0, cap};
    const int levels = 1 + (int)(rnd() % 7);
    const int ops = 1000 + (int)(rnd() % 150000);
    int id = 0;
    for (int op = 0; op < ops; op++) {
      const bool pop = !real.empty() && (rnd() % 100) < (unsigned)(trial % 2 ? 55 : 40);
      if (pop) {
        float k; int r;
        const auto top = real.top();
        real.pop();
        mns_heap_pop(h, &k, &r);
        if (k != top.first || r != top.second) { printf("pop order mismatch trial %d op %d\n", trial, op); return 1; }
        checks++;
      } else {
        const float k = (float)(rnd() % levels) * 0.25f;
        real.push(std::make_pair(k, id));
        if (!mns_heap_push(h, k, id)) { printf("heap full\n"); return 1; }
        id++;
      }
    }
    while (!real.empty()) {
      float k; int r;
      const auto top = real.top();
      real.pop();
      mns_heap_pop(h, &k, &r);
      if (k != top.first || r != top.second) { printf("drain order mismatch trial %d\n", trial); return 1; }
      checks++;
    }
    if (h.n != 0) { printf("heap not empty\n"); return 1; }
  }
  if (total_collections == 0) { printf("the collection never ran\n"); return 1; }
  printf("ok %lld checks, %lld collections\n", checks, total_collections);
  return 0;
}
