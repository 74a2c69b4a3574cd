// dev check (host): the 64-bin log(1 - s) evaluation of mn_edge_warp_kernel, emulated with fma(), against
// the host libm on the whole clipped domain.  built and run by tests/test_libm_parity.py::test_log1m_64bin_table_exhaustive (gcc -O2 [-mfma] ... -lm)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../mergenet_b200/csrc/mn_log1m_tab.h"
typedef struct { double invc, logc; } T;
static const T tab[64] = {MN_LOG1M64_TABLE};
int main(void) {
  const double Ln2 = 0x1.62e42fefa39efp-1;
  long long bad = 0, amb = 0, n = 0;
  for (uint32_t b = 0x34000000u; b <= 0x3f7ffffeu; b++, n++) {
    float s; memcpy(&s, &b, 4);
    double x = 1.0 - (double)s;
    uint64_t xb; memcpy(&xb, &x, 8);
    uint32_t hx = (uint32_t)(xb >> 32);
    uint32_t tmp = hx - (uint32_t)(MN_LOG1M64_OFF >> 32);
    int i = (tmp >> 14) & 63;
    int k = (int32_t)tmp >> 20;
    uint64_t zb = ((uint64_t)(hx - (tmp & 0xfff00000u)) << 32) | (uint32_t)xb;
    double z; memcpy(&z, &zb, 8);
    double r = fma(z, tab[i].invc, -1.0);
    double t = fma((double)k, Ln2, tab[i].logc);
    double q = fma(r, -1.0 / 6, 0.2);
    q = fma(r, q, -0.25); q = fma(r, q, 1.0 / 3); q = fma(r, q, -0.5);
    double r2 = r * r;
    double y = fma(r2, q, r) + t;
    uint64_t yb; memcpy(&yb, &y, 8);
    uint32_t c = ((uint32_t)yb << 3) + ((0x4000u - 0x10000000u) << 3);
    if (c < (0x8000u << 3)) { amb++; continue; }
    float got = (float)y, want = (float)log(x);
    if (memcmp(&got, &want, 4)) { if (bad < 5) printf("bad %08x got %a want %a\n", b, got, want); bad++; }
  }
  printf("inputs %lld ambiguous %lld (%.2e) mismatches outside the fallback set %lld\n", n, amb, (double)amb / n, bad);
  return bad != 0;
}
