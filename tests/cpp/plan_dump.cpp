// Prints the pass plans of the MODWT and DWT planners as JSON lines (host-only; used by tests/test_plan_bounds.py).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../jwave-pro_b200/csrc/jwc_dwt_plan.cuh"
#include "../../jwave-pro_b200/csrc/jwc_modwt_plan.cuh"

using namespace jwc;

int main(int argc, char** argv) {
  if (argc < 7) return 2;
  const char* kind = argv[1];
  const int64_t n = atoll(argv[2]);
  const int levels = atoi(argv[3]), L = atoi(argv[4]), inverse = atoi(argv[5]), budget = atoi(argv[6]);
  if (!strcmp(kind, "fastdiv")) {   // argv[2] = divisor: the multiply-high quotient against the real one
    const unsigned d = (unsigned)n;
    const FastDiv f = make_fastdiv(d);
    unsigned long long bad = 0, checked = 0;
    auto chk = [&](unsigned long long v) {
      if (v >= (1ull << 31)) return;
      checked++;
      if (fastdiv((unsigned)v, f) != (unsigned)(v / d)) bad++;
    };
    for (unsigned long long v = 0; v < 300000; v++) chk(v);
    for (unsigned long long q = 1; q < 200000; q++) { chk(q * d); chk(q * d - 1); chk(q * d + 1); }
    for (unsigned long long v = (1ull << 31) - 300000; v < (1ull << 31); v++) chk(v);
    unsigned long long x = 88172645463325252ull;
    for (int i = 0; i < 2000000; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; chk(x & 0x7fffffffull); }
    printf("{\"ok\": %d, \"checked\": %llu, \"m\": %u, \"l\": %d}\n", bad == 0 ? 1 : 0, checked, f.m, f.l);
    return 0;
  }
  if (!strcmp(kind, "modwt")) {
    ModwtPlanInput in{};
    in.n = n; in.J = levels; in.L = L; in.aligned16 = true; in.smem_budget = budget; in.inverse = inverse != 0;
    ModwtPlan p = modwt_plan(in);
    printf("{\"all_fused\": %d, \"generic_from\": %d, \"R\": %d, \"passes\": [", (int)p.all_fused, p.generic_from, kModwtR);
    for (size_t i = 0; i < p.passes.size(); i++) {
      const ModwtPass& q = p.passes[i];
      printf("%s{\"j0\": %d, \"k\": %d, \"logP\": %d, \"T2\": %d, \"Hp\": %d, \"mode\": %d, \"vcap\": %d, \"threads\": %d, \"smem\": %zu, \"cycles\": %lld}",
             i ? ", " : "", q.j0, q.k, q.logP, q.T2, q.Hp, q.mode, q.vcap, q.threads, q.smem, (long long)modwt_cycles(n, q.j0));
    }
    printf("]}\n");
  } else {
    DwtPlanInput in{};
    in.n = n; in.levels = levels; in.L = L; in.tree = !strcmp(kind, "wpt"); in.inverse = inverse != 0; in.aligned16 = true;
    in.smem_budget = budget;
    if (argc > 7) in.group_override = atoi(argv[7]);   // optional: maximum levels per pass
    DwtPlan p = dwt_plan(in, levels);
    printf("{\"ok\": %d, \"R\": %d, \"passes\": [", (int)p.ok, kDwtR);
    for (size_t i = 0; i < p.passes.size(); i++) {
      const DwtPass& q = p.passes[i];
      printf("%s{\"l0\": %d, \"k\": %d, \"T\": %d, \"cap\": %d, \"mode\": %d, \"threads\": %d, \"smem\": %zu}", i ? ", " : "",
             q.l0, q.k, q.T, q.cap, q.mode, q.threads, q.smem);
    }
    printf("]}\n");
  }
  return 0;
}
