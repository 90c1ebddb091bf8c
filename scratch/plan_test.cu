#include <cstdio>
#include "../votca_b200/csrc/gemm_dmma.cuh"
using namespace gwbse;
int main() {
  struct S { int M, N; long long Ko, Ki; int Z; bool lower, w, nsc; const char* name; };
  S shapes[] = {{544784, 3177, 1, 3177, 1, false, false, false, "mul_right"},
                {3177, 3177, 144, 1105, 1, true, true, false, "eps syrk"},
                {287, 270, 3177, 144, 1, false, false, false, "hd2 step2 old"},
                {287, 2160, 3177, 144, 1, false, false, false, "hd2 step2 big"},
                {144, 2160, 3177, 144, 1, false, false, false, "hd step2 big"},
                {144, 270, 3177, 144, 1, false, false, false, "hd step2 old"},
                {4320, 28593, 1, 287, 1, false, false, true, "hd step1"},
                {4320, 228744, 1, 287, 1, false, false, true, "hd step1 big"},
                {1249, 431, 1, 1249, 64, false, false, false, "fill"},
                {431, 431, 144, 3177, 1, true, false, false, "sigma_x"},
                {41328, 30, 1, 3177, 1, false, false, false, "hx2"},
                {3177, 30, 144, 287, 1, false, false, false, "hx1"},
                {3177, 3177, 1, 3177, 1, false, false, false, "dense"},
                {13, 1, 1, 82656, 1, false, false, false, "tiny"}};
  double w = 1.0;
  for (auto& s : shapes) {
    GemmParams p;
    p.M = s.M; p.N = s.N; p.Ko = (int)s.Ko; p.Ki = (int)s.Ki; p.Z1 = s.Z;
    p.lower_only = s.lower; p.w = s.w ? &w : nullptr; p.nscale = s.nsc ? &w : nullptr;
    int cfg, sw, sk;
    gemm_plan_describe(p, 148, -1, 0, &cfg, &sw, &sk);
    printf("%-16s M=%d N=%d K=%lld Z=%d -> cfg%d swap=%d splitk=%d\n", s.name, s.M, s.N, s.Ko * s.Ki, s.Z, cfg, sw, sk);
  }
}
