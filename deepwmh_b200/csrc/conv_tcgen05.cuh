// tcgen05 / TMEM / TMA implicit-GEMM conv3d (placeholder until the kernel lands: nothing enabled).
#pragma once
#include <string>
#include <vector>
#include "common.cuh"

namespace dwmh {

struct TcLayer {
  bool enabled = false;
};

inline void tc_free(TcLayer& t) { t.enabled = false; }

inline int tc_prepare(TcLayer& t, const std::vector<float>& w, int c0, int c1, int cout, const int k[3], const int s[3],
                      const int in_sp[3], const int out_sp[3], int maxN, bool bf16, const void* in0, const void* in1,
                      void* out, std::string* why) {
  (void)t; (void)w; (void)c0; (void)c1; (void)cout; (void)k; (void)s; (void)in_sp; (void)out_sp; (void)maxN; (void)bf16;
  (void)in0; (void)in1; (void)out; (void)why;
  return 0;
}

inline int tc_init_attributes(bool bf16) { (void)bf16; return 0; }

template <typename T>
int tc_launch(TcLayer& t, int nb, double* sums, int num_sms, cudaStream_t st, std::string* err) {
  (void)t; (void)nb; (void)sums; (void)num_sms; (void)st;
  if (err) *err = "tcgen05 path not built";
  return 1;
}

}  // namespace dwmh
