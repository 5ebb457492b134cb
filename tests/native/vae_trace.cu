// Test harness (CPU): the native VAE runtime (csrc/vae_engine.cu) compiled with SVR2_HOST_TEST — host memory as the
// workspace, every kernel entry point replaced by a stub that prints its name and scalar arguments (pointers as p0 / p1
// = NULL / not NULL).  tests/test_native_vae_cpu.py compares the trace with the op sequence of the Python module
// (vae.py) on the same clip, and the dry-run workspace size with the extent the real run touches.
// usage: vae_trace <weights manifest> enc|dec T H W slice_frames
#define SVR2_HOST_TEST 1
#include "../../comfyui-seedvr2_videoupscaler_b200/csrc/vae_engine.cu"
#include <stdarg.h>
#include <stdlib.h>

#include <fstream>
#include <sstream>

namespace svr2 {
static char g_msg[512];
int set_error(int code, const char* m) { snprintf(g_msg, sizeof g_msg, "%s", m ? m : ""); return code; }
}
static char* g_lo = nullptr;
static char* g_hi = nullptr;       // workspace bounds: every non-NULL pointer into it must stay inside
static size_t g_touch = 0;
static const char* P_(const void* p) {
  if (p && (const char*)p >= g_lo && (const char*)p < g_hi) {
    const size_t off = (const char*)p - g_lo;
    if (off > g_touch) g_touch = off;
  }
  return p ? "p1" : "p0";
}
extern "C" {
const char* svr2_last_error(void) { return svr2::g_msg; }
int svr2_groupnorm_from_stats_bf16(const void* x, void* y, int frames, int hw, int C, const void* gamma, const void* beta,
                                   float eps, int silu, int out_t_pad, int out_dup_head, const void* stat_partial,
                                   int stat_slots, void* coef_scratch, void* stream) {
  printf("svr2_groupnorm_from_stats_bf16 %s %s %d %d %d %s %s %.5g %d %d %d %s %d %s %s\n", P_(x), P_(y), frames, hw, C, P_(gamma),
         P_(beta), eps, silu, out_t_pad, out_dup_head, P_(stat_partial), stat_slots, P_(coef_scratch), P_(stream));
  return 0;
}
int svr2_groupnorm_bf16(const void* x, void* y, int frames, int hw, int C, const void* gamma, const void* beta, float eps,
                        int silu, int out_t_pad, int out_dup_head, double* scratch, int64_t scratch_bytes, void* stream) {
  printf("svr2_groupnorm_bf16 %s %s %d %d %d %s %s %.5g %d %d %d %s %lld %s\n", P_(x), P_(y), frames, hw, C, P_(gamma), P_(beta), eps,
         silu, out_t_pad, out_dup_head, P_(scratch), (long long)scratch_bytes, P_(stream));
  return 0;
}
int svr2_conv3d_bf16(const void* x, int T_in_total, int H, int W, int Cin, const void* w, int Cout, int kt, int kh, int kw,
                     int stride_t, int stride_hw, int pad_hw, int T_out, int epi_flags, const void* bias, const void* residual,
                     void* y, int out_t_pad, int out_dup_head, int ldc, void* stream) {
  printf("svr2_conv3d_bf16 %s %d %d %d %d %s %d %d %d %d %d %d %d %d %d %s %s %s %d %d %d %s\n", P_(x), T_in_total, H, W, Cin, P_(w),
         Cout, kt, kh, kw, stride_t, stride_hw, pad_hw, T_out, epi_flags, P_(bias), P_(residual), P_(y), out_t_pad, out_dup_head,
         ldc, P_(stream));
  return 0;
}
int svr2_conv3d_stats_bf16(const void* x, int T_in_total, int H, int W, int Cin, const void* w, int Cout, int kt, int kh,
                           int kw, int stride_t, int stride_hw, int pad_hw, int T_out, int epi_flags, const void* bias,
                           const void* residual, void* y, int out_t_pad, int out_dup_head, int ldc, void* stat_partial,
                           int64_t stat_bytes, int* stat_slots, void* stream) {
  printf("svr2_conv3d_stats_bf16 %s %d %d %d %d %s %d %d %d %d %d %d %d %d %d %s %s %s %d %d %d %s %lld %s %s\n", P_(x), T_in_total, H,
         W, Cin, P_(w), Cout, kt, kh, kw, stride_t, stride_hw, pad_hw, T_out, epi_flags, P_(bias), P_(residual), P_(y), out_t_pad,
         out_dup_head, ldc, P_(stat_partial), (long long)stat_bytes, P_(stat_slots), P_(stream));
  *stat_slots = svr2_conv_stat_slots(Cout, stride_hw == 1 ? H : H / 2, stride_hw == 1 ? W : W / 2);
  return 0;
}
int svr2_conv3d_shortcut_stats_bf16(const void* x, int T_in_total, int H, int W, int Cin, const void* w, int Cout, int kt,
                                    int kh, int kw, int T_out, const void* bias, const void* x2, int C2, void* y,
                                    int out_t_pad, int out_dup_head, void* stat_partial, int64_t stat_bytes,
                                    int* stat_slots, void* stream) {
  printf("svr2_conv3d_shortcut_stats_bf16 %s %d %d %d %d %s %d %d %d %d %d %s %s %d %s %d %d %s %lld %s %s\n", P_(x), T_in_total, H, W,
         Cin, P_(w), Cout, kt, kh, kw, T_out, P_(bias), P_(x2), C2, P_(y), out_t_pad, out_dup_head, P_(stat_partial),
         (long long)stat_bytes, P_(stat_slots), P_(stream));
  *stat_slots = svr2_conv_stat_slots(Cout, H, W);
  return 0;
}
int svr2_upsample_shuffle_bf16(const void* x, int F, int H, int W, int C, const void* w, const void* bias, int temporal,
                               int drop_head, void* y, int out_t_pad, int out_dup_head, void* stream) {
  printf("svr2_upsample_shuffle_bf16 %s %d %d %d %d %s %s %d %d %s %d %d %s\n", P_(x), F, H, W, C, P_(w), P_(bias), temporal, drop_head,
         P_(y), out_t_pad, out_dup_head, P_(stream));
  return 0;
}
int svr2_linear_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, int M, int N, int K, int epi_flags,
                     const void* bias, const float* gate, const void* residual, void* out, int64_t ldc, float out_scale,
                     void* stream) {
  printf("svr2_linear_bf16 %s %lld %s %lld %d %d %d %d %s %s %s %s %lld %.5g %s\n", P_(a), (long long)lda, P_(w), (long long)ldw, M, N, K,
         epi_flags, P_(bias), P_(gate), P_(residual), P_(out), (long long)ldc, out_scale, P_(stream));
  return 0;
}
int svr2_linear_ex_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, int M, int N, int K, int epi_flags,
                        const void* bias, const float* gate, const void* residual, void* out, int64_t ldc, float out_scale,
                        const float* rowscale, void* stat_out, int64_t ld_stat, const int* run_if, void* stream) {
  printf("svr2_linear_ex_bf16 %s %lld %s %lld %d %d %d %d %s %s %s %s %lld %.5g %s %s %lld %s %s\n", P_(a), (long long)lda, P_(w),
         (long long)ldw, M, N, K, epi_flags, P_(bias), P_(gate), P_(residual), P_(out), (long long)ldc, out_scale, P_(rowscale),
         P_(stat_out), (long long)ld_stat, P_(run_if), P_(stream));
  return 0;
}
int svr2_rowstat_max(const void* partial, int slots, int64_t ld, float* mhat, int rows, int* flag_reset, void* stream) {
  printf("svr2_rowstat_max %s %d %lld %s %d %s %s\n", P_(partial), slots, (long long)ld, P_(mhat), rows, P_(flag_reset), P_(stream));
  return 0;
}
int svr2_pexp_stat_combine(const void* partial, int slots, int64_t ld, const float* mhat, float* rowscale, int rows, int* flag,
                           void* stream) {
  printf("svr2_pexp_stat_combine %s %d %lld %s %s %d %s %s\n", P_(partial), slots, (long long)ld, P_(mhat), P_(rowscale), rows, P_(flag),
         P_(stream));
  return 0;
}
int svr2_rowstat_combine(const void* partial, int slots, int64_t ld, float* lse, int rows, void* stream) {
  printf("svr2_rowstat_combine %s %d %lld %s %d %s\n", P_(partial), slots, (long long)ld, P_(lse), rows, P_(stream));
  return 0;
}
int svr2_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int rows, int cols, void* stream) {
  printf("svr2_transpose_bf16 %s %lld %s %lld %d %d %s\n", P_(in), (long long)ld_in, P_(out), (long long)ld_out, rows, cols, P_(stream));
  return 0;
}
int svr2_im2col3_bf16(const void* x, int T, int H, int W, int C, int ld_in, void* out, int ld_out, void* stream) {
  printf("svr2_im2col3_bf16 %s %d %d %d %d %d %s %d %s\n", P_(x), T, H, W, C, ld_in, P_(out), ld_out, P_(stream));
  return 0;
}
}
namespace svr2 {
// the channel stride is the one argument the public entry points (what vae.py calls) do not carry: printed last, after '|'
int ncdhw_to_ndhwc_strided(const void* in, int in_dtype, int C, int T, int H, int W, int64_t cs, void* out, int C_pad,
                           int out_t_pad, float div, void* stream) {
  printf("svr2_ncdhw_to_ndhwc_bf16 %s %d %d %d %d %d %s %d %d %.5g %s | %lld\n", P_(in), in_dtype, C, T, H, W, P_(out), C_pad, out_t_pad,
         div, P_(stream), (long long)cs);
  return 0;
}
int ndhwc_to_ncdhw_strided(const void* in, int ld_in, int C, int T, int H, int W, void* out, int out_dtype, int64_t cs,
                           void* stream) {
  printf("svr2_ndhwc_to_ncdhw %s %d %d %d %d %d %s %d %s | %lld\n", P_(in), ld_in, C, T, H, W, P_(out), out_dtype, P_(stream), (long long)cs);
  return 0;
}
int conv_tap_gather_strided(const float* z, int64_t ldz, int co_n, const void* bias, int T, int H, int W, void* out,
                            int out_dtype, int64_t cs, void* stream) {
  printf("svr2_conv_tap_gather %s %lld %d %s %d %d %d %s %d %s | %lld\n", P_(z), (long long)ldz, co_n, P_(bias), T, H, W, P_(out), out_dtype,
         P_(stream), (long long)cs);
  return 0;
}
}  // namespace svr2

// Arena fuzz: a random alloc / release / alloc_top sequence replayed on an unbounded arena (the dry run) and on one capped
// at the dry run's need() must make identical placement decisions, never overlap two live blocks and stay inside the cap.
static int arena_fuzz(unsigned seed, int ops) {
  struct Blk { size_t off, bytes; };
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed >> 8; };
  std::vector<int> script;           // >0: alloc of that many bytes; 0: release a pseudo-random live block; <0: alloc_top
  std::vector<size_t> pick;
  for (int i = 0; i < ops; ++i) {
    const unsigned r = rnd() % 100;
    if (r < 55) script.push_back(1 + (int)(rnd() % (1 << (4 + rnd() % 18))));
    else if (r < 95) script.push_back(0);
    else script.push_back(-(1 + (int)(rnd() % 100000)));
    pick.push_back(rnd());
  }
  auto replay = [&](Arena& A, std::vector<size_t>* trace) -> bool {
    std::vector<Blk> live, top;
    for (size_t i = 0; i < script.size(); ++i) {
      const int op = script[i];
      if (op > 0) {
        const size_t off = A.alloc((size_t)op);
        if (off == NONE) return false;
        const size_t bytes = align_up((size_t)op);
        for (const Blk& b : live)
          if (off < b.off + b.bytes && b.off < off + bytes) return false;          // overlap with a live block
        if (off + bytes > A.cap - A.top_used) return false;
        live.push_back({off, bytes});
        trace->push_back(off);
      } else if (op == 0) {
        if (live.empty()) continue;
        const size_t k = pick[i] % live.size();
        A.release(live[k].off, live[k].bytes);
        live.erase(live.begin() + k);
      } else {
        const size_t off = A.alloc_top((size_t)(-op));
        if (off == NONE) return false;
        trace->push_back(A.top_used);
        for (const Blk& b : live)
          if (b.off + b.bytes > A.cap - A.top_used) return false;                   // the top region ran into a live block
      }
    }
    return true;
  };
  Arena dry(~(size_t)0 / 2);
  std::vector<size_t> t0, t1;
  if (!replay(dry, &t0)) return 1;
  Arena real(dry.need());
  if (!replay(real, &t1)) return 2;
  if (t0 != t1) return 3;
  if (real.need() != dry.need()) return 4;
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && std::string(argv[1]) == "fuzz") {
    for (unsigned seed = 1; seed <= 200; ++seed) {
      const int rc = arena_fuzz(seed, 400);
      if (rc) { fprintf(stderr, "arena fuzz seed %u failed (%d)\n", seed, rc); return 10 + rc; }
    }
    printf("arena fuzz ok\n");
    return 0;
  }
  if (argc < 7) return 2;
  svr2_engine eng;
  eng.desc.variant = 2;
  std::ifstream f(argv[1]);
  std::string line;
  while (std::getline(f, line)) {       // name rank d0 d1 ...
    std::istringstream is(line);
    std::string name;
    Tensor t;
    is >> name >> t.rank;
    for (int i = 0; i < t.rank; ++i) is >> t.shape[i];
    t.ptr = (void*)0x1000;
    eng.w[name] = t;
  }
  const bool enc = std::string(argv[2]) == "enc";
  const int T = atoi(argv[3]), H = atoi(argv[4]), W = atoi(argv[5]), slice = atoi(argv[6]);
  const size_t need = svr2_vae_workspace_bytes(&eng, enc ? 0 : 1, T, H, W, slice);
  if (!need) { fprintf(stderr, "plan failed: %s\n", eng.err); return 3; }
  void* ws = nullptr;
  if (posix_memalign(&ws, 256, need)) return 4;
  g_lo = (char*)ws;
  g_hi = g_lo + need;
  static char in_buf[16], out_buf[16];
  const int rc = enc ? svr2_vae_encode(&eng, in_buf, 1, T, H, W, slice, out_buf, ws, need, nullptr)
                     : svr2_vae_decode(&eng, in_buf, 1, T, H, W, slice, out_buf, ws, need, nullptr);
  if (rc) { fprintf(stderr, "run failed (%d): %s\n", rc, eng.err); return 5; }
  // a workspace one byte short of the plan must be refused
  if ((enc ? svr2_vae_encode(&eng, in_buf, 1, T, H, W, slice, out_buf, ws, need - 256, nullptr)
           : svr2_vae_decode(&eng, in_buf, 1, T, H, W, slice, out_buf, ws, need - 256, nullptr)) == 0) return 6;
  printf("# workspace %zu touched_max_offset %zu launches %lld\n", need, g_touch, (long long)svr2_vae_last_launches(&eng));
  free(ws);
  vae_state_destroy(&eng);
  return 0;
}
