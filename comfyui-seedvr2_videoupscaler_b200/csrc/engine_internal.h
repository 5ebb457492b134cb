// Internal (non-ABI): the svr2_t handle shared by the native host runtimes (engine.cu: NaDiT, vae_engine.cu: video VAE).
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "svr2_internal.h"

namespace svr2 {

struct Tensor {
  void* ptr = nullptr;
  int dtype = 1;          // 0 f32, 1 bf16, 2 f16
  int rank = 0;
  int64_t shape[5] = {0, 0, 0, 0, 0};
  bool owned = false;
  int64_t numel() const {
    int64_t n = 1;
    for (int i = 0; i < rank; ++i) n *= shape[i];
    return n;
  }
};

inline size_t dtype_size(int dt) { return dt == 0 ? 4 : 2; }
inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct Geometry;      // window / RoPE index tables of one clip geometry (engine.cu)
struct VaeState;      // engine-owned VAE workspace bookkeeping (vae_engine.cu)

}  // namespace svr2

struct svr2_engine {
  int device = 0;
  svr2_model_desc desc{};
  std::unordered_map<std::string, svr2::Tensor> w;
  std::map<std::vector<int>, svr2::Geometry*> geo;      // (T, Hp, Wp, l) -> tables
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
  std::vector<void*> retired_workspaces;          // outgrown blocks: a captured CUDA graph may still replay into them
  svr2::VaeState* vae = nullptr;                  // VAE handles (desc.variant == 2): engine-owned workspace state
  char err[256] = "";
};

namespace svr2 {
inline int fail(svr2_engine* e, int code, const char* msg) {
  if (e) snprintf(e->err, sizeof e->err, "%s", msg);
  return set_error(code, msg);
}
inline const Tensor* find(svr2_engine* e, const std::string& name) {
  auto it = e->w.find(name);
  return it == e->w.end() ? nullptr : &it->second;
}
void vae_state_destroy(svr2_engine* e);          // vae_engine.cu
}  // namespace svr2
