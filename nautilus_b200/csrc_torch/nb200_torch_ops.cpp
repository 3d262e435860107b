// torch.ops.nautilus_b200.* -- the hot-path entry points registered through
// PyTorch's STABLE C ABI (torch/csrc/stable/library.h: boxed kernels over
// StableIValue stacks, tensors as AtenTensorHandle; no ATen / c10 C++ ABI
// dependence), as SURVEY.md 8(b) asks.  Every op is a thin shim: it checks
// its tensors, allocates the outputs with the caching allocator, fetches the
// current CUDA stream of the tensors' device through the C shim
// (aoti_torch_get_current_cuda_stream) and calls the plain C ABI of
// include/nautilus_b200.h -- the same functions the ctypes binding uses.
//
//   ell_contains(points, c, B_inv) -> u8[n]        Ellipsoid.contains, basic.py:344-360
//   shell_stats(log_l, log_l_min) -> (lse f64[4], counters i64[8])
//                                                   update_shell_info sums, sampler.py:925-943
//   bound_contains(meta_h, meta, data, bound, which, points, mlp_mode) -> u8[n]
//                                                   NautilusBound.contains, nautilus.py:146-169
//   shell_cycle(meta_h, meta, data, bound, first_later, n_later, n, seed,
//               offset, stream_id, like_id, like_params, log_l_min, mlp_mode)
//       -> (points f64[n,d], log_l f64[n], code u8[n], lse f64[4], counters i64[8])
//                                                   add_samples over one raw batch, sampler.py:1093-1144
#include <torch/csrc/inductor/aoti_torch/c/shim.h>
#include <torch/csrc/stable/library.h>
#include <torch/csrc/stable/ops.h>
#include <torch/csrc/stable/tensor.h>

#include <dlfcn.h>

#include <string>
#include <tuple>

#include "../../include/nautilus_b200.h"

using torch::headeronly::ScalarType;
using torch::stable::Tensor;

namespace {

// The C ABI is bound at run time (dlopen of libnautilus_b200.so next to this
// file, dlsym per entry point) instead of at link time: the nvcc-built
// library carries pieces of a statically linked C++ runtime, and having it
// in this module's link-time lookup scope broke exception unwinding out of
// the boxed kernels (a failed STD_TORCH_CHECK crashed instead of raising).
void* core_library() {
  static void* handle = [] {
    Dl_info info;
    STD_TORCH_CHECK(dladdr(reinterpret_cast<void*>(&core_library), &info) &&
                        info.dli_fname,
                    "cannot locate libnautilus_b200_torch.so");
    std::string path(info.dli_fname);
    path = path.substr(0, path.find_last_of('/') + 1) + "libnautilus_b200.so";
    void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    STD_TORCH_CHECK(h != nullptr, "cannot load ", path, ": ", dlerror());
    return h;
  }();
  return handle;
}

void* core_symbol(const char* name) {
  void* p = dlsym(core_library(), name);
  STD_TORCH_CHECK(p != nullptr, "libnautilus_b200.so does not export ", name);
  return p;
}

// NB(nb200_cycle)(...) calls the entry point with the header's signature
#define NB(fn)                                                          \
  ([]() {                                                               \
    static auto* p = reinterpret_cast<decltype(&fn)>(core_symbol(#fn)); \
    return p;                                                           \
  }())

void* cuda_stream_of(const Tensor& t) {
  void* stream = nullptr;
  TORCH_ERROR_CODE_CHECK(
      aoti_torch_get_current_cuda_stream(t.get_device_index(), &stream));
  return stream;
}

void check(int rc) { STD_TORCH_CHECK(rc == 0, NB(nb200_last_error)()); }

void want(const Tensor& t, ScalarType dtype, bool cuda, const char* what) {
  STD_TORCH_CHECK(t.scalar_type() == dtype && t.is_contiguous() &&
                      (cuda ? t.is_cuda() : t.is_cpu()),
                  what);
}

Tensor ell_contains(Tensor points, Tensor c, Tensor B_inv) {
  want(points, ScalarType::Double, true,
       "points must be a contiguous CUDA float64 [n, d] tensor");
  STD_TORCH_CHECK(points.dim() == 2, "points must be [n, d]");
  const int64_t n = points.size(0), d = points.size(1);
  want(c, ScalarType::Double, true, "c must be CUDA float64 [d]");
  want(B_inv, ScalarType::Double, true, "B_inv must be CUDA float64 [d, d]");
  STD_TORCH_CHECK(c.numel() == d && B_inv.numel() == d * d,
                  "ellipsoid parameters do not match the points");
  Tensor out = torch::stable::new_empty(points, {n}, ScalarType::Byte);
  check(NB(nb200_ell_contains)(
      static_cast<const double*>(points.data_ptr()), n, (int)d,
      static_cast<const double*>(c.data_ptr()),
      static_cast<const double*>(B_inv.data_ptr()),
      static_cast<uint8_t*>(out.data_ptr()), nullptr, cuda_stream_of(points)));
  return out;
}

std::tuple<Tensor, Tensor> shell_stats(Tensor log_l, double log_l_min) {
  want(log_l, ScalarType::Double, true, "log_l must be CUDA float64 [n]");
  Tensor lse = torch::stable::new_empty(log_l, {NB200_N_LSE});
  Tensor counters =
      torch::stable::new_empty(log_l, {NB200_N_CNT}, ScalarType::Long);
  const int64_t bytes = (int64_t)NB(nb200_workspace_bytes)(1, 1);
  Tensor ws = torch::stable::new_empty(log_l, {bytes}, ScalarType::Byte);
  check(NB(nb200_stats)(static_cast<const double*>(log_l.data_ptr()), nullptr,
                    log_l.numel(), log_l_min,
                    static_cast<double*>(lse.data_ptr()),
                    static_cast<int64_t*>(counters.data_ptr()), ws.data_ptr(),
                    (size_t)bytes, cuda_stream_of(log_l)));
  return {lse, counters};
}

Tensor bound_contains(Tensor meta_h, Tensor meta, Tensor data, int64_t bound,
                      int64_t which, Tensor points, int64_t mlp_mode) {
  want(meta_h, ScalarType::Int, false, "meta_h must be a CPU int32 tensor");
  want(meta, ScalarType::Int, true, "meta must be a CUDA int32 tensor");
  want(data, ScalarType::Double, true, "data must be a CUDA float64 tensor");
  want(points, ScalarType::Double, true,
       "points must be a contiguous CUDA float64 [n, d] tensor");
  STD_TORCH_CHECK(points.dim() == 2, "points must be [n, d]");
  const int64_t n = points.size(0), d = points.size(1);
  Tensor out = torch::stable::new_empty(points, {n}, ScalarType::Byte);
  const int64_t bytes = (int64_t)NB(nb200_workspace_bytes)(n, (int)d);
  Tensor ws = torch::stable::new_empty(points, {bytes}, ScalarType::Byte);
  check(NB(nb200_bound_contains)(
      static_cast<const int32_t*>(meta_h.data_ptr()),
      static_cast<const int32_t*>(meta.data_ptr()),
      static_cast<const double*>(data.data_ptr()), (int)bound, (int)which,
      static_cast<const double*>(points.data_ptr()), nullptr, n,
      static_cast<uint8_t*>(out.data_ptr()), (int)mlp_mode, ws.data_ptr(),
      (size_t)bytes, cuda_stream_of(points)));
  return out;
}

std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor> shell_cycle(
    Tensor meta_h, Tensor meta, Tensor data, int64_t bound,
    int64_t first_later, int64_t n_later, int64_t n, int64_t seed,
    int64_t offset, int64_t stream_id, int64_t like_id, Tensor like_params,
    double log_l_min, int64_t mlp_mode) {
  want(meta_h, ScalarType::Int, false, "meta_h must be a CPU int32 tensor");
  want(meta, ScalarType::Int, true, "meta must be a CUDA int32 tensor");
  want(data, ScalarType::Double, true, "data must be a CUDA float64 tensor");
  want(like_params, ScalarType::Double, true,
       "like_params must be a CUDA float64 tensor");
  const int32_t* mh = static_cast<const int32_t*>(meta_h.data_ptr());
  STD_TORCH_CHECK(meta_h.numel() > 1 && bound >= 0 && bound < mh[0],
                  "bound index out of range");
  const int d = mh[mh[1 + bound] + 2];
  int pairs = 0;                       // (later bound, neural bound) pairs
  for (int64_t l = first_later; l < first_later + n_later && l < mh[0]; ++l)
    pairs += mh[mh[1 + l] + 4];
  const int64_t rows = n > 0 ? n : 1;
  Tensor points = torch::stable::new_empty(data, {rows, (int64_t)d});
  Tensor log_l = torch::stable::new_empty(data, {rows});
  Tensor code = torch::stable::new_empty(data, {rows}, ScalarType::Byte);
  Tensor lse = torch::stable::new_empty(data, {NB200_N_LSE});
  Tensor counters =
      torch::stable::new_empty(data, {NB200_N_CNT}, ScalarType::Long);
  const int64_t bytes = (int64_t)NB(nb200_cycle_workspace_bytes)(rows, d, pairs);
  Tensor ws = torch::stable::new_empty(data, {bytes}, ScalarType::Byte);
  check(NB(nb200_cycle)(
      mh, static_cast<const int32_t*>(meta.data_ptr()),
      static_cast<const double*>(data.data_ptr()), (int)bound,
      (int)first_later, (int)n_later, n, (uint64_t)seed, (uint64_t)offset,
      (uint32_t)stream_id, (int)like_id,
      static_cast<const double*>(like_params.data_ptr()),
      (int)like_params.numel(), log_l_min, (int)mlp_mode,
      static_cast<double*>(points.data_ptr()),
      static_cast<double*>(log_l.data_ptr()),
      static_cast<uint8_t*>(code.data_ptr()),
      static_cast<double*>(lse.data_ptr()),
      static_cast<int64_t*>(counters.data_ptr()), ws.data_ptr(), (size_t)bytes,
      cuda_stream_of(data)));
  return {points, log_l, code, lse, counters};
}

}  // namespace

STABLE_TORCH_LIBRARY(nautilus_b200, m) {
  m.def("ell_contains(Tensor points, Tensor c, Tensor B_inv) -> Tensor");
  m.def("shell_stats(Tensor log_l, float log_l_min) -> (Tensor, Tensor)");
  m.def(
      "bound_contains(Tensor meta_h, Tensor meta, Tensor data, int bound, "
      "int which, Tensor points, int mlp_mode) -> Tensor");
  m.def(
      "shell_cycle(Tensor meta_h, Tensor meta, Tensor data, int bound, "
      "int first_later, int n_later, int n, int seed, int offset, "
      "int stream_id, int like_id, Tensor like_params, float log_l_min, "
      "int mlp_mode) -> (Tensor, Tensor, Tensor, Tensor, Tensor)");
}

STABLE_TORCH_LIBRARY_IMPL(nautilus_b200, CUDA, m) {
  m.impl("ell_contains", TORCH_BOX(&ell_contains));
  m.impl("shell_stats", TORCH_BOX(&shell_stats));
  m.impl("bound_contains", TORCH_BOX(&bound_contains));
  m.impl("shell_cycle", TORCH_BOX(&shell_cycle));
}
