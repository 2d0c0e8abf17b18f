// User-defined forward-model families, compiled at run time with NVRTC.
//
// Reference: NonLinearForwardModel.of(evaluation = <closure>, differential | jacobian = <closure>, ...)
//   /root/reference/src/main/scala/thylacine/model/components/forwardmodel/NonLinearForwardModel.scala:44-85.
// A JVM closure cannot run on a GPU.  The device-side family is f(theta)_i = link((A theta + offset)_i); round 1 offered a
// fixed registry of six links.  Here the caller hands the link -- and optionally its derivative -- as CUDA C source text in
// the variable z; it is compiled for sm_100a with NVRTC into the element-wise step of the wide DMMA path (the step between
// the forward GEMM Z = A X and the adjoint GEMM G = A^T W, kernels_wide.cu k_wide_link), so the two GEMMs stay on the fp64
// tensor cores and the user's code only ever sees one pre-activation at a time.  Without a derivative the link's own
// forward difference (link(z + h) - link(z)) / h with h = the model's `differential` is used: the same limit as the
// reference's FiniteDifferenceJacobian (core/computation/FiniteDifferenceJacobian.scala:30-55) for this family, since
// d f_i / d theta_j = link'(z_i) A_ij.
// NVRTC and the driver API are loaded with dlopen (libnvrtc.so.12, libcuda.so.1): models that never register a link need
// neither.  Compilation needs no GPU (the "does it build" check covers it); loading the module does.
#include "user_link.h"
#include <dlfcn.h>
#include <cstring>

namespace thy {

namespace {

typedef int nvrtcResult;
typedef struct _nvrtcProgram* nvrtcProgram;
struct NvrtcApi {
  void* lib = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
  bool ok = false;
};
NvrtcApi& nvrtc() {
  static NvrtcApi a;
  if (a.lib) return a;
  const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/targets/x86_64-linux/lib/libnvrtc.so.12",
                         "/usr/local/cuda/lib64/libnvrtc.so.12"};
  for (const char* n : names) {
    a.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (a.lib) break;
  }
  if (!a.lib) return a;
#define THY_SYM(field, name) a.field = (decltype(a.field))dlsym(a.lib, name)
  THY_SYM(CreateProgram, "nvrtcCreateProgram");
  THY_SYM(CompileProgram, "nvrtcCompileProgram");
  THY_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
  THY_SYM(GetProgramLog, "nvrtcGetProgramLog");
  THY_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
  THY_SYM(GetCUBIN, "nvrtcGetCUBIN");
  THY_SYM(DestroyProgram, "nvrtcDestroyProgram");
  THY_SYM(GetErrorString, "nvrtcGetErrorString");
#undef THY_SYM
  a.ok = a.CreateProgram && a.CompileProgram && a.GetProgramLogSize && a.GetProgramLog && a.GetCUBINSize && a.GetCUBIN &&
         a.DestroyProgram;
  return a;
}

typedef int CUresult;
struct DriverApi {
  void* lib = nullptr;
  CUresult (*ModuleLoadData)(void**, const void*) = nullptr;
  CUresult (*ModuleGetFunction)(void**, void*, const char*) = nullptr;
  CUresult (*ModuleUnload)(void*) = nullptr;
  CUresult (*LaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
};
DriverApi& driver() {
  static DriverApi a;
  if (a.lib) return a;
  a.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
  if (!a.lib) return a;
  a.ModuleLoadData = (decltype(a.ModuleLoadData))dlsym(a.lib, "cuModuleLoadData");
  a.ModuleGetFunction = (decltype(a.ModuleGetFunction))dlsym(a.lib, "cuModuleGetFunction");
  a.ModuleUnload = (decltype(a.ModuleUnload))dlsym(a.lib, "cuModuleUnload");
  a.LaunchKernel = (decltype(a.LaunchKernel))dlsym(a.lib, "cuLaunchKernel");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "cuGetErrorString");
  a.ok = a.ModuleLoadData && a.ModuleGetFunction && a.ModuleUnload && a.LaunchKernel;
  return a;
}

// The element-wise step of the wide path with the user's link spliced in.  Mirrors kernels_wide.cu k_wide_link: one warp
// per (chain, 128-observation tile) on the tiled pre-activations, in place; tile geometry arrives as arguments.
const char* kKernelTemplate = R"SRC(
__device__ __forceinline__ double thy_user_link(double z) { return (@EVAL@); }
__device__ __forceinline__ double thy_user_link_deriv(double z, double f, double h) { (void)f; (void)h; return (@DERIV@); }
extern "C" __global__ void __launch_bounds__(256) thy_user_link_step(
    double* __restrict__ WTl, const double* __restrict__ off, const double2* __restrict__ yiv, double* __restrict__ partS,
    long long B, int n, int n_tiles, int ictiles, int write_deriv, double fd_h, int WT, int WK, int WP, int TILE_KC) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= B * n_tiles) return;
  const long long b = wid / n_tiles;
  const int it = (int)(wid % n_tiles);
  const long long bt = b / WT;
  const int m = (int)(b % WT);
  double s = 0.0;
  for (int q = 0; q < WT / WK; ++q) {
    const int ic = it * (WT / WK) + q;
    const int i = ic * WK + lane;
    double* ptr = WTl + ((size_t)bt * ictiles + ic) * TILE_KC + m * WP + lane;
    double w = 0.0;
    if (i < n) {
      const double z = *ptr + off[i];
      const double f = thy_user_link(z);
      const double2 yv = yiv[i];
      const double res = yv.x - f;
      w = res * yv.y;
      s = fma(res, w, s);
      if (write_deriv) w *= thy_user_link_deriv(z, f, fd_h);
    }
    *ptr = w;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) partS[b * n_tiles + it] = s;
}
)SRC";

std::string splice(std::string src, const std::string& key, const std::string& text) {
  for (size_t pos = src.find(key); pos != std::string::npos; pos = src.find(key, pos + text.size()))
    src.replace(pos, key.size(), text);
  return src;
}

}  // namespace

UserLink::~UserLink() {
  if (module && driver().ok) driver().ModuleUnload(module);
}

// Compiles the link; the cubin is kept until the first launch loads it into the current context.
thy_status user_link_compile(const char* eval_expr, const char* deriv_expr, std::unique_ptr<UserLink>* out) {
  if (!eval_expr || !*eval_expr) return fail(THY_ERR_INVALID_ARGUMENT, "empty link expression");
  if (!nvrtc().ok) return fail(THY_ERR_UNSUPPORTED, "libnvrtc.so.12 could not be loaded: user-defined links are unavailable");
  auto link = std::make_unique<UserLink>();
  link->eval = eval_expr;
  link->has_deriv = deriv_expr && *deriv_expr;
  link->deriv = link->has_deriv ? deriv_expr : "(thy_user_link(z + h) - f) / h";
  const std::string src = splice(splice(kKernelTemplate, "@EVAL@", link->eval), "@DERIV@", link->deriv);
  nvrtcProgram prog = nullptr;
  nvrtcResult r = nvrtc().CreateProgram(&prog, src.c_str(), "thy_user_link.cu", 0, nullptr, nullptr);
  if (r != 0) return fail(THY_ERR_CUDA, "nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
  r = nvrtc().CompileProgram(prog, 4, opts);
  if (r != 0) {
    size_t n = 0;
    nvrtc().GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) nvrtc().GetProgramLog(prog, log.data());
    nvrtc().DestroyProgram(&prog);
    return fail(THY_ERR_INVALID_ARGUMENT, "user link does not compile (z is the pre-activation; f and h are also in scope in "
                                          "the derivative): " + log);
  }
  size_t n = 0;
  nvrtc().GetCUBINSize(prog, &n);
  link->cubin.resize(n);
  r = nvrtc().GetCUBIN(prog, link->cubin.data());
  nvrtc().DestroyProgram(&prog);
  if (r != 0 || n == 0) return fail(THY_ERR_CUDA, "nvrtcGetCUBIN failed");
  *out = std::move(link);
  return THY_OK;
}

thy_status user_link_launch(UserLink& link, double* WTl, const double* off, const double2* yiv, double* partS, long long B,
                            int n, int n_tiles, int ictiles, int write_deriv, double fd_h, int WT, int WK, int WP, int TILE_KC,
                            cudaStream_t s) {
  if (!driver().ok) return fail(THY_ERR_CUDA, "libcuda.so.1 could not be loaded");
  if (!link.function) {
    THY_CUDA_CHECK(cudaFree(nullptr));   // make sure the runtime's primary context is current
    if (driver().ModuleLoadData(&link.module, link.cubin.data()) != 0) return fail(THY_ERR_CUDA, "cuModuleLoadData failed for a user link");
    if (driver().ModuleGetFunction(&link.function, link.module, "thy_user_link_step") != 0)
      return fail(THY_ERR_CUDA, "cuModuleGetFunction failed for a user link");
  }
  void* args[] = {&WTl, &off, &yiv, &partS, &B, &n, &n_tiles, &ictiles, &write_deriv, &fd_h, &WT, &WK, &WP, &TILE_KC};
  const long long warps = B * n_tiles;
  const unsigned grid = (unsigned)((warps + 7) / 8);
  const CUresult r = driver().LaunchKernel(link.function, grid, 1, 1, 256, 1, 1, 0, (void*)s, args, nullptr);
  if (r != 0) {
    const char* msg = nullptr;
    if (driver().GetErrorString) driver().GetErrorString(r, &msg);
    return fail(THY_ERR_CUDA, std::string("cuLaunchKernel (user link): ") + (msg ? msg : "error"));
  }
  count_launch();
  return THY_OK;
}

}  // namespace thy
