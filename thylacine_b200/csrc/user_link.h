// NVRTC-compiled user links (user_link.cu).
#pragma once
#include "common.cuh"
#include <memory>
#include <string>
#include <vector>

namespace thy {

struct UserLink {
  std::string eval, deriv;
  bool has_deriv = false;
  std::vector<char> cubin;
  void* module = nullptr;     // CUmodule, loaded at the first launch
  void* function = nullptr;   // CUfunction thy_user_link_step
  UserLink() = default;
  UserLink(const UserLink&) = delete;
  UserLink& operator=(const UserLink&) = delete;
  ~UserLink();
};

thy_status user_link_compile(const char* eval_expr, const char* deriv_expr, std::unique_ptr<UserLink>* out);
thy_status user_link_launch(UserLink& link, double* WTl, const double* off, const double2* yiv, double* partS, long long B,
                            int n, int n_tiles, int ictiles, int write_deriv, double fd_h, int WT, int WK, int WP, int TILE_KC,
                            cudaStream_t s);

}  // namespace thy
