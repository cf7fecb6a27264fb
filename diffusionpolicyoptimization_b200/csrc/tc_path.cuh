// tcgen05 (bf16 x bf16 -> fp32, TMEM accumulators, TMA operand staging) path for large row counts.
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"

struct FwdBufs;
struct TcState { int dummy; };

static int tc_init(dppo_handle* h) { (void)h; return 0; }
static void tc_destroy(dppo_handle* h) { (void)h; }
static int tc_refresh_net(dppo_handle* h, int net, cudaStream_t s) { (void)h; (void)net; (void)s; return 0; }
static bool tc_eligible(const dppo_handle* h, int rows) { (void)h; (void)rows; return false; }
static int tc_actor_forward(dppo_handle* h, cudaStream_t s, int net, const float* x, const float* obs, int obs_div, int N,
                            const int* trow, int tconst, float* eps) {
    DPPO_FAIL(-7, "tensor path not built");
}
static int tc_actor_forward_keep(dppo_handle* h, cudaStream_t s, int net, const float* x, const float* obs, int obs_div, int N,
                                 const int* trow, int tconst, FwdBufs& b) {
    DPPO_FAIL(-7, "tensor path not built");
}
static int tc_ppo_step(dppo_handle* h, cudaStream_t s, const float* obs, const float* prev, const float* nxt, const int32_t* inds,
                       const float* returns, const float* oldvalues, const float* advantages, const float* oldlogp,
                       int N, int64_t N_global, float adv_mean, float adv_std) {
    DPPO_FAIL(-7, "tensor path not built");
}
