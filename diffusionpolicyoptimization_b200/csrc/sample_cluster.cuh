// Persistent small-batch DDPM chain sampler (VPGDiffusion.call, diffusion_vpg.py:249-339).
//
// ONE launch per rollout step; all T denoising steps loop on-chip.  The grid is a set of 16-CTA
// thread-block clusters (one per GPC).  A cluster owns R env rows for the whole chain; inside a
// cluster every CTA owns 32 of the 512 hidden columns:
//   * its [512 x 32] fp32 slices of block.l1 / block.l2 and its [32 x A] slice of the output layer
//     stay resident in shared memory (reloaded once at the base -> fine-tuned switch, t = K-1);
//   * its 512 threads each keep one column of the layer-0 x-rows (W_in[0:A, c]) and the
//     obs contribution  obs @ W_in[A+td:, c]  in registers (the obs term is constant over the chain);
//   * the time embedding enters as a per-t bias row bt[t] (precomputed per weight version).
// Per denoising step: L0 (local, redundant) -> L1 slice -> DSMEM all-gather of relu(h1) ->
// cluster barrier -> L2 slice (+ residual) -> partial output layer over the CTA's 32 columns ->
// DSMEM all-gather of the [R x A] partials -> cluster barrier -> posterior / noise / x update
// (redundant, bit-identical in all 16 CTAs).  Two cluster barriers per step, 2T per chain.
// Strict fp32 FFMA; noise is injected (parity) or drawn in-kernel by Philox4x32-10.
#pragma once
#include "simt_kernels.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

constexpr int CS_H = 512;        // hidden width this kernel is specialised for
constexpr int CS_C = 16;         // CTAs per cluster
constexpr int CS_COLS = CS_H / CS_C;   // 32 hidden columns per CTA
constexpr int CS_THREADS = 512;

struct ClusterSampleP {
    const float* w[2];        // flat actor parameters: [0] base, [1] fine-tuned
    const float* bt[2];       // per-t layer-0 bias rows [T][H]
    ActorOff o;
    const float* obs; const float* xT; const float* noise;
    float* actions; float* chains;
    const float* sch;
    int B, A, Do, T, K, td, nchunks, use_base_policy;
    SampleHyper hp;
    uint64_t seed, offset; int64_t row_offset;
    // env-side epilogue: the first act_cols = act_steps * Da entries of every action chunk, un-normalised like
    // MujocoLocomotionLowdimWrapper.unnormalize_action, written where the env workers read them (pinned host memory or device)
    float* raw_actions; const float* act_min; const float* act_max; int act_cols, Da;
};
// env/gym_utils/wrapper/mujoco_locomotion_lowdim.py:60-62 in NumPy's fp32 evaluation order (no FMA contraction)
__device__ __forceinline__ float env_unnormalize_action(float a, float amin, float amax) {
    const float a01 = __fdiv_rn(__fadd_rn(a, 1.f), 2.f);
    return __fadd_rn(__fmul_rn(a01, __fsub_rn(amax, amin)), amin);
}

template <int AP, int R>
constexpr size_t cluster_sample_smem_floats() {
    return (size_t)2 * CS_H * CS_COLS      // W1s, W2s
         + (size_t)CS_COLS * AP            // W3s
         + (size_t)3 * R * CS_H            // actA, actB, red
         + (size_t)3 * R * CS_COLS         // uown, out1, vs
         + (size_t)R * AP * (2 + CS_C)     // pl, xs, part[16]
         + 2 * CS_COLS + AP;               // b1s, b2s, b3s
}

template <int AP, int R>
__global__ void __launch_bounds__(CS_THREADS, 1) sample_cluster_kernel(const ClusterSampleP p) {
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int cid = blockIdx.x / CS_C, nclusters = gridDim.x / CS_C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col0 = crank * CS_COLS;
    const int A = p.A, T = p.T, K = p.K, B = p.B;

    extern __shared__ __align__(16) float smem[];
    float* W1s = smem;                         // [512][32]
    float* W2s = W1s + CS_H * CS_COLS;         // [512][32]
    float* W3s = W2s + CS_H * CS_COLS;         // [32][AP]
    float* actA = W3s + CS_COLS * AP;          // [R][512] relu(u)
    float* actB = actA + R * CS_H;             // [R][512] relu(h1)
    float* red = actB + R * CS_H;              // [16][R][32] cross-warp partials (also obs staging)
    float* uown = red + R * CS_H;              // [R][32] raw u, own columns (residual)
    float* out1 = uown + R * CS_COLS;          // [R][32]
    float* vs = out1 + R * CS_COLS;            // [R][32]
    float* pl = vs + R * CS_COLS;              // [R][AP] partial eps of this CTA
    float* xs = pl + R * AP;                   // [R][AP] current x
    float* part = xs + R * AP;                 // [16][R][AP]
    float* b1s = part + CS_C * R * AP;         // [32]
    float* b2s = b1s + CS_COLS;                // [32]
    float* b3s = b2s + CS_COLS;                // [AP]

    float wx[AP];        // W_in[a][tid]
    float cobs[R];       // obs[r] @ W_in[A+td:, tid]

    for (int chunk = cid; chunk < p.nchunks; chunk += nclusters) {
        const int row0 = chunk * R;
        // ---- initialise x (x_T injected or Philox slot 0)
        if (tid < R * AP) {
            int r = tid / AP, a = tid % AP;
            float v = 0.f;
            if (a < A && row0 + r < B) {
                v = p.xT ? p.xT[(size_t)(row0 + r) * A + a]
                         : philox_normal(p.seed, p.offset, p.row_offset + row0 + r, 0, a);
                if (crank == 0 && p.chains && K == T) p.chains[((size_t)(row0 + r) * (K + 1)) * A + a] = v;
            }
            xs[tid] = v;
        }
        __syncthreads();
        int cur_net = -1;

        for (int i = 0; i < T; ++i) {
            const int t = T - 1 - i;
            const int net = (t < K && !p.use_base_policy) ? 1 : 0;
            if (net != cur_net) {
                // ---- (re)load this net's resident slices and recompute its obs term
                const float* w = p.w[net];
                for (int idx = tid; idx < CS_H * CS_COLS; idx += CS_THREADS) {
                    int k = idx >> 5, j = idx & 31;
                    W1s[idx] = w[p.o.w1 + (size_t)k * CS_H + col0 + j];
                    W2s[idx] = w[p.o.w2 + (size_t)k * CS_H + col0 + j];
                }
                for (int idx = tid; idx < CS_COLS * AP; idx += CS_THREADS) {
                    int k = idx / AP, a = idx % AP;
                    W3s[idx] = a < A ? w[p.o.w3 + (size_t)(col0 + k) * A + a] : 0.f;
                }
                if (tid < CS_COLS) { b1s[tid] = w[p.o.b1 + col0 + tid]; b2s[tid] = w[p.o.b2 + col0 + tid]; }
                if (tid < AP) b3s[tid] = tid < A ? w[p.o.b3 + tid] : 0.f;
#pragma unroll
                for (int a = 0; a < AP; ++a) wx[a] = a < A ? w[p.o.win + (size_t)a * CS_H + tid] : 0.f;
#pragma unroll
                for (int r = 0; r < R; ++r) cobs[r] = 0.f;
                for (int j = 0; j < p.Do; ++j) {
                    float wv = w[p.o.win + (size_t)(A + p.td + j) * CS_H + tid];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float ov = (row0 + r < B) ? p.obs[(size_t)(row0 + r) * p.Do + j] : 0.f;
                        cobs[r] = fmaf(ov, wv, cobs[r]);
                    }
                }
                cur_net = net;
                __syncthreads();
            }
            // ---- L0 (every CTA computes all 512 columns): u = [x|obs] @ W_in + bt[t]
            {
                const float btv = p.bt[net][(size_t)t * CS_H + tid];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float u = cobs[r] + btv;
#pragma unroll
                    for (int a = 0; a < AP; ++a) u = fmaf(xs[r * AP + a], wx[a], u);
                    actA[r * CS_H + tid] = fmaxf(u, 0.f);
                    if (tid >= col0 && tid < col0 + CS_COLS) uown[r * CS_COLS + tid - col0] = u;
                }
            }
            __syncthreads();
            // ---- L1 slice: h1[:, col0:col0+32] = relu(u) @ W1[:, cols] + b1
            {
                float acc[R];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = 0.f;
                const float* Wp = W1s + (warp * 32) * CS_COLS + lane;
                const float* ap = actA + warp * 32;
#pragma unroll
                for (int kk = 0; kk < 32; kk += 4) {
                    float w0 = Wp[(kk + 0) * CS_COLS], w1 = Wp[(kk + 1) * CS_COLS];
                    float w2 = Wp[(kk + 2) * CS_COLS], w3 = Wp[(kk + 3) * CS_COLS];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float4 a = *reinterpret_cast<const float4*>(ap + r * CS_H + kk);
                        acc[r] = fmaf(a.x, w0, acc[r]); acc[r] = fmaf(a.y, w1, acc[r]);
                        acc[r] = fmaf(a.z, w2, acc[r]); acc[r] = fmaf(a.w, w3, acc[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) red[(warp * R + r) * 32 + lane] = acc[r];
            }
            __syncthreads();
            if (tid < R * 32) {
                int r = tid >> 5;
                float s = b1s[lane];
#pragma unroll
                for (int w = 0; w < 16; ++w) s += red[(w * R + r) * 32 + lane];
                out1[r * 32 + lane] = fmaxf(s, 0.f);
            }
            __syncthreads();
            {   // all-gather relu(h1): warp w writes this CTA's slice into CTA w's actB
                float* peer = cluster.map_shared_rank(actB, warp);
#pragma unroll
                for (int r = 0; r < R; ++r) peer[r * CS_H + col0 + lane] = out1[r * 32 + lane];
            }
            cluster.sync();
            // ---- L2 slice + residual: v[:, cols] = relu(h1) @ W2[:, cols] + b2 + u[:, cols]
            {
                float acc[R];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = 0.f;
                const float* Wp = W2s + (warp * 32) * CS_COLS + lane;
                const float* ap = actB + warp * 32;
#pragma unroll
                for (int kk = 0; kk < 32; kk += 4) {
                    float w0 = Wp[(kk + 0) * CS_COLS], w1 = Wp[(kk + 1) * CS_COLS];
                    float w2 = Wp[(kk + 2) * CS_COLS], w3 = Wp[(kk + 3) * CS_COLS];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float4 a = *reinterpret_cast<const float4*>(ap + r * CS_H + kk);
                        acc[r] = fmaf(a.x, w0, acc[r]); acc[r] = fmaf(a.y, w1, acc[r]);
                        acc[r] = fmaf(a.z, w2, acc[r]); acc[r] = fmaf(a.w, w3, acc[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) red[(warp * R + r) * 32 + lane] = acc[r];
            }
            __syncthreads();
            if (tid < R * 32) {
                int r = tid >> 5;
                float s = b2s[lane];
#pragma unroll
                for (int w = 0; w < 16; ++w) s += red[(w * R + r) * 32 + lane];
                vs[r * 32 + lane] = s + uown[r * 32 + lane];
            }
            __syncthreads();
            // ---- partial output layer over this CTA's 32 columns
            if (tid < R * AP) {
                int r = tid / AP, a = tid % AP;
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < CS_COLS; ++k) s = fmaf(vs[r * 32 + k], W3s[k * AP + a], s);
                pl[tid] = s;
            }
            __syncthreads();
            {
                float* peer = cluster.map_shared_rank(part, warp) + crank * (R * AP);
                for (int j = lane; j < R * AP; j += 32) peer[j] = pl[j];
            }
            cluster.sync();
            // ---- posterior mean / clipped noise / x update (identical in all 16 CTAs)
            if (tid < R * AP) {
                int r = tid / AP, a = tid % AP;
                if (a < A && row0 + r < B) {
                    float e = b3s[a];
#pragma unroll
                    for (int src = 0; src < CS_C; ++src) e += part[src * (R * AP) + tid];
                    const int64_t row = row0 + r;
                    float nz = p.noise ? p.noise[((size_t)i * B + row) * A + a]
                                       : philox_normal(p.seed, p.offset, p.row_offset + row, 1 + i, a);
                    float xn = ddpm_step_elem(xs[tid], e, nz, t, p.sch, T, p.hp, t == 0);
                    xs[tid] = xn;
                    if (crank == 0) {
                        if (p.chains && t <= K) p.chains[((size_t)row * (K + 1) + (K - t)) * A + a] = xn;
                        if (t == 0) {
                            p.actions[(size_t)row * A + a] = xn;
                            if (p.raw_actions && a < p.act_cols)
                                p.raw_actions[(size_t)row * p.act_cols + a] = env_unnormalize_action(xn, p.act_min[a % p.Da], p.act_max[a % p.Da]);
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
    // no CTA may exit while peers can still write into its shared memory
    cluster.sync();
}
