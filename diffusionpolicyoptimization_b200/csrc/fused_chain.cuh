// Fused residual-MLP layer chain on tcgen05 (sm_100a): one persistent CTA per SM walks 128-row tiles and
// runs EVERY layer of the chain for its tile with the activations resident on chip.
//
//   shared memory  X   [128][H] bf16   the current activation, stored as H/64 K-major 128B-swizzled tiles
//                                      (exactly the image TMA would produce), i.e. directly the A operand
//                                      of the next layer's tcgen05.mma and the source of a TMA store
//                  H0  [128][64] bf16  the packed layer-0 input tile [x | obs | onehot(t) | 1 | 0]
//                                      (TMA-loaded, or rebuilt in place by the sampler every denoising step)
//                  W   4 x 16 KB ring  weight tiles [32 k][256 n] (MN-major, TMA) streamed from L2
//   tensor memory  512 columns         the whole [128][H] fp32 accumulator of one layer
//
//   warp 0     TMA producer (weight ring, H0 tiles)
//   warp 1     tcgen05.mma issuer (one elected thread), TMEM allocation
//   warps 2-5  epilogue: tcgen05.ld -> bias / ReLU / ReLU-mask -> bf16 -> swizzled st.shared into X
//              (+ TMA store of X to HBM when the backward pass needs the tensor), and the final-layer
//              math: eps store, Gaussian log-prob, or the DDPM posterior step of the sampler.
//
// MMA and epilogue of one tile alternate (X is updated in place); the weight ring keeps streaming across
// the epilogue, so the tensor pipe restarts immediately.  Three programs run on this engine:
//   forward   L0 relu(H0 W0) -> L1 relu(X W1 + b1) -> L2 X W2 + H0 W0 + b2 (residual by K-concatenation)
//             -> L3 X W3 + b3 ;  final = eps | log-prob | (training) eps + stored a0, a1, v + ReLU bit masks
//   backward  B1 dv = deps W3^T -> B2 dh1 = (dv W2^T) . m1 -> B3 du = (dh1 W1^T) . m0   (each stored by TMA)
//   sampler   T denoising steps x forward, x kept in the epilogue threads' registers, noise from Philox
//             (or injected), chain / actions written once: ONE launch per rollout step.
#pragma once
#include "tc_gemm.cuh"
#include "simt_kernels.cuh"

namespace fc {
using namespace tc;

constexpr int FBM = 128, WK = 32, NSTAGE = 4, FTHREADS = 192, MAXL = 4, NMAPS = 8;
constexpr int STAGE_BYTES = WK * 256 * 2;      // 16 KB
constexpr int H0_BYTES = FBM * 64 * 2;         // 16 KB

enum { FINAL_STORE = 0, FINAL_EPS = 1, FINAL_LOGP = 2, FINAL_SAMPLE = 3 };

struct Layer {
    int a_src;                 // 0: H0 tile only (K = 64); 1: X (K = H); 2: X then H0 (K = H + 64)
    int wmap;                  // tensor map of the B operand: bf16 [K][N] (MN-major), box [WK][64]
    int wrow_x, wrow_h0;       // first k-row of the X part / the H0 part inside that map
    int n;                     // output width: H, or 64 for the output layer
    int h0_last;               // this layer is the tile's last reader of H0 (TMA mode): release it afterwards
    const float* bias;         // [n] fp32 or null
    int act;                   // 0 none, 1 ReLU
    const uint32_t* mask_in;   // [rows][H/32] ReLU bit masks to multiply with, or null
    uint32_t* mask_out;        // [rows][H/32] ReLU bit masks to record, or null
    int store_map;             // tensor map for the TMA store of X ([rows][H] bf16, box [128][64]), or -1
};

struct Params {
    int rows, nlayers, final_mode, h0_from_tma;
    Layer L[2][MAXL];          // [net][layer]; net 1 = fine-tuned actor (sampler only)
    int A, Do, T, K;
    const float* sch;
    // FINAL_EPS / FINAL_LOGP
    float* out;                // eps [rows][A] or logp [rows][A]
    const float* prev; const float* next; const float* chains; const int* trow;
    float dcv, min_lp_std;
    // FINAL_SAMPLE
    const float* obs; const float* xT; const float* noise;
    float* actions; float* chains_out;
    SampleHyper hp; int use_base_policy;
    uint64_t seed, offset; int64_t row_offset;
};
struct Maps { CUtensorMap m[NMAPS]; };

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

template <int H> constexpr size_t chain_smem_bytes() {
    return (size_t)(H / 64) * 16384 + H0_BYTES + (size_t)NSTAGE * STAGE_BYTES + 1024 + 256;
}

// write this thread's row of the H0 tile: [x (A) | obs (Do) | onehot(t) (T) | 1 | 0..], 128B-swizzled
__device__ __forceinline__ void build_h0_row(uint32_t h0_addr, int rloc, const float (&x)[32], const float* __restrict__ obs_row,
                                             int A, int Do, int T, int t, bool valid) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = c * 8 + j * 2 + e;
                float v = 0.f;
                if (valid) {
                    if (k < 32 && k < A) v = x[k < 32 ? k : 0];
                    else if (k < A + Do) v = __ldg(obs_row + (k - A));
                    else if (k < A + Do + T) v = (k - A - Do == t) ? 1.f : 0.f;
                    else if (k == A + Do + T) v = 1.f;
                }
                v2[e] = v;
            }
            w[j] = pack_bf16(v2[0], v2[1]);
        }
        st_shared_v4(h0_addr + rloc * 128 + ((c ^ (rloc & 7)) << 4), w[0], w[1], w[2], w[3]);
    }
}

template <int H>
__global__ void __launch_bounds__(FTHREADS, 1) chain_kernel(const __grid_constant__ Maps maps, const Params p) {
    constexpr int XT = H / 64;
    constexpr int X_BYTES = XT * 16384;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sX = smem;
    uint8_t* sH0 = sX + X_BYTES;
    uint8_t* sW = sH0 + H0_BYTES;
    uint64_t* bars = (uint64_t*)(sW + NSTAGE * STAGE_BYTES);
    uint64_t* w_full = bars;
    uint64_t* w_empty = bars + NSTAGE;
    uint64_t* h0_full = bars + 2 * NSTAGE;
    uint64_t* h0_empty = h0_full + 1;
    uint64_t* acc_full = h0_full + 2;
    uint64_t* x_full = h0_full + 3;
    uint32_t* tmem_slot = (uint32_t*)(h0_full + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (p.rows + FBM - 1) / FBM;
    const bool sampler = p.final_mode == FINAL_SAMPLE;
    const int nsteps = sampler ? p.T : 1;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < NMAPS; ++i) tma_prefetch_desc(&maps.m[i]);
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        mbar_init(h0_full, 1); mbar_init(h0_empty, 1); mbar_init(acc_full, 1); mbar_init(x_full, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, h0_phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                if (p.h0_from_tma) {
                    mbar_wait(h0_empty, h0_phase ^ 1);
                    mbar_expect_tx(h0_full, H0_BYTES);
                    tma_load_2d(sH0, &maps.m[0], h0_full, 0, tile * FBM);
                    h0_phase ^= 1;
                }
                for (int step = 0; step < nsteps; ++step) {
                    const int net = (sampler && (p.T - 1 - step) < p.K && !p.use_base_policy) ? 1 : 0;
                    for (int l = 0; l < p.nlayers; ++l) {
                        const Layer& L = p.L[net][l];
                        const int n_cur = L.n < 256 ? L.n : 256, nhc = L.n / n_cur;
                        const int kx = L.a_src >= 1 ? H / WK : 0, kh = L.a_src != 1 ? 64 / WK : 0;
                        for (int nh = 0; nh < nhc; ++nh) {
                            for (int s = 0; s < kx + kh; ++s) {
                                const int krow = s < kx ? L.wrow_x + s * WK : L.wrow_h0 + (s - kx) * WK;
                                mbar_wait(&w_empty[stage], phase ^ 1);
                                mbar_expect_tx(&w_full[stage], (uint32_t)(WK * n_cur * 2));
                                uint8_t* dst = sW + stage * STAGE_BYTES;
                                for (int j = 0; j < n_cur / 64; ++j)
                                    tma_load_2d(dst + j * (WK * 128), &maps.m[L.wmap], &w_full[stage], nh * n_cur + j * 64, krow);
                                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, h0_phase = 0, x_phase = 0;
            const uint32_t x_addr = smem_u32(sX), h0_addr = smem_u32(sH0);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                bool h0_waited = false;
                for (int step = 0; step < nsteps; ++step) {
                    const int net = (sampler && (p.T - 1 - step) < p.K && !p.use_base_policy) ? 1 : 0;
                    for (int l = 0; l < p.nlayers; ++l) {
                        const Layer& L = p.L[net][l];
                        const int n_cur = L.n < 256 ? L.n : 256, nhc = L.n / n_cur;
                        const int kx = L.a_src >= 1 ? H / WK : 0, kh = L.a_src != 1 ? 64 / WK : 0;
                        const uint32_t idesc = make_idesc(FBM, n_cur, false, true);
                        // previous epilogue done: X / H0 written, TMEM drained
                        mbar_wait(x_full, x_phase); x_phase ^= 1;
                        if (L.a_src != 1 && p.h0_from_tma && !h0_waited) { mbar_wait(h0_full, h0_phase); h0_phase ^= 1; h0_waited = true; }
                        tcgen05_fence_after();
                        for (int nh = 0; nh < nhc; ++nh) {
                            const uint32_t tmem_d = tmem_base + (uint32_t)(nh * 256);
                            for (int s = 0; s < kx + kh; ++s) {
                                mbar_wait(&w_full[stage], phase);
                                tcgen05_fence_after();
                                const uint32_t b0 = smem_u32(sW + stage * STAGE_BYTES);
#pragma unroll
                                for (int q = 0; q < WK / 16; ++q) {
                                    uint32_t a_at;
                                    if (s < kx) { const int k = s * WK + q * 16; a_at = x_addr + (k >> 6) * 16384 + (k & 63) * 2; }
                                    else { const int k = (s - kx) * WK + q * 16; a_at = h0_addr + k * 2; }
                                    const uint64_t da = make_desc(a_at, 16, 1024);
                                    const uint64_t db = make_desc(b0 + q * 2048, WK * 128, 1024);
                                    umma_bf16(tmem_d, da, db, idesc, (s > 0 || q > 0) ? 1u : 0u);
                                }
                                tcgen05_commit(&w_empty[stage]);
                                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                            }
                        }
                        tcgen05_commit(acc_full);
                        if (L.h0_last && p.h0_from_tma) tcgen05_commit(h0_empty);
                    }
                }
            }
        }
    } else {
        // ===================================================== epilogue warps
        const int quad = warp & 3;
        const int rloc = quad * 32 + lane;                       // row inside the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const uint32_t x_addr = smem_u32(sX), h0_addr = smem_u32(sH0);
        const bool store_thread = (warp == 2 && lane == 0);
        uint32_t acc_phase = 0;
        const int A = p.A;
        float x[32];
#pragma unroll
        for (int a = 0; a < 32; ++a) x[a] = 0.f;
        bool first = true;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int row = tile * FBM + rloc;
            const bool valid = row < p.rows;
            if (sampler) {
                // tile prologue: x_T (injected or Philox slot 0), first H0 image
#pragma unroll
                for (int a = 0; a < 32; ++a) {
                    float v = 0.f;
                    if (valid && a < A) {
                        v = p.xT ? p.xT[(size_t)row * A + a] : philox_normal(p.seed, p.offset, p.row_offset + row, 0, a);
                        if (p.chains_out && p.K == p.T) p.chains_out[((size_t)row * (p.K + 1)) * A + a] = v;
                    }
                    x[a] = v;
                }
                build_h0_row(h0_addr, rloc, x, p.obs + (size_t)(valid ? row : 0) * p.Do, A, p.Do, p.T, p.T - 1, valid);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(x_full);
            } else if (first) {
                if (lane == 0) mbar_arrive(x_full);              // "TMEM is free" for the very first layer
            }
            first = false;
            for (int step = 0; step < nsteps; ++step) {
                const int t_s = p.T - 1 - step;
                const int net = (sampler && t_s < p.K && !p.use_base_policy) ? 1 : 0;
                for (int l = 0; l < p.nlayers; ++l) {
                    const Layer& L = p.L[net][l];
                    const bool final_layer = (l == p.nlayers - 1) && p.final_mode != FINAL_STORE;
                    mbar_wait(acc_full, acc_phase); acc_phase ^= 1;
                    tcgen05_fence_after();
                    if (!final_layer) {
                        // ---------------- generic layer: TMEM -> X (in place)
                        if (store_thread) tma_store_wait_read();       // earlier TMA stores must have finished reading X
                        epi_barrier();
                        const uint32_t* min_row = L.mask_in ? L.mask_in + (size_t)row * (H / 32) : nullptr;
                        uint32_t* mout_row = L.mask_out ? L.mask_out + (size_t)row * (H / 32) : nullptr;
                        uint32_t mo[4] = {0u, 0u, 0u, 0u};
                        uint4 mi4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
                        for (int c = 0; c < L.n / 32; ++c) {
                            uint32_t r[32];
                            tmem_ld32(tmem_base + lane_base + (uint32_t)(c * 32), r);
                            const int n0 = c * 32;
                            float v[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                            if (L.bias) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] += __ldg(L.bias + n0 + j);
                            }
                            if (L.act == 1) {
                                uint32_t bits = 0u;
#pragma unroll
                                for (int j = 0; j < 32; ++j) { bits |= (v[j] > 0.f ? 1u : 0u) << j; v[j] = fmaxf(v[j], 0.f); }
                                mo[c & 3] = bits;
                                if (mout_row && (c & 3) == 3 && valid) *reinterpret_cast<uint4*>(mout_row + (c - 3)) = make_uint4(mo[0], mo[1], mo[2], mo[3]);
                            }
                            if (min_row) {
                                if ((c & 3) == 0) mi4 = valid ? *reinterpret_cast<const uint4*>(min_row + c) : make_uint4(0u, 0u, 0u, 0u);
                                const uint32_t bits = (c & 3) == 0 ? mi4.x : ((c & 3) == 1 ? mi4.y : ((c & 3) == 2 ? mi4.z : mi4.w));
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) ? v[j] : 0.f;
                            }
                            const uint32_t tile_addr = x_addr + (uint32_t)(n0 >> 6) * 16384u + (uint32_t)rloc * 128u;
                            const int cb = (n0 & 63) >> 3;
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                st_shared_v4(tile_addr + (uint32_t)(((cb + q) ^ (rloc & 7)) << 4),
                                             pack_bf16(v[q * 8 + 0], v[q * 8 + 1]), pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                                             pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
                        }
                        tcgen05_fence_before();
                        fence_async_smem();
                        epi_barrier();
                        if (L.store_map >= 0 && store_thread) {
                            for (int kb = 0; kb < L.n / 64; ++kb) tma_store_2d(&maps.m[L.store_map], sX + kb * 16384, kb * 64, tile * FBM);
                            tma_store_commit();
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(x_full);
                    } else {
                        // ---------------- final layer: this thread's row of eps sits in TMEM columns [0, 32)
                        uint32_t r[32];
                        tmem_ld32(tmem_base + lane_base, r);
                        tcgen05_fence_before();
                        float eps[32];
#pragma unroll
                        for (int a = 0; a < 32; ++a) eps[a] = __uint_as_float(r[a]) + ((L.bias && a < A) ? __ldg(L.bias + a) : 0.f);
                        if (p.final_mode == FINAL_EPS) {
                            if (valid) {
#pragma unroll
                                for (int a = 0; a < 32; ++a) if (a < A) p.out[(size_t)row * A + a] = eps[a];
                            }
                        } else if (p.final_mode == FINAL_LOGP) {
                            if (valid) {
                                const float *pv, *nx; int t;
                                if (p.chains) {
                                    const int b = row / p.K, k = row % p.K;
                                    pv = p.chains + ((size_t)b * (p.K + 1) + k) * A; nx = pv + A; t = p.K - 1 - k;
                                } else { pv = p.prev + (size_t)row * A; nx = p.next + (size_t)row * A; t = p.trow[row]; }
#pragma unroll
                                for (int a = 0; a < 32; ++a) if (a < A)
                                    p.out[(size_t)row * A + a] = logprob_elem(pv[a], eps[a], nx[a], t, p.sch, p.T, p.dcv, p.min_lp_std, nullptr, nullptr, nullptr);
                            }
                        } else {   // FINAL_SAMPLE: posterior mean, clipped noise, x update (diffusion_vpg.py:198-206,239-243,301-338)
                            if (valid) {
#pragma unroll
                                for (int a = 0; a < 32; ++a) if (a < A) {
                                    const float nz = p.noise ? p.noise[((size_t)step * p.rows + row) * A + a]
                                                             : philox_normal(p.seed, p.offset, p.row_offset + row, 1 + step, a);
                                    const float xn = ddpm_step_elem(x[a], eps[a], nz, t_s, p.sch, p.T, p.hp, t_s == 0);
                                    x[a] = xn;
                                    if (p.chains_out && t_s <= p.K) p.chains_out[((size_t)row * (p.K + 1) + (p.K - t_s)) * A + a] = xn;
                                    if (t_s == 0) p.actions[(size_t)row * A + a] = xn;
                                }
                            }
                            if (step + 1 < nsteps) {
                                build_h0_row(h0_addr, rloc, x, p.obs + (size_t)(valid ? row : 0) * p.Do, A, p.Do, p.T, t_s - 1, valid);
                                fence_async_smem();
                            }
                        }
                        __syncwarp();
                        // the sampler's last step hands over to the next tile's prologue instead
                        if (lane == 0 && !(sampler && step + 1 == nsteps)) mbar_arrive(x_full);
                    }
                }
            }
        }
        if (store_thread) tma_store_wait_all();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------ host side
// weight operand [K][N] bf16 row-major: box [WK k][64 n]
static int weight_map(CUtensorMap* m, const void* base, uint64_t K, uint64_t N) { return make_map(m, base, K, N, N, WK, 64); }
// row tile source / destination [rows][cols] bf16 row-major: box [128 rows][64 cols]
static int rowtile_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols) { return make_map(m, base, rows, cols, cols, FBM, 64); }

template <int H>
static int launch_chain_t(dppo_handle* h, cudaStream_t s, const Maps& maps, const Params& p, double flops) {
    auto kern = chain_kernel<H>;
    static bool attr_set = false;
    if (!attr_set) { CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem_bytes<H>())); attr_set = true; }
    const int ntiles = (p.rows + FBM - 1) / FBM;
    const int grid = ntiles < h->sm_count ? ntiles : h->sm_count;
    prof_begin(h, s);
    kern<<<grid, FTHREADS, chain_smem_bytes<H>(), s>>>(maps, p);
    prof_end(h, s, flops);
    h->launches++; h->tc_launches++; h->fused_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) DPPO_FAIL(-3, "fused chain launch failed: %s", cudaGetErrorString(e));
    return 0;
}
static int launch_chain(dppo_handle* h, cudaStream_t s, int H, const Maps& maps, const Params& p, double flops) {
    if (H == 512) return launch_chain_t<512>(h, s, maps, p, flops);
    if (H == 256) return launch_chain_t<256>(h, s, maps, p, flops);
    DPPO_FAIL(-7, "fused chain: unsupported hidden width %d", H);
}

}  // namespace fc
