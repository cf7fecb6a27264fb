// Split-precision tcgen05 path (DPPO_PREC_BF16X3): fp32-faithful results on the bf16 tensor pipe.
//
// Every fp32 operand x is carried as bf16 PLANES p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1) (8 mantissa bits
// each) and every product is evaluated as a sum of exact bf16 x bf16 plane products accumulated in fp32 in tensor memory,
// several tcgen05.mma per k-step into ONE accumulator:
//   P = 2 (16-bit operands, backward / weight-gradient GEMMs):  A B ~= A0 B0 + A1 B0 + A0 B1              (dropped: 2^-18)
//   P = 3 (24-bit operands, forward GEMMs):                     A B ~= A0 B0 + A0 B1 + A1 B0 + A1 B1 + A0 B2 + A2 B0   (2^-27)
// Why the forward pass needs P = 3: the loss is only piecewise smooth (ReLU kinks, the +-1 clip of x0, the PPO ratio clip).  A
// forward deviation eps from the reference flips about eps * density units across a kink, and every flipped unit switches a
// whole gradient term on or off: the actor-gradient error against the oracle grows like sqrt(eps) - 2e-3 of the largest
// entry with 16-bit forward operands (measured, tools/flip_probe.py), against the 1e-3 north_star's fp32 mode is held to.
// With 24-bit forward operands the forward pass is as close to the oracle as the FFMA path is; the backward and
// weight-gradient GEMMs are smooth in their operands and keep P = 2 (1e-5 relative).
//
//   ts::split_gemm_kernel<BN, A_MN, B_MN, P>   D[M,N] = sum_seg sum_(i,j) A_i B_j
//     * a shared-memory stage holds the 2 P plane tiles of one 64-deep k-block (TMA, 128-byte swizzle): each byte that
//       enters the SM feeds 1.5 (P = 2) or 2 (P = 3) MMAs instead of the 1 of a K-concatenated formulation
//     * 128 x BN accumulators double-buffered in TMEM; 4 epilogue warps: bias / activation / act' mask / residual add, outputs
//       as hi + lo planes (the next GEMM's operands), fp32, or split-K partials (deterministic fixed-order reduction)
//     * operands K-major or MN-major as stored (activations [rows][features], weights [in][out]): nothing is transposed
//   host programs (same data flow as the per-layer bf16 programs of tc_path.cuh, every activation / gradient as two planes):
//     forward L0..L3 (residual by K-concatenation [a1 | h0] x [W2 ; W0]), backward dv / dh1 / du, the four weight-gradient
//     products per net, bias gradients by column sums, the time-embedding backward from dW0's one-hot rows.
#pragma once
#include "tc_path.cuh"

namespace ts {
using namespace tc;

constexpr int SBM = 128, SBK = 64, STHREADS = 192;   // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue
constexpr int MAXP = 3;

struct Epi {
    int M, N;                                         // valid extents of the output
    const float* bias;                                // [N]
    int act;                                          // 0 none, 1 relu, 2 mish
    const bf16* mask0; const bf16* mask1; int ldmask; int mask_mode;   // 1: *= (mask0 > 0), 2: *= mish'(mask0 + mask1)
    const bf16* add0; const bf16* add1; int ldadd;                     // += add0 + add1
    float* out_f32; int ld_f32; size_t split_stride;  // fp32 row-major (+ split * split_stride)
    bf16* out[MAXP]; int out_planes; int ld_out;      // post-activation planes (out_planes = 0: none)
    bf16* pre[2]; int ld_pre;                         // pre-activation (Mish nets), two planes
};
struct Params {
    int m_blocks, n_blocks, splits;
    int kblocks, kb_per_split;                        // K blocks (of 64) in total / per split
    int ka_blocks;                                    // K blocks taken from A segment 0 (the rest from segment 1)
    Epi epi;
};
struct Maps { CUtensorMap a[2][MAXP], b[MAXP]; };    // [segment][plane], [plane]

// stage = P A-plane tiles + P B-plane tiles of one k-block
template <int BN, int P> __host__ __device__ constexpr int stage_bytes() { return P * (SBM * SBK * 2 + BN * SBK * 2); }
template <int BN, int P> __host__ __device__ constexpr int stages() { return (220 * 1024) / stage_bytes<BN, P>() > 4 ? 4 : (220 * 1024) / stage_bytes<BN, P>(); }
template <int BN, int P> constexpr size_t smem_bytes() { return (size_t)stages<BN, P>() * stage_bytes<BN, P>() + 1024 + 256; }

__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
    hi = __float2bfloat16(v);
    lo = __float2bfloat16(v - __bfloat162float(hi));
}
__device__ __forceinline__ void split_bf16_3(float v, bf16& p0, bf16& p1, bf16& p2) {
    p0 = __float2bfloat16(v);
    const float r1 = v - __bfloat162float(p0);            // exact
    p1 = __float2bfloat16(r1);
    p2 = __float2bfloat16(r1 - __bfloat162float(p1));     // exact remainder, rounded to the third plane
}
// 32 consecutive outputs of one row as NP bf16 planes
template <int NP>
__device__ __forceinline__ void store_planes32(bf16* const (&dst)[MAXP], size_t off, const float (&v)[32], bool vec, int nvalid) {
    if (vec) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            __align__(16) bf16 t[MAXP][8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (NP == 3) split_bf16_3(v[q * 8 + j], t[0][j], t[1][j], t[2][j]); else split_bf16(v[q * 8 + j], t[0][j], t[1][j]);
            }
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) *reinterpret_cast<uint4*>(dst[pl] + off + q * 8) = *reinterpret_cast<const uint4*>(t[pl]);
        }
    } else {
        for (int j = 0; j < nvalid; ++j) {
            bf16 t0, t1, t2 = __float2bfloat16(0.f);
            if (NP == 3) split_bf16_3(v[j], t0, t1, t2); else split_bf16(v[j], t0, t1);
            dst[0][off + j] = t0; dst[1][off + j] = t1; if (NP == 3) dst[2][off + j] = t2;
        }
    }
}

template <int BN, bool A_MN, bool B_MN, int P>
__global__ void __launch_bounds__(STHREADS, 1) split_gemm_kernel(const __grid_constant__ Maps maps, const Params p) {
    constexpr int STG = stages<BN, P>();
    constexpr int A_TILE = SBM * SBK * 2, B_TILE = BN * SBK * 2, STAGE = stage_bytes<BN, P>();
    // P = 3: the five correction products go to a SECOND accumulator and are added to the main one (A0 B0) in fp32 registers by
    // the epilogue.  tcgen05 accumulation truncates: every MMA into an accumulator costs up to one ulp of ITS magnitude, biased
    // (measured: 192 accumulations per element at K = 512 left 5e-6 relative whether the operands carried 16 or 24 bits).  The
    // correction sum is 2^-8 of the main one, so its truncations vanish and the main accumulator sees K/16 of them instead of 6 K/16.
    constexpr int NACC = P == 3 ? 2 : 1;
    constexpr int ACC_COLS = NACC * BN;
    constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512)));
    static_assert(STG >= 2, "a stage must fit twice");
    static_assert(2 * ACC_COLS <= 512, "accumulators exceed tensor memory");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + STG * STAGE);
    uint64_t* full = bars;                 // [STG]  TMA -> MMA
    uint64_t* empty = bars + STG;          // [STG]  MMA -> TMA
    uint64_t* tfull = bars + 2 * STG;      // [2]    MMA -> epilogue
    uint64_t* tempty = tfull + 2;          // [2]    epilogue -> MMA
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = p.m_blocks * p.n_blocks * p.splits;

    if (warp == 0 && lane == 0) {
        for (int pl = 0; pl < P; ++pl) { tma_prefetch_desc(&maps.a[0][pl]); tma_prefetch_desc(&maps.b[pl]); }
        for (int i = 0; i < STG; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ===================== TMA producer: the 2 P plane tiles of one k-block per stage =====================
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int split = tile / (p.m_blocks * p.n_blocks);
            const int rem = tile % (p.m_blocks * p.n_blocks);
            const int m_blk = rem / p.n_blocks, n_blk = rem % p.n_blocks;
            const int kb0 = split * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (elect_one_lane()) {
                    mbar_expect_tx(&full[stage], STAGE);
                    uint8_t* sa = smem + stage * STAGE; uint8_t* sb = sa + P * A_TILE;
                    const int seg = kb < p.ka_blocks ? 0 : 1;
                    const int ka = (seg ? kb - p.ka_blocks : kb) * SBK;
#pragma unroll
                    for (int pl = 0; pl < P; ++pl) {
                        uint8_t* a = sa + pl * A_TILE; uint8_t* b = sb + pl * B_TILE;
                        if (!A_MN) tma_load_2d(a, &maps.a[seg][pl], &full[stage], ka, m_blk * SBM);
                        else {
#pragma unroll
                            for (int j = 0; j < SBM / 64; ++j) tma_load_2d(a + j * (64 * SBK * 2), &maps.a[seg][pl], &full[stage], m_blk * SBM + j * 64, ka);
                        }
                        if (!B_MN) tma_load_2d(b, &maps.b[pl], &full[stage], kb * SBK, n_blk * BN);
                        else {
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j) tma_load_2d(b + j * (64 * SBK * 2), &maps.b[pl], &full[stage], n_blk * BN + j * 64, kb * SBK);
                        }
                    }
                }
                __syncwarp();
                if (++stage == STG) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: all plane products of a k-step into one accumulator =====================
        constexpr uint32_t idesc = make_idesc(SBM, BN, A_MN, B_MN);
        // (A plane, B plane) pairs, small terms first
        constexpr int NPROD = P == 3 ? 6 : 3;
        constexpr int PA[6] = {2, 0, 1, 1, 0, 0}, PB[6] = {0, 2, 1, 0, 1, 0};       // P = 3
        constexpr int QA[3] = {1, 0, 0}, QB[3] = {0, 1, 0};                          // P = 2
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int split = tile / (p.m_blocks * p.n_blocks);
            const int kb0 = split * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[stage], phase);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + stage * STAGE), sb = sa + P * A_TILE;
                if (elect_one_lane()) {
#pragma unroll
                    for (int k = 0; k < SBK / 16; ++k) {
                        // K-major: 16 bf16 = 32 B inside the 128 B swizzle row; SBO = 8 rows * 128 B.
                        // MN-major: 16 k-rows = 2048 B; LBO = next 64-wide MN atom (64*SBK*2 B), SBO = 8 k-rows.
                        uint64_t da[P], db[P];
#pragma unroll
                        for (int pl = 0; pl < P; ++pl) {
                            da[pl] = A_MN ? make_desc(sa + pl * A_TILE + k * 2048, 64 * SBK * 2, 1024) : make_desc(sa + pl * A_TILE + k * 32, 16, 1024);
                            db[pl] = B_MN ? make_desc(sb + pl * B_TILE + k * 2048, 64 * SBK * 2, 1024) : make_desc(sb + pl * B_TILE + k * 32, 16, 1024);
                        }
#pragma unroll
                        for (int q = 0; q < NPROD; ++q) {
                            const int ia = P == 3 ? PA[q] : QA[q], ib = P == 3 ? PB[q] : QB[q];
                            if (NACC == 2) {       // main product (last in the list) -> columns [0, BN), corrections -> [BN, 2 BN)
                                const bool main = q == NPROD - 1;
                                umma_bf16(tmem_d + (main ? 0u : (uint32_t)BN), da[ia], db[ib], idesc, (kb > kb0 || k > 0 || (!main && q > 0)) ? 1u : 0u);
                            } else umma_bf16(tmem_d, da[ia], db[ib], idesc, (kb > kb0 || k > 0 || q > 0) ? 1u : 0u);
                        }
                    }
                    tcgen05_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == STG) { stage = 0; phase ^= 1; }
            }
            if (elect_one_lane()) tcgen05_commit(&tfull[acc]);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================== epilogue warps (TMEM -> registers -> global planes) =====================
        const int quad = warp & 3;
        const Epi& e = p.epi;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int split = tile / (p.m_blocks * p.n_blocks);
            const int rem = tile % (p.m_blocks * p.n_blocks);
            const int m_blk = rem / p.n_blocks, n_blk = rem % p.n_blocks;
            mbar_wait(&tfull[acc], acc_phase);
            tcgen05_fence_after();
            const int m = m_blk * SBM + quad * 32 + lane;
            const bool row_ok = m < e.M;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * ACC_COLS + c * 32), r);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (NACC == 2) {
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * ACC_COLS + BN + c * 32), r);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r[j]);
                }
                const int n0 = n_blk * BN + c * 32;
                if (row_ok && n0 < e.N) {
                    const bool full32 = n0 + 32 <= e.N;
                    const int nvalid = full32 ? 32 : e.N - n0;
                    if (e.bias) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (full32 || j < nvalid) v[j] += __ldg(e.bias + n0 + j);
                    }
                    if (e.pre[0]) {
                        bf16* const pd[MAXP] = {e.pre[0], e.pre[1], nullptr};
                        store_planes32<2>(pd, (size_t)m * e.ld_pre + n0, v, full32 && (e.ld_pre & 7) == 0, nvalid);
                    }
                    if (e.act == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                    } else if (e.act == 2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = mish_f(v[j]);
                    }
                    if (e.mask0) {
                        const bf16* sh = e.mask0 + (size_t)m * e.ldmask + n0;
                        const bf16* sl = e.mask1 ? e.mask1 + (size_t)m * e.ldmask + n0 : nullptr;
                        if (full32 && (e.ldmask & 7) == 0) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 uh = *reinterpret_cast<const uint4*>(sh + q * 8);
                                const bf16* hb = reinterpret_cast<const bf16*>(&uh);
                                if (e.mask_mode == 1) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) v[q * 8 + j] = __bfloat162float(hb[j]) > 0.f ? v[q * 8 + j] : 0.f;
                                } else {
                                    const uint4 ul = *reinterpret_cast<const uint4*>(sl + q * 8);
                                    const bf16* lb = reinterpret_cast<const bf16*>(&ul);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) v[q * 8 + j] *= mish_grad_f(__bfloat162float(hb[j]) + __bfloat162float(lb[j]));
                                }
                            }
                        } else {
                            for (int j = 0; j < nvalid; ++j) {
                                const float mh = __bfloat162float(sh[j]);
                                v[j] *= (e.mask_mode == 1) ? (mh > 0.f ? 1.f : 0.f) : mish_grad_f(mh + __bfloat162float(sl[j]));
                            }
                        }
                    }
                    if (e.add0) {
                        const bf16* sh = e.add0 + (size_t)m * e.ldadd + n0;
                        const bf16* sl = e.add1 + (size_t)m * e.ldadd + n0;
                        if (full32 && (e.ldadd & 7) == 0) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 uh = *reinterpret_cast<const uint4*>(sh + q * 8), ul = *reinterpret_cast<const uint4*>(sl + q * 8);
                                const bf16* hb = reinterpret_cast<const bf16*>(&uh); const bf16* lb = reinterpret_cast<const bf16*>(&ul);
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[q * 8 + j] += __bfloat162float(hb[j]) + __bfloat162float(lb[j]);
                            }
                        } else {
                            for (int j = 0; j < nvalid; ++j) v[j] += __bfloat162float(sh[j]) + __bfloat162float(sl[j]);
                        }
                    }
                    if (e.out_planes == 3) store_planes32<3>(e.out, (size_t)m * e.ld_out + n0, v, full32 && (e.ld_out & 7) == 0, nvalid);
                    else if (e.out_planes == 2) store_planes32<2>(e.out, (size_t)m * e.ld_out + n0, v, full32 && (e.ld_out & 7) == 0, nvalid);
                    if (e.out_f32) {
                        float* dst = e.out_f32 + (size_t)split * e.split_stride + (size_t)m * e.ld_f32 + n0;
                        if (full32 && (e.ld_f32 & 3) == 0) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                *reinterpret_cast<float4*>(dst + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                        } else {
                            for (int j = 0; j < nvalid; ++j) dst[j] = v[j];
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ host side
// One split operand: `planes` planes with the same shape.  K-major: memory is [mn][k]; MN-major: memory is [k][mn].
struct Operand { const bf16* p[MAXP]; bool mn_major; int64_t mn, k, ld; };
struct Gemm {
    Operand A, A2, B;      // A2.p[0] == nullptr: no K concatenation; B spans the concatenated K
    int planes;            // 2 or 3 (every operand must carry that many)
    int M, N, splits;
    double alg_flops;      // algorithmic flops (un-padded dims, ONE product per MAC) for the live roofline
    Epi epi;
};

template <int BN, bool A_MN, bool B_MN, int P>
static int launch_t(dppo_handle* h, cudaStream_t s, const Gemm& g) {
    Maps mp;
    const Operand& A = g.A; const Operand& B = g.B;
    auto amap = [&](CUtensorMap* m, const bf16* ptr, const Operand& o) -> int {
        return A_MN ? make_map(m, ptr, o.k, o.mn, o.ld, SBK, 64) : make_map(m, ptr, o.mn, o.k, o.ld, SBM, SBK);
    };
    for (int pl = 0; pl < MAXP; ++pl) {
        const int src = pl < P ? pl : 0;
        DPPO_TRY(amap(&mp.a[0][pl], A.p[src], A));
        if (g.A2.p[0]) DPPO_TRY(amap(&mp.a[1][pl], g.A2.p[src], g.A2)); else mp.a[1][pl] = mp.a[0][pl];
        if (!B_MN) DPPO_TRY(make_map(&mp.b[pl], B.p[src], B.mn, B.k, B.ld, BN, SBK));
        else DPPO_TRY(make_map(&mp.b[pl], B.p[src], B.k, B.mn, B.ld, SBK, 64));
    }
    Params p;
    p.m_blocks = (g.M + SBM - 1) / SBM; p.n_blocks = (g.N + BN - 1) / BN;
    const int ka = (int)((A.k + SBK - 1) / SBK), ka2 = g.A2.p[0] ? (int)((g.A2.k + SBK - 1) / SBK) : 0;
    p.kblocks = ka + ka2; p.ka_blocks = ka;
    int splits = g.splits < 1 ? 1 : g.splits;
    if (splits > p.kblocks) splits = p.kblocks;
    p.kb_per_split = (p.kblocks + splits - 1) / splits;
    p.splits = (p.kblocks + p.kb_per_split - 1) / p.kb_per_split;
    p.epi = g.epi;
    auto kern = split_gemm_kernel<BN, A_MN, B_MN, P>;
    static bool attr_set_dev[64] = {};      // function attributes are per device
    bool& attr_set = attr_set_dev[h->device & 63];
    if (!attr_set) { CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<BN, P>())); attr_set = true; }
    const int tiles = p.m_blocks * p.n_blocks * p.splits;
    const int grid = tiles < h->sm_count ? tiles : h->sm_count;
    prof_begin(h, s);
    kern<<<grid, STHREADS, smem_bytes<BN, P>(), s>>>(mp, p);
    prof_end(h, s, g.alg_flops > 0 ? g.alg_flops : 2.0 * (double)g.M * (double)g.N * (double)(A.k + (g.A2.p[0] ? g.A2.k : 0)), 1);
    h->launches++; h->tc_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) DPPO_FAIL(-3, "split gemm launch failed: %s", cudaGetErrorString(e));
    return p.splits;
}
// returns the number of split-K partials written (>= 1) or a negative error.  Only the combinations the programs below use
// are instantiated: forward (K-major A, MN-major B; P = 2 or 3), output layer / backward (K-major both), weight gradients
// (MN-major both, P = 2).
static int launch(dppo_handle* h, cudaStream_t s, const Gemm& g) {
    const bool a = g.A.mn_major, b = g.B.mn_major;
    if (g.planes == 3) {
        if (!a && b) return launch_t<128, false, true, 3>(h, s, g);
        if (!a && !b && g.N <= 32) return launch_t<32, false, false, 3>(h, s, g);
        DPPO_FAIL(-7, "split gemm: this 3-plane operand layout is not instantiated");
    }
    if (!a && b) return launch_t<256, false, true, 2>(h, s, g);
    if (!a && !b) return g.N <= 32 ? launch_t<32, false, false, 2>(h, s, g) : launch_t<256, false, false, 2>(h, s, g);
    if (a && b) return g.N <= 64 ? launch_t<64, true, true, 2>(h, s, g) : launch_t<256, true, true, 2>(h, s, g);
    DPPO_FAIL(-7, "split gemm: operand layout (A MN-major, B K-major) is not instantiated");
}
}  // namespace ts

// =====================================================================================================================
// state: plane copies of the weights per net (3 planes each; the backward GEMMs read the first two)
struct TsW { bf16* p[ts::MAXP]; };
struct TsNetW {
    TsW w2w0, w1, w3t, w3p;
    float* bias2;                         // critic: b2 + b_in (the residual's input-layer bias rides with block.l2's)
    int H;
};
struct TsState { TsNetW net[4]; int KP0; };
struct SplitT { bf16* p[ts::MAXP]; };
static inline SplitT split_null() { SplitT t; t.p[0] = t.p[1] = t.p[2] = nullptr; return t; }

__device__ __forceinline__ void ts_put(const TsW& W, size_t i, float v) {
    bf16 a, b, c; ts::split_bf16_3(v, a, b, c);
    W.p[0][i] = a; W.p[1][i] = b; W.p[2][i] = c;
}
// same operand layouts as tc_pack_actor_body (w2w0 = [W2 ; W0 rows in h0 order], w3t = W3^T padded to 64 rows, w3p = W3 padded to
// 128 columns), each as three planes; no transposed copies (the backward GEMMs read W1 / W2 K-major as stored)
__global__ void ts_pack_actor_kernel(const float* __restrict__ w, ActorOff o, int A, int td, int Do, int T, int H, int KP0,
                                     const float* __restrict__ bt, const TsNetW W) {
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = i0; i < (size_t)H * H; i += stride) { ts_put(W.w2w0, i, w[o.w2 + i]); ts_put(W.w1, i, w[o.w1 + i]); }
    for (size_t i = i0; i < (size_t)KP0 * H; i += stride) {
        const int k = (int)(i / H), c = (int)(i % H);
        float v = 0.f;
        if (k < A) v = w[o.win + (size_t)k * H + c];
        else if (k < A + Do) v = w[o.win + (size_t)(k + td) * H + c];
        else if (k < A + Do + T) v = bt[(size_t)(k - A - Do) * H + c];
        ts_put(W.w2w0, (size_t)H * H + i, v);
    }
    for (size_t i = i0; i < (size_t)64 * H; i += stride) {
        const int a = (int)(i / H), k = (int)(i % H);
        ts_put(W.w3t, i, a < A ? w[o.w3 + (size_t)k * A + a] : 0.f);
    }
    for (size_t i = i0; i < (size_t)H * 128; i += stride) {
        const int k = (int)(i / 128), a = (int)(i % 128);
        ts_put(W.w3p, i, a < A ? w[o.w3 + (size_t)k * A + a] : 0.f);
    }
}
__global__ void ts_pack_critic_kernel(const float* __restrict__ w, CriticOff o, int A, int Do, int Hc, int KP0, const TsNetW W) {
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = i0; i < (size_t)Hc * Hc; i += stride) { ts_put(W.w2w0, i, w[o.w2 + i]); ts_put(W.w1, i, w[o.w1 + i]); }
    for (size_t i = i0; i < (size_t)KP0 * Hc; i += stride) {
        const int k = (int)(i / Hc), c = (int)(i % Hc);
        ts_put(W.w2w0, (size_t)Hc * Hc + i, (k >= A && k < A + Do) ? w[o.win + (size_t)(k - A) * Hc + c] : 0.f);
    }
    for (size_t i = i0; i < (size_t)64 * Hc; i += stride) { const int a = (int)(i / Hc), k = (int)(i % Hc); ts_put(W.w3t, i, a == 0 ? w[o.w3 + k] : 0.f); }
    for (size_t i = i0; i < (size_t)Hc * 128; i += stride) { const int k = (int)(i / 128), a = (int)(i % 128); ts_put(W.w3p, i, a == 0 ? w[o.w3 + k] : 0.f); }
    for (size_t i = i0; i < (size_t)Hc; i += stride) W.bias2[i] = w[o.b2 + i] + w[o.bin + i];
}
// h0[r] = [x[r] | obs[r / obs_div] | onehot(t_r) | 1 | 0..] as three planes (tc_pack_h0_kernel's layout); one thread per 8 columns
__global__ void ts_pack_h0_kernel(const float* __restrict__ x, const float* __restrict__ obs, const int* __restrict__ trow, int tconst,
                                  int N, int A, int Do, int T, int KP0, int obs_div, const SplitT h0, int chainK) {
    const int g8 = KP0 / 8;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * g8) return;
    const int r = (int)(i / g8), k0 = (int)(i % g8) * 8;
    const int t = trow ? (tconst < 0 ? -tconst - 1 - trow[r] : trow[r]) : tconst;
    const size_t xrow = chainK > 0 ? (size_t)(r / chainK) * (chainK + 1) + (r % chainK) : (size_t)r;
    const size_t orow = (size_t)(r / obs_div);
    __align__(16) bf16 o3[3][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = k0 + j;
        float v = 0.f;
        if (k < A) v = x ? x[xrow * A + k] : 0.f;
        else if (k < A + Do) v = obs[orow * Do + (k - A)];
        else if (k < A + Do + T) v = (k - A - Do == t) ? 1.f : 0.f;
        else if (k == A + Do + T) v = 1.f;
        ts::split_bf16_3(v, o3[0][j], o3[1][j], o3[2][j]);
    }
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) *reinterpret_cast<uint4*>(h0.p[pl] + (size_t)r * KP0 + k0) = *reinterpret_cast<const uint4*>(o3[pl]);
}
// dst[r][0:64] = [src[r][0:ncols] | 0..] as two planes
__global__ void ts_pad64_kernel(const float* __restrict__ src, int N, int ncols, bf16* __restrict__ hi, bf16* __restrict__ lo) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * 8) return;
    const int r = (int)(i >> 3), k0 = (int)(i & 7) * 8;
    __align__(16) bf16 oh[8]; __align__(16) bf16 ol[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ts::split_bf16(k0 + j < ncols ? src[(size_t)r * ncols + k0 + j] : 0.f, oh[j], ol[j]);
    *reinterpret_cast<uint4*>(hi + (size_t)r * 64 + k0) = *reinterpret_cast<const uint4*>(oh);
    *reinterpret_cast<uint4*>(lo + (size_t)r * 64 + k0) = *reinterpret_cast<const uint4*>(ol);
}
// column sums of a two-plane matrix [N][ncols] (ncols even, ncols/2 <= 256): part[blk][ncols]
__global__ void __launch_bounds__(256) ts_colsum_kernel(const bf16* __restrict__ Dh, const bf16* __restrict__ Dl, int N, int ncols, int rows_per_block,
                                                        float* __restrict__ part) {
    __shared__ float2 red[256];
    const int tpr = ncols / 2, groups = 256 / tpr;
    const int cg = threadIdx.x % tpr, rg = threadIdx.x / tpr;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
    float2 acc = make_float2(0.f, 0.f);
    if (rg < groups) {
        for (int r = r0 + rg; r < r1; r += groups) {
            const float2 fh = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Dh + (size_t)r * ncols + 2 * cg));
            const float2 fl = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Dl + (size_t)r * ncols + 2 * cg));
            acc.x += fh.x + fl.x; acc.y += fh.y + fl.y;
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < tpr) {
        float2 sacc = red[threadIdx.x];
        for (int g = 1; g < groups; ++g) { const float2 o = red[g * tpr + threadIdx.x]; sacc.x += o.x; sacc.y += o.y; }
        part[(size_t)blockIdx.x * ncols + 2 * threadIdx.x] = sacc.x;
        part[(size_t)blockIdx.x * ncols + 2 * threadIdx.x + 1] = sacc.y;
    }
}

static const int TS_MIN_ROWS = 2048;
static bool ts_shapes_ok(const dppo_handle* h) { return h->ts && h->ts->net[0].w1.p[0] != nullptr; }
static bool ts_eligible(const dppo_handle* h, int rows) {
    return h->cfg.precision == DPPO_PREC_BF16X3 && ts_shapes_ok(h) && rows >= TS_MIN_ROWS;
}
static int ts_alloc_w(TsW& w, size_t n) {
    CUDA_TRY(cudaMalloc(&w.p[0], ts::MAXP * n * sizeof(bf16)));
    for (int pl = 1; pl < ts::MAXP; ++pl) w.p[pl] = w.p[0] + pl * n;
    return 0;
}
static int ts_init(dppo_handle* h) {
    if (h->cfg.precision != DPPO_PREC_BF16X3) return 0;
    const Geom& g = h->g;
    TsState* st = new TsState();
    memset(st, 0, sizeof(*st));
    h->ts = st;
    st->KP0 = round_up(g.A + g.Do + g.T + 1, 64);
    if ((g.H % 64) || (g.Hc % 64) || g.A > 32) return 0;     // shapes this path does not cover: stays on FFMA
    for (int net = 0; net < 4; ++net) {
        TsNetW& w = st->net[net];
        const size_t H = net == DPPO_NET_CRITIC ? g.Hc : g.H;
        w.H = (int)H;
        DPPO_TRY(ts_alloc_w(w.w2w0, (H + st->KP0) * H)); DPPO_TRY(ts_alloc_w(w.w1, H * H));
        DPPO_TRY(ts_alloc_w(w.w3t, 64 * H)); DPPO_TRY(ts_alloc_w(w.w3p, H * 128));
        CUDA_TRY(cudaMalloc(&w.bias2, H * sizeof(float)));
    }
    return 0;
}
static void ts_destroy(dppo_handle* h) {
    if (!h->ts) return;
    for (int net = 0; net < 4; ++net) { TsNetW& w = h->ts->net[net]; cudaFree(w.w2w0.p[0]); cudaFree(w.w1.p[0]); cudaFree(w.w3t.p[0]); cudaFree(w.w3p.p[0]); cudaFree(w.bias2); }
    delete h->ts; h->ts = nullptr;
}
// rebuild the plane copies of one net (after set_weights / an optimizer step; the actor's bt table must be current)
static int ts_refresh_net(dppo_handle* h, int net, cudaStream_t s) {
    if (h->cfg.precision != DPPO_PREC_BF16X3 || !ts_shapes_ok(h)) return 0;
    const Geom& g = h->g; const TsNetW& w = h->ts->net[net];
    if (net == DPPO_NET_CRITIC) ts_pack_critic_kernel<<<128, 256, 0, s>>>(h->net_w[net], g.co, g.A, g.Do, g.Hc, h->ts->KP0, w);
    else ts_pack_actor_kernel<<<256, 256, 0, s>>>(h->net_w[net], g.ao, g.A, g.td, g.Do, g.T, g.H, h->ts->KP0, h->ad[net].bt, w);
    TC_KCHECK(h);
    return 0;
}

// ------------------------------------------------------------------ one residual MLP on N rows, every tensor as planes
struct TsMlp {
    const TsNetW* W; int H, NO, act1, KP0, net, din;
    int fp;                                // planes of the FORWARD GEMMs: 3 when the net (or what follows it) has kinks, else 2
    const float *b0, *b1, *b2, *b3;
    SplitT h0, a0, a1, v, pre0, pre1, dv, dh1, du;
    float* out;                            // [N][NO] fp32
};
static void ts_mlp_clear(TsMlp& m) {
    memset(&m, 0, sizeof(m));
}
static void ts_actor_mlp(const dppo_handle* h, int net, TsMlp& m) {
    const Geom& g = h->g; const float* w = h->net_w[net];
    ts_mlp_clear(m);
    m.W = &h->ts->net[net]; m.H = g.H; m.NO = g.A; m.act1 = h->cfg.actor_act + 1; m.KP0 = h->ts->KP0; m.net = net; m.din = g.Din;
    m.fp = 3;                                                                        // ReLU kinks and / or the +-1 clip of x0 behind eps
    m.b0 = nullptr; m.b1 = w + g.ao.b1; m.b2 = w + g.ao.b2; m.b3 = w + g.ao.b3;      // b_in rides in W0's one-hot rows (bt table)
}
static void ts_critic_mlp(const dppo_handle* h, TsMlp& m) {
    const Geom& g = h->g; const float* w = h->net_w[DPPO_NET_CRITIC];
    ts_mlp_clear(m);
    m.W = &h->ts->net[DPPO_NET_CRITIC]; m.H = g.Hc; m.NO = 1; m.act1 = h->cfg.critic_act + 1; m.KP0 = h->ts->KP0; m.net = DPPO_NET_CRITIC; m.din = g.Do;
    m.fp = h->cfg.critic_act == DPPO_ACT_RELU ? 3 : 2;                               // Mish is smooth
    m.b0 = w + g.co.bin; m.b1 = w + g.co.b1; m.b2 = m.W->bias2; m.b3 = w + g.co.b3;
}
static size_t ts_mlp_ws_bytes(int N, int H, int KP0, bool mish, bool bwd) {
    return 3 * (ws_bytes((size_t)N * KP0, 2) + 3 * ws_bytes((size_t)N * H, 2)) + 2 * ws_bytes((size_t)N * H, 2) * (size_t)((mish ? 2 : 0) + (bwd ? 3 : 0));
}
static SplitT ts_take(dppo_handle* h, size_t n, int planes) {
    SplitT t = split_null();
    for (int pl = 0; pl < planes; ++pl) t.p[pl] = ws_take<bf16>(h, n);
    return t;
}
static void ts_mlp_take(dppo_handle* h, int N, TsMlp& m, bool bwd) {
    const size_t n = (size_t)N * m.H;
    m.h0 = ts_take(h, (size_t)N * m.KP0, 3);
    m.a0 = ts_take(h, n, 3); m.a1 = ts_take(h, n, 3); m.v = ts_take(h, n, 3);
    if (m.act1 == 2) { m.pre0 = ts_take(h, n, 2); m.pre1 = ts_take(h, n, 2); }
    if (bwd) { m.dv = ts_take(h, n, 2); m.dh1 = ts_take(h, n, 2); m.du = ts_take(h, n, 2); }
}
static ts::Operand ts_op(const bf16* const* p, size_t off, bool mn_major, int64_t mn, int64_t k, int64_t ld) {
    ts::Operand o; o.mn_major = mn_major; o.mn = mn; o.k = k; o.ld = ld;
    for (int pl = 0; pl < ts::MAXP; ++pl) o.p[pl] = p[pl] ? p[pl] + off : nullptr;
    return o;
}
static ts::Operand tsK(const SplitT& t, int64_t mn, int64_t k, int64_t ld) { return ts_op(t.p, 0, false, mn, k, ld); }
static ts::Operand tsMN(const SplitT& t, int64_t mn, int64_t k, int64_t ld) { return ts_op(t.p, 0, true, mn, k, ld); }
static ts::Operand tswK(const TsW& w, size_t off, int64_t mn, int64_t k, int64_t ld) { return ts_op(w.p, off, false, mn, k, ld); }
static ts::Operand tswMN(const TsW& w, size_t off, int64_t mn, int64_t k, int64_t ld) { return ts_op(w.p, off, true, mn, k, ld); }
static ts::Gemm ts_gemm_of(ts::Operand A, ts::Operand B, int M, int N, int planes) {
    ts::Gemm g; memset(&g, 0, sizeof(g));
    g.A = A; g.B = B; g.M = M; g.N = N; g.splits = 1; g.planes = planes; g.epi.M = M; g.epi.N = N;
    return g;
}
static void ts_epi_out(ts::Gemm& g, const SplitT& t, int planes, int ld) {
    for (int pl = 0; pl < ts::MAXP; ++pl) g.epi.out[pl] = pl < planes ? t.p[pl] : nullptr;
    g.epi.out_planes = planes; g.epi.ld_out = ld;
}
static int ts_run(dppo_handle* h, cudaStream_t s, const ts::Gemm& g) { const int r = ts::launch(h, s, g); return r < 0 ? r : 0; }

static int ts_mlp_forward(dppo_handle* h, cudaStream_t s, const TsMlp& m, int N) {
    const int H = m.H, KP0 = m.KP0, P = m.fp; const TsNetW& W = *m.W;
    const size_t w0 = (size_t)H * H;
    // L0: a0 = act(h0 W0 (+ b0))
    ts::Gemm g = ts_gemm_of(tsK(m.h0, N, KP0, KP0), tswMN(W.w2w0, w0, H, KP0, H), N, H, P);
    g.epi.bias = m.b0; g.epi.act = m.act1; ts_epi_out(g, m.a0, P, H);
    g.epi.pre[0] = m.pre0.p[0]; g.epi.pre[1] = m.pre0.p[1]; g.epi.ld_pre = H;
    g.alg_flops = 2.0 * N * (double)m.din * H;
    DPPO_TRY(ts_run(h, s, g));
    // L1: a1 = act(a0 W1 + b1)
    g = ts_gemm_of(tsK(m.a0, N, H, H), tswMN(W.w1, 0, H, H, H), N, H, P);
    g.epi.bias = m.b1; g.epi.act = m.act1; ts_epi_out(g, m.a1, P, H);
    g.epi.pre[0] = m.pre1.p[0]; g.epi.pre[1] = m.pre1.p[1]; g.epi.ld_pre = H;
    DPPO_TRY(ts_run(h, s, g));
    // L2 + residual: v = [a1 | h0] [W2 ; W0] + b2 (+ b0)
    g = ts_gemm_of(tsK(m.a1, N, H, H), tswMN(W.w2w0, 0, H, H + KP0, H), N, H, P);
    g.A2 = tsK(m.h0, N, KP0, KP0);
    g.epi.bias = m.b2; ts_epi_out(g, m.v, P, H);
    g.alg_flops = 2.0 * N * (double)H * H;                                       // the re-accumulated residual is not algorithmic work
    DPPO_TRY(ts_run(h, s, g));
    // L3: out = v W3 + b3
    g = ts_gemm_of(tsK(m.v, N, H, H), tswK(W.w3t, 0, 32, H, H), N, m.NO, P);
    g.epi.bias = m.b3; g.epi.out_f32 = m.out; g.epi.ld_f32 = m.NO;
    DPPO_TRY(ts_run(h, s, g));
    return 0;
}
// dW[out_rows][out_cols] = X^T D  (X [rows][M], D [rows][Nd], two planes each), split-K over the rows with a FIXED-order reduction
static int ts_dw(dppo_handle* h, cudaStream_t s, const SplitT& X, int M, const SplitT& D, int Nd, int rows, float* part, float* out, int out_rows, int out_cols,
                 int ld_out, int alg_rows) {
    ts::Gemm g = ts_gemm_of(tsMN(X, M, rows, M), tsMN(D, Nd, rows, Nd), M, Nd, 2);
    g.splits = tc_splits_for(h, M, Nd, rows);
    g.alg_flops = 2.0 * (double)rows * (double)alg_rows * (double)out_cols;
    g.epi.out_f32 = part; g.epi.ld_f32 = Nd; g.epi.split_stride = (size_t)M * Nd;
    const int S = ts::launch(h, s, g);
    if (S < 0) return S;
    tc_reduce2d_kernel<<<tc_nblk((size_t)out_rows * out_cols, 256), 256, 0, s>>>(part, S, (size_t)M * Nd, out_rows, out_cols, Nd, out, ld_out);
    TC_KCHECK(h);
    return 0;
}
static int ts_colsum(dppo_handle* h, cudaStream_t s, const SplitT& D, int N, int ncols, float* part, float* out) {
    int nb = 2 * h->sm_count; if (nb > (N + 63) / 64) nb = (N + 63) / 64; if (nb < 1) nb = 1;
    int rpb = (N + nb - 1) / nb; nb = (N + rpb - 1) / rpb;
    if (ncols / 2 > 256 || (ncols & 1)) DPPO_FAIL(-7, "ts_colsum: %d columns unsupported", ncols);
    ts_colsum_kernel<<<nb, 256, 0, s>>>(D.p[0], D.p[1], N, ncols, rpb, part); TC_KCHECK(h);
    reduce_partials_kernel<<<tc_nblk(ncols, 256), 256, 0, s>>>(part, nb, (size_t)ncols, (size_t)ncols, out, 1.f); TC_KCHECK(h);
    return 0;
}
// backward of the residual MLP from dout [N][64] (two planes, zero padded).  Writes gradients of W1,b1,W2,b2,W3 into gnet at the
// given offsets and dW0 (in h0 row order) into dw0 [KP0][H].
static int ts_mlp_backward(dppo_handle* h, cudaStream_t s, const TsMlp& m, const SplitT& dout, int N, float* part,
                           float* gnet, size_t ow1, size_t ob1, size_t ow2, size_t ob2, size_t ow3, float* dw0) {
    const int H = m.H, KP0 = m.KP0; const TsNetW& W = *m.W;
    // dv = dout W3^T
    ts::Gemm g = ts_gemm_of(tsK(dout, N, 64, 64), tswK(W.w3p, 0, H, 64, 128), N, H, 2);
    ts_epi_out(g, m.dv, 2, H);
    g.alg_flops = 2.0 * N * (double)m.NO * H;
    DPPO_TRY(ts_run(h, s, g));
    // dh1 = (dv W2^T) * act'(h1)
    g = ts_gemm_of(tsK(m.dv, N, H, H), tswK(W.w2w0, 0, H, H, H), N, H, 2);
    if (m.act1 == 2) { g.epi.mask0 = m.pre1.p[0]; g.epi.mask1 = m.pre1.p[1]; } else { g.epi.mask0 = m.a1.p[0]; g.epi.mask1 = nullptr; }
    g.epi.ldmask = H; g.epi.mask_mode = m.act1;
    ts_epi_out(g, m.dh1, 2, H);
    DPPO_TRY(ts_run(h, s, g));
    // du = (dh1 W1^T) * act'(u) + dv
    g = ts_gemm_of(tsK(m.dh1, N, H, H), tswK(W.w1, 0, H, H, H), N, H, 2);
    if (m.act1 == 2) { g.epi.mask0 = m.pre0.p[0]; g.epi.mask1 = m.pre0.p[1]; } else { g.epi.mask0 = m.a0.p[0]; g.epi.mask1 = nullptr; }
    g.epi.ldmask = H; g.epi.mask_mode = m.act1;
    g.epi.add0 = m.dv.p[0]; g.epi.add1 = m.dv.p[1]; g.epi.ldadd = H;
    ts_epi_out(g, m.du, 2, H);
    DPPO_TRY(ts_run(h, s, g));
    // weight gradients
    DPPO_TRY(ts_dw(h, s, m.v, H, dout, 64, N, part, gnet + ow3, H, m.NO, m.NO, H));
    DPPO_TRY(ts_dw(h, s, m.a1, H, m.dv, H, N, part, gnet + ow2, H, H, H, H));
    DPPO_TRY(ts_dw(h, s, m.a0, H, m.dh1, H, N, part, gnet + ow1, H, H, H, H));
    DPPO_TRY(ts_dw(h, s, m.h0, KP0, m.du, H, N, part, dw0, KP0, H, H, m.din));
    // bias gradients of block.l1 / block.l2 (the input-layer bias comes out of dw0's constant rows)
    DPPO_TRY(ts_colsum(h, s, m.dv, N, H, part, gnet + ob2));
    DPPO_TRY(ts_colsum(h, s, m.dh1, N, H, part, gnet + ob1));
    return 0;
}

// ------------------------------------------------------------------ forward-only programs
static size_t ts_actor_forward_ws(const dppo_handle* h, int N) { return ts_mlp_ws_bytes(N, h->g.H, h->ts->KP0, h->cfg.actor_act == DPPO_ACT_MISH, false); }
// eps[N][A] = actor(x, t, obs); appends to the workspace (the caller may hold pointers below ws.used)
static int ts_actor_forward(dppo_handle* h, cudaStream_t s, int net, const float* x, const float* obs, int obs_div, int N,
                            const int* trow, int tconst, float* eps, int chainK = 0) {
    const Geom& g = h->g; const int KP0 = h->ts->KP0;
    const size_t need = h->ws.used + ts_actor_forward_ws(h, N);
    if (need > h->ws.cap) DPPO_FAIL(-7, "ts_actor_forward: workspace too small (%zu > %zu)", need, h->ws.cap);
    const size_t mark = h->ws.used;
    TsMlp m; ts_actor_mlp(h, net, m);
    ts_mlp_take(h, N, m, false);
    m.out = eps;
    ts_pack_h0_kernel<<<tc_nblk((size_t)N * (KP0 / 8), 256), 256, 0, s>>>(x, obs, trow, tconst, N, g.A, g.Do, g.T, KP0, obs_div, m.h0, chainK);
    TC_KCHECK(h);
    const int r = ts_mlp_forward(h, s, m, N);
    h->ws.used = mark;
    return r;
}
static int ts_value(dppo_handle* h, cudaStream_t s, const float* obs, int N, float* v) {
    const Geom& g = h->g; const int KP0 = h->ts->KP0;
    DPPO_TRY(ws_reserve(h, ts_mlp_ws_bytes(N, g.Hc, KP0, h->cfg.critic_act == DPPO_ACT_MISH, false), s));
    TsMlp m; ts_critic_mlp(h, m);
    ts_mlp_take(h, N, m, false);
    m.out = v;
    ts_pack_h0_kernel<<<tc_nblk((size_t)N * (KP0 / 8), 256), 256, 0, s>>>(nullptr, obs, nullptr, -1, N, g.A, g.Do, g.T, KP0, 1, m.h0, 0);
    TC_KCHECK(h);
    return ts_mlp_forward(h, s, m, N);
}

// ------------------------------------------------------------------ gradients
// actor backward from deps [N][A] fp32: fills gnet[0 : nA]
static int ts_actor_grads(dppo_handle* h, cudaStream_t s, int net, const TsMlp& m, const float* deps, const SplitT& depsb, int N, float* part, float* dw0, float* gnet) {
    const Geom& g = h->g; const float* w = h->net_w[net]; const ActorDerived& d = h->ad[net];
    DPPO_TRY(ts_mlp_backward(h, s, m, depsb, N, part, gnet, g.ao.w1, g.ao.b1, g.ao.w2, g.ao.b2, g.ao.w3, dw0));
    DPPO_TRY(colsum(h, s, deps, g.A, N, g.A, nullptr, 1, part, gnet + g.ao.b3));
    // dw0 rows [A+Do, A+Do+T) are the per-t column sums of du: the gradient of the bt table
    const size_t sm = (size_t)(g.T * g.td * 3) * sizeof(float);
    time_backward_kernel<<<1, 512, sm, s>>>(w, g.ao, g.A, g.td, g.H, g.T, dw0 + (size_t)(g.A + g.Do) * g.H, d.sinemb, d.thpre, d.temb, gnet);
    TC_KCHECK(h);
    unpack_dw0_kernel<<<tc_nblk((size_t)(g.A + g.Do) * g.H, 256), 256, 0, s>>>(dw0, g.A, g.td, g.Do, g.H, gnet + g.ao.win);
    TC_KCHECK(h);
    return 0;
}
static int ts_critic_grads(dppo_handle* h, cudaStream_t s, const TsMlp& m, const float* dval, const SplitT& dvalb, int N, float* part, float* dw0, float* gnet) {
    const Geom& g = h->g;
    DPPO_TRY(ts_mlp_backward(h, s, m, dvalb, N, part, gnet, g.co.w1, g.co.b1, g.co.w2, g.co.b2, g.co.w3, dw0));
    DPPO_TRY(colsum(h, s, dval, 1, N, 1, nullptr, 1, part, gnet + g.co.b3));
    unpack_dw0_kernel<<<tc_nblk((size_t)g.Do * g.Hc, 256), 256, 0, s>>>(dw0 + (size_t)g.A * g.Hc, 0, 0, g.Do, g.Hc, gnet + g.co.win);
    TC_KCHECK(h);
    // the ones column of h0 collects the input-layer bias gradient
    CUDA_TRY(cudaMemcpyAsync(gnet + g.co.bin, dw0 + (size_t)(g.A + g.Do + g.T) * g.Hc, g.Hc * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return 0;
}

// PPODiffusion.c_loss + tape.gradient (diffusion_ppo.py:32-132, train_ppo_diffusion_agent.py:340-346).  Leaves
// [actor_ft grads | critic grads | 8 metrics] in h->grads.  The loss itself is the fp32 parity path's kernel.
static int ts_ppo_step(dppo_handle* h, cudaStream_t s, const float* obs, const float* prev, const float* nxt, const int32_t* inds,
                       const float* returns, const float* oldvalues, const float* advantages, const float* oldlogp,
                       int N, int64_t N_global, float adv_mean, float adv_std) {
    const Geom& g = h->g; const int KP0 = h->ts->KP0;
    const size_t nA = g.ao.n, nC = g.co.n;
    float* gr = h->grads;
    const bool amish = h->cfg.actor_act == DPPO_ACT_MISH, cmish = h->cfg.critic_act == DPPO_ACT_MISH;
    const int nlb = tc_nblk(N, 128);
    const size_t pf = tc_part_floats(h, g.H);
    const size_t need = ts_mlp_ws_bytes(N, g.H, KP0, amish, true) + ts_mlp_ws_bytes(N, g.Hc, KP0, cmish, true)
                      + 2 * ws_bytes((size_t)N * g.A, 4) + 2 * ws_bytes(N, 4) + 4 * ws_bytes((size_t)N * 64, 2)
                      + ws_bytes(pf, 4) + ws_bytes((size_t)KP0 * g.H, 4) + ws_bytes((size_t)KP0 * g.Hc, 4) + ws_bytes((size_t)nlb * 5, 8);
    DPPO_TRY(ws_reserve(h, need, s));
    TsMlp ma, mc; ts_actor_mlp(h, DPPO_NET_ACTOR_FT, ma); ts_critic_mlp(h, mc);
    ts_mlp_take(h, N, ma, true); ts_mlp_take(h, N, mc, true);
    float* eps = ws_take<float>(h, (size_t)N * g.A); float* deps = ws_take<float>(h, (size_t)N * g.A);
    float* val = ws_take<float>(h, N); float* dval = ws_take<float>(h, N);
    SplitT depsb = ts_take(h, (size_t)N * 64, 2), dvalb = ts_take(h, (size_t)N * 64, 2);
    float* part = ws_take<float>(h, pf);
    float* dw0a = ws_take<float>(h, (size_t)KP0 * g.H); float* dw0c = ws_take<float>(h, (size_t)KP0 * g.Hc);
    double* bsum = ws_take<double>(h, (size_t)nlb * 5);
    ma.out = eps; mc.out = val;
    if (adv_std < 0.f) { adv_stats_kernel<<<1, 1024, 0, s>>>(advantages, N, h->scalars); TC_KCHECK(h); }
    else { set_scalars_kernel<<<1, 1, 0, s>>>(h->scalars, adv_mean, adv_std); TC_KCHECK(h); }
    // h0 straight from (prev, obs, K-1-inds): tconst = -(K) flags "t = K-1-trow[r]"; the critic reads the same tile (its x / one-hot rows of W0 are zero)
    ts_pack_h0_kernel<<<tc_nblk((size_t)N * (KP0 / 8), 256), 256, 0, s>>>(prev, obs, inds, -g.K, N, g.A, g.Do, g.T, KP0, 1, ma.h0, 0);
    TC_KCHECK(h);
    mc.h0 = ma.h0;
    DPPO_TRY(ts_mlp_forward(h, s, ma, N));
    DPPO_TRY(ts_mlp_forward(h, s, mc, N));
    PpoHyper hp;
    hp.A = g.A; hp.Da = h->cfg.action_dim; hp.K = g.K; hp.T = g.T; hp.reward_horizon = h->cfg.reward_horizon; hp.norm_adv = h->cfg.norm_adv;
    hp.dcv = h->cfg.denoised_clip_value; hp.min_lp_std = h->cfg.min_logprob_denoising_std;
    hp.lp_lo = h->cfg.logprob_clip_lo; hp.lp_hi = h->cfg.logprob_clip_hi; hp.gamma_d = h->cfg.gamma_denoising;
    hp.clip_coef = h->cfg.clip_ploss_coef; hp.clip_base = h->cfg.clip_ploss_coef_base; hp.clip_rate = h->cfg.clip_ploss_coef_rate;
    hp.clip_v = h->cfg.clip_vloss_coef; hp.vf_coef = h->cfg.vf_coef; hp.inv_nglobal = 1.0f / (float)N_global;
    ppo_loss_kernel<<<nlb, 128, 0, s>>>(prev, nxt, eps, inds, returns, oldvalues, advantages, oldlogp, val, h->scalars, h->sched, hp, N, deps, dval, bsum);
    TC_KCHECK(h);
    ppo_metrics_kernel<<<1, 256, 0, s>>>(bsum, nlb, hp.inv_nglobal, (float)((double)N / (double)N_global), gr + nA + nC); TC_KCHECK(h);
    ts_pad64_kernel<<<tc_nblk((size_t)N * 8, 256), 256, 0, s>>>(deps, N, g.A, depsb.p[0], depsb.p[1]); TC_KCHECK(h);
    ts_pad64_kernel<<<tc_nblk((size_t)N * 8, 256), 256, 0, s>>>(dval, N, 1, dvalb.p[0], dvalb.p[1]); TC_KCHECK(h);
    DPPO_TRY(ts_actor_grads(h, s, DPPO_NET_ACTOR_FT, ma, deps, depsb, N, part, dw0a, gr));
    DPPO_TRY(ts_critic_grads(h, s, mc, dval, dvalb, N, part, dw0c, gr + nA));
    return 0;
}

// DiffusionModel.c_loss / p_losses (diffusion.py:179-202) + tape.gradient: loss -> h->grads[nA], gradients -> h->grads[0 : nA]
static int ts_pretrain_grads(dppo_handle* h, cudaStream_t s, const float* actions, const float* obs, int N, int64_t N_global,
                             int64_t row_offset, const int32_t* t_in, const float* noise_in, uint64_t seed, uint64_t offset) {
    const Geom& g = h->g; const int KP0 = h->ts->KP0;
    const size_t nA = g.ao.n; float* gr = h->grads;
    const size_t ne = (size_t)N * g.A;
    const int nlb = tc_nblk(ne, 256);
    const size_t pf = tc_part_floats(h, g.H);
    const size_t need = ts_mlp_ws_bytes(N, g.H, KP0, h->cfg.actor_act == DPPO_ACT_MISH, true) + 4 * ws_bytes(ne, 4) + ws_bytes(N, 4)
                      + 2 * ws_bytes((size_t)N * 64, 2) + ws_bytes(pf, 4) + ws_bytes((size_t)KP0 * g.H, 4) + ws_bytes(nlb, 8);
    DPPO_TRY(ws_reserve(h, need, s));
    TsMlp ma; ts_actor_mlp(h, DPPO_NET_ACTOR, ma);
    ts_mlp_take(h, N, ma, true);
    float* eps = ws_take<float>(h, ne); float* deps = ws_take<float>(h, ne); float* noise = ws_take<float>(h, ne); float* xn = ws_take<float>(h, ne);
    int* trow = ws_take<int>(h, N);
    SplitT depsb = ts_take(h, (size_t)N * 64, 2);
    float* part = ws_take<float>(h, pf);
    float* dw0 = ws_take<float>(h, (size_t)KP0 * g.H);
    double* bsum = ws_take<double>(h, nlb);
    ma.out = eps;
    pretrain_prep_kernel<<<tc_nblk(ne, 256), 256, 0, s>>>(actions, t_in, noise_in, N, g.A, g.T, h->sched, seed, offset, row_offset, trow, noise, xn); TC_KCHECK(h);
    ts_pack_h0_kernel<<<tc_nblk((size_t)N * (KP0 / 8), 256), 256, 0, s>>>(xn, obs, trow, 0, N, g.A, g.Do, g.T, KP0, 1, ma.h0, 0); TC_KCHECK(h);
    DPPO_TRY(ts_mlp_forward(h, s, ma, N));
    const float scale = 1.0f / ((float)N_global * (float)g.A);
    mse_loss_kernel<<<nlb, 256, 0, s>>>(eps, noise, ne, scale, deps, bsum); TC_KCHECK(h);
    sum_blocks_kernel<<<1, 256, 0, s>>>(bsum, nlb, scale, gr + nA); TC_KCHECK(h);
    ts_pad64_kernel<<<tc_nblk((size_t)N * 8, 256), 256, 0, s>>>(deps, N, g.A, depsb.p[0], depsb.p[1]); TC_KCHECK(h);
    DPPO_TRY(ts_actor_grads(h, s, DPPO_NET_ACTOR, ma, deps, depsb, N, part, dw0, gr));
    return 0;
}
