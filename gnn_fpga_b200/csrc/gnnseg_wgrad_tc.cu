// Weight gradients of the dense backward step on tcgen05 (hidden_dim 32 / 64): every contraction of
// Estimator.training_step's backward (gnn/estimator.py:49-60 through gnn/model.py:113-125,140-156) that sums over the
// NODES,
//     dWP^T[o][k] = sum_n dproj[n][o] * [H_t | X | 1][n][k]        (d[W1a|W1b|W3a|W3b|W3c] and their biases: the ones column)
//     dW4[o][k]   = sum_n dz[n][o] * h1[n][k]                        (node network, second layer)
//     [dWin|dbin] / db4 = sum_n dz[n][o] * [X | 1][n]                (input network at t = 0, b4 otherwise)
// as 3xTF32 products (hi.hi + lo.hi + hi.lo, fp32 accumulation in tensor memory over ALL node tiles of the CTA).
//
// Both operands of such a product are MN-major: the contraction index (the node) is the row of the node-major arrays.
// tcgen05.mma kind::tf32 takes MN-major operands in ONE shared-memory layout, SWIZZLE_128B_BASE32B
// (scripts/micro/umma_mn32_probe.cu; SWIZZLE_NONE returns zeros): per node a 128-byte row of 32 consecutive columns,
// four rows per 512-byte atom, the 32-byte chunks of a row XOR-ed with (node & 3); blocks of 32 columns LBO apart,
// groups of four nodes SBO = 512 bytes apart.  A float4 of a node-major row lands in it as ONE 16-byte store: no
// transposition anywhere, and the image of a block serves as A (its columns = rows of D) and as B alike.
//
// Two kernels: wgrad_tc_kernel right below (hidden_dim 64, and 32 with GNNSEG_WGRAD=split): two products per k-step;
// wgrad_tc_merged_kernel further down (hidden_dim 32, the default there): one product per k-step, bulk-copy fed.
// wgrad_tc_kernel: one CTA per SM, 9 warps: 8 loader warps (a thread owns one 16-byte chunk of one node per block: coalesced 128-byte
// row pieces in, tf32 hi / lo split, two 16-byte stores out) fill a ring of three stages of SN nodes; one thread issues
// the MMAs of a stage and commits them to the stage's `empty` barrier.  Nothing is read back per tile: the accumulators
// stay in tensor memory until the CTA's last stage, then four warps add them to the CTA's slot of the partial buffer
// (layout NodePart of gnnseg_backward.cu; the same grid on every launch => a fixed summation order, bit-reproducible).
// D rows beyond the real columns (M = 128 reads four blocks, whatever lies behind) are never looked at.
#include <cstdio>
#include <cstdlib>
#include "gnnseg_tc.cuh"

namespace gnnseg {

namespace {

#ifdef GNNSEG_DTRACE
#define DT_DECL(n) long long dt_acc[n] = {}; const long long dt_start = clock64()
#define DT_WAIT(i, ...) do { const long long dt_t0 = clock64(); __VA_ARGS__; dt_acc[i] += clock64() - dt_t0; } while (0)
#else
#define DT_DECL(n)
#define DT_WAIT(i, ...) do { __VA_ARGS__; } while (0)
#endif

__device__ __forceinline__ void mbar_arrive_w(const uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
// MN-major, SWIZZLE_128B_BASE32B (layout type 1), version 1
__device__ __forceinline__ uint64_t smem_desc_mn32(const uint32_t addr, const uint32_t lbo, const uint32_t sbo) {
    return smem_desc(addr, lbo, sbo) | ((uint64_t)1 << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32_mn(const int M, const int N) {
    return idesc_tf32(M, N) | (1u << 15) | (1u << 16);
}
// byte offset of the 16-byte chunk c16 (columns 4 c16 .. 4 c16 + 3 of a 32-column block) of node `node` inside a block image
__device__ __forceinline__ int mn32_off(const int node, const int c16) {
    return ((node >> 2) << 9) + ((node & 3) << 7) + ((((c16 >> 1) ^ node) & 3) << 5) + ((c16 & 1) << 4);
}

template <int H, int NB>
struct WCfg {
    static constexpr int SN = (H == 32) ? 32 : 16;           // nodes per stage
    static constexpr int KS = SN / 8;                         // MMA k-steps per stage
    static constexpr int BI = SN * 128;                       // bytes of one block image
    static constexpr int PB = NB * H / 32, HB = H / 32;       // blocks of dproj; of H, dz, h1
    // One group of blocks per precision (hi, lo), laid out so that every operand's blocks are equally spaced:
    //   HB = 1: [h1 | dz | H | X]            HB = 2: [h1.0 | dz.0 | dz.1 | h1.1 | H.0 | H.1 | X]
    // [H.. | X] are the rows of the projection products' A operand (contiguous), [h1.. | X] the columns of the dz
    // products' B operand (three blocks apart), dz.. the rows of their A operand (contiguous).
    static constexpr int GB = 3 * HB + 1;
    static constexpr int O_H1 = 0, H1_STRIDE = 3 * BI, O_DZ = BI, O_HXH = 2 * HB * BI, O_HXX = 3 * HB * BI;
    static constexpr int G_HI = 0, G_LO = GB * BI, P_HI = 2 * GB * BI, P_LO = P_HI + PB * BI;
    static constexpr int STAGE = P_LO + PB * BI;
    static constexpr int STAGES = 3;
    static constexpr int SMEM_BYTES = STAGES * STAGE + 1024;  // + alignment slack
    static constexpr int MA = (HB + 1) * 32 <= 64 ? 64 : 128; // rows of the projection products: [H | X 1 0 ..]
    static constexpr int NA = NB * H, NA_SPLIT = NA > 256 ? 2 : 1, NA1 = NA / NA_SPLIT;
    static constexpr int MC = 64;                             // rows of the dz products (H <= 64)
    static constexpr int COL_A = 0, COL_C = NA, COLS = NA + H + 16;
    static constexpr int TMEM_COLS = COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
    static constexpr int NT = 288;                            // 8 loader warps + the MMA warp
    static_assert(STAGES * STAGE + 1024 <= 232448, "shared memory");
    static_assert(COLS <= 512 && NA1 <= 256 && NA1 % 16 == 0, "tensor memory / N");
};

template <int H>
struct NodePartW {       // = NodePart<H> of gnnseg_backward.cu
    static constexpr int D4 = H + 4;
    static constexpr int WP = 0, BP = WP + D4 * 5 * H, W4 = BP + 5 * H, B4 = W4 + H * H, WIN = B4 + H, BIN = WIN + 4 * H,
                         SIZE = BIN + H;
};

// FIRST: the input step (t = 0): no h1, dz feeds dWin / dbin.  accumulate: 0 = store, 1 = add to the projection slots
// (WIN / BIN are written once), 2 = add to the W4 / B4 slots as well.
template <int H, int NB, bool FIRST>
__global__ void __launch_bounds__(WCfg<H, NB>::NT, 1)
wgrad_tc_kernel(const float* __restrict__ dproj, const float* __restrict__ H_in, const float* __restrict__ X4,
                const float* __restrict__ dz, const float* __restrict__ h1_prev, const int n_nodes, const int n_chunks,
                float* __restrict__ part, const int accumulate) {
    using C = WCfg<H, NB>;
    using NP = NodePartW<H>;
    constexpr int SN = C::SN, BI = C::BI, S = C::STAGES;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bars[2 * S + 1];
    __shared__ uint32_t tmem_slot;
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full = [&](const int s) { return bar0 + 8u * s; };
    auto empty = [&](const int s) { return bar0 + 8u * (S + s); };
    const uint32_t done = bar0 + 8u * (2 * S);
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full(s), 8); mbar_init(empty(s), 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int n_mine = (n_chunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // >= 1: grid <= n_chunks

    if (warp < 8) {
        // ---------------- loaders ----------------
        constexpr int PER_BLK = SN * 8;                      // 16-byte chunks of one block image
        constexpr int NP_ITEMS = C::PB * PER_BLK / 256, NH_ITEMS = (C::HB * PER_BLK + 255) / 256;
        static_assert(C::PB * PER_BLK % 256 == 0 && C::HB * PER_BLK % 256 == 0, "loader split");
        DT_DECL(1);
        struct StageRegs { float4 vp[NP_ITEMS], vh[NH_ITEMS], vz[NH_ITEMS], v1[NH_ITEMS], vx; };
        // the global loads of stage `it` (a thread's chunks of every block), all issued before anything is consumed
        auto load = [&](StageRegs& R, const int it) {
            const int node0 = ((int)blockIdx.x + it * (int)gridDim.x) * SN;
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < NP_ITEMS; ++j) {
                const int i = tid + 256 * j, blk = i / PER_BLK, r = i % PER_BLK, ng = node0 + (r >> 3);
                R.vp[j] = ng < n_nodes ? ldg4(dproj + (size_t)ng * (NB * H) + 32 * blk + 4 * (r & 7)) : zero;
            }
#pragma unroll
            for (int j = 0; j < NH_ITEMS; ++j) {
                const int i = tid + 256 * j, blk = i / PER_BLK, r = i % PER_BLK, ng = node0 + (r >> 3);
                const bool ok = ng < n_nodes;
                const size_t off = (size_t)ng * H + 32 * blk + 4 * (r & 7);
                R.vh[j] = ok ? ldg4(H_in + off) : zero;
                R.vz[j] = ok ? ldg4(dz + off) : zero;
                R.v1[j] = (!FIRST && ok) ? ldg4(h1_prev + off) : zero;
            }
            R.vx = zero;
            if (tid < SN * 4) {                               // the X block: [x0 x1 x2 x3 | 1 0 0 0 | 0 ... ]: 16 columns are read
                const int ng = node0 + (tid >> 2), c16 = tid & 3;
                if (ng < n_nodes) {
                    if (c16 == 0) R.vx = ldg4(X4 + (size_t)ng * 4);
                    else if (c16 == 1) R.vx.x = 1.f;
                }
            }
        };
        // tf32 hi / lo split and the two 16-byte stores per chunk, once the MMAs that read the stage last have completed
        auto store = [&](const StageRegs& R, const int it) {
            const int s = it % S, round = it / S;
            DT_WAIT(0, mbar_wait(empty(s), (round & 1) ^ 1));
            unsigned char* st = smem + s * C::STAGE;
            auto put = [&](const float4 v, const int o) {     // o: offset inside the hi group / the hi P blocks
                float4 hi, lo;
                split3(v.x, hi.x, lo.x); split3(v.y, hi.y, lo.y); split3(v.z, hi.z, lo.z); split3(v.w, hi.w, lo.w);
                *reinterpret_cast<float4*>(st + o) = hi;
                *reinterpret_cast<float4*>(st + o + (o >= C::P_HI ? C::PB * BI : C::G_LO)) = lo;
            };
#pragma unroll
            for (int j = 0; j < NP_ITEMS; ++j) {
                const int i = tid + 256 * j, blk = i / PER_BLK, r = i % PER_BLK;
                put(R.vp[j], C::P_HI + blk * BI + mn32_off(r >> 3, r & 7));
            }
#pragma unroll
            for (int j = 0; j < NH_ITEMS; ++j) {
                const int i = tid + 256 * j, blk = i / PER_BLK, r = i % PER_BLK, o = mn32_off(r >> 3, r & 7);
                put(R.vh[j], C::O_HXH + blk * BI + o);
                put(R.vz[j], C::O_DZ + blk * BI + o);
                if (!FIRST) put(R.v1[j], C::O_H1 + blk * C::H1_STRIDE + o);
            }
            if (tid < SN * 4) put(R.vx, C::O_HXX + mn32_off(tid >> 2, tid & 3));
            fence_async_smem();                               // generic-proxy writes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive_w(full(s));
        };
        StageRegs r0, r1, r2;                                 // two stages of loads in flight behind the one being stored
        load(r0, 0);
        if (1 < n_mine) load(r1, 1);
        for (int it = 0; it < n_mine; it += 3) {
            if (it + 2 < n_mine) load(r2, it + 2);
            store(r0, it);
            if (it + 1 < n_mine) { if (it + 3 < n_mine) load(r0, it + 3); store(r1, it + 1); }
            if (it + 2 < n_mine) { if (it + 4 < n_mine) load(r1, it + 4); store(r2, it + 2); }
        }
#ifdef GNNSEG_DTRACE
        if (blockIdx.x == 0 && tid == 0) printf("wgrad<%d,%d> loader : total %lld, wait empty %lld (%d stages)\n", NB, (int)FIRST, clock64() - dt_start, dt_acc[0], n_mine);
#endif
    } else {
      if (lane == 0) {
        // ---------------- MMA issuer ----------------
        constexpr int NC = FIRST ? 16 : H + 16;
        constexpr uint32_t ID_A = idesc_tf32_mn(C::MA, C::NA1), ID_C = idesc_tf32_mn(C::MC, NC);
        const uint32_t sbase = smem_u32(smem);
        DT_DECL(1);
        for (int it = 0; it < n_mine; ++it) {
            const int s = it % S, round = it / S;
            DT_WAIT(0, mbar_wait(full(s), round & 1));
            tc_fence_after();
            const uint32_t st = sbase + s * C::STAGE;
#pragma unroll
            for (int ks = 0; ks < C::KS; ++ks) {
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {        // hi.hi, lo.hi, hi.lo
                    const uint32_t ga = st + (pass == 1 ? C::G_LO : C::G_HI) + ks * 1024;      // A side: [H | X], dz
                    const uint32_t gb = st + (pass == 2 ? C::G_LO : C::G_HI) + ks * 1024;      // B side: [h1 | X]
                    const uint32_t pb = st + (pass == 2 ? C::P_LO : C::P_HI) + ks * 1024;      // B side: dproj
                    const uint32_t acc = (it > 0 || ks > 0 || pass > 0) ? 1u : 0u;
                    // D_A[k][o] += sum_n [H | X 1][n][k] * dproj[n][o]
                    const uint64_t a_hx = smem_desc_mn32(ga + C::O_HXH, BI, 512);
#pragma unroll
                    for (int h = 0; h < C::NA_SPLIT; ++h)
                        umma_ss(tmem + C::COL_A + h * C::NA1, a_hx, smem_desc_mn32(pb + h * (C::NA1 / 32) * BI, BI, 512), ID_A, acc);
                    // D_C[o][k | f] += sum_n dz[n][o] * [h1 | X 1][n][k | f]
                    const uint64_t a_z = smem_desc_mn32(ga + C::O_DZ, BI, 512);
                    const uint64_t b_c = FIRST ? smem_desc_mn32(gb + C::O_HXX, BI, 512) : smem_desc_mn32(gb + C::O_H1, C::H1_STRIDE, 512);
                    umma_ss(tmem + C::COL_C, a_z, b_c, ID_C, acc);
                }
            }
            umma_commit(empty(s));                            // arrives when the MMAs above have read the stage
        }
        umma_commit(done);
#ifdef GNNSEG_DTRACE
        if (blockIdx.x == 0) printf("wgrad<%d,%d> mma    : total %lld, wait full %lld\n", NB, (int)FIRST, clock64() - dt_start, dt_acc[0]);
#endif
      }
      __syncwarp();
    }

    // ---------------- epilogue: accumulators -> the CTA's partial ----------------
    // row r of an M = 64 product sits in TMEM lane 32 (r / 16) + r % 16, of an M = 128 product in lane r
    if (warp < 4) {
        mbar_wait(done, 0);
        tc_fence_after();
        float* mine = part + (size_t)blockIdx.x * NP::SIZE;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        const bool add_p = accumulate != 0, add_4 = accumulate > 1;
        {
            const int k = C::MA == 64 ? (lane < 16 ? 16 * warp + lane : -1) : 32 * warp + lane;      // row of D_A: column of [H | X 1]
            float* dst = nullptr;                             // this row of the partial: 5H floats apart from the next
            if (k >= 0 && k < NP::D4) dst = mine + NP::WP + k * 5 * H;
            else if (k == NP::D4) dst = mine + NP::BP;        // the ones column: bias sums
#pragma unroll 1
            for (int c0 = 0; c0 < C::NA; c0 += 16) {
                float v[16];
                tmem_ld16(lane_base + C::COL_A + c0, v);
                if (dst) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        float4 w = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        float* p = dst + c0 + i;
                        if (add_p) { const float4 o = *reinterpret_cast<const float4*>(p); w.x += o.x; w.y += o.y; w.z += o.z; w.w += o.w; }
                        st4(p, w);
                    }
                }
            }
        }
        {
            const int o = lane < 16 ? 16 * warp + lane : -1;  // row of D_C: column of dz
            const bool ok = o >= 0 && o < H;
            if (!FIRST) {
#pragma unroll 1
                for (int c0 = 0; c0 < H; c0 += 16) {
                    float v[16];
                    tmem_ld16(lane_base + C::COL_C + c0, v);
                    if (ok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float* p = mine + NP::W4 + o * H + c0 + i;
                            *p = add_4 ? *p + v[i] : v[i];
                        }
                    }
                }
            }
            float v[16];
            tmem_ld16(lane_base + C::COL_C + (FIRST ? 0 : H), v);
            if (ok) {
                if (FIRST) {
#pragma unroll
                    for (int f = 0; f < 4; ++f) mine[NP::WIN + f * H + o] = v[f];
                    mine[NP::BIN + o] = v[4];
                } else {
                    float* p = mine + NP::B4 + o;
                    *p = add_4 ? *p + v[4] : v[4];
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS) : "memory");
}

// ---- hidden_dim 32: ONE product per k-step and precision pair ------------------------------------------------------
// The issuing thread needs ~50 - 110 cycles per tcgen05.mma whatever its size (wait-time counters, make dtrace: the kernel
// above spends 86 % of its time issuing 36 small products per 32 nodes), so at hidden_dim 32 all of a step's
// contractions go into one M = 128, N <= 208 product:
//     A rows (m-blocks)   [ H_t (32) | X 1 0.. (32) | dz (32) | (32 rows of whatever follows: ignored) ]
//     B columns (n-blocks) [ dproj (32 NB) | h1 (32, t > 0) | X 1 0.. (16) ]
// D rows 0..36 x the dproj columns = d[W1a|W1b|W3a|W3b|W3c]^T and their biases, rows 64..95 x the h1 columns = dW4, x the
// last block = db4 (t > 0) or dWin | dbin (t = 0); the other quadrants are computed and never read.  Stages of 16 nodes
// (two k-steps): a raw ring of five (bulk copies) in front of an image ring of three.
template <int NB, bool FIRST>
struct MCfg {
    static constexpr int H = 32, SN = 16, KS = SN / 8, BI = SN * 128, PB = NB;
    static constexpr int O_HXH = 0, O_XA = BI, O_DZ = 2 * BI, O_P = 3 * BI, O_H1 = (3 + PB) * BI, O_XB = (3 + PB + (FIRST ? 0 : 1)) * BI;
    static constexpr int GB = 3 + PB + (FIRST ? 0 : 1) + 1;
    static constexpr int G_LO = GB * BI, STAGE = 2 * GB * BI, STAGES = 3;
    // raw ring: the node-major rows of a stage as they lie in global memory (contiguous per array), brought in by bulk
    // copies: the loaders' own loads are capped by the L1's outstanding misses (~12 B/clk/SM measured, 3.4 TB/s chip-wide)
    static constexpr int ROW_P = NB * H * 4, ROW_H = H * 4;
    static constexpr int R_P = 0, R_H = SN * ROW_P, R_Z = R_H + SN * ROW_H, R_1 = R_Z + SN * ROW_H, R_X = R_1 + (FIRST ? 0 : SN * ROW_H);
    static constexpr int RAW = R_X + SN * 16, RAW_STAGES = 5;
    static constexpr int OFF_RAW = STAGES * STAGE;
    static constexpr int SMEM_BYTES = OFF_RAW + RAW_STAGES * RAW + 1024;
    static constexpr int N = 32 * PB + (FIRST ? 0 : 32) + 16;
    static constexpr int COL_H1 = 32 * PB, COL_X = 32 * PB + (FIRST ? 0 : 32);
    static constexpr int TMEM_COLS = N <= 128 ? 128 : 256;
    static constexpr int ARRAYS = PB + (FIRST ? 2 : 3);       // 32-column row arrays per stage: dproj blocks, H, dz, h1
    static constexpr int ITEMS = (ARRAYS * SN * 8 + 255) / 256;
    static constexpr int NT = 320;                            // 8 loader warps, the MMA warp, the copy warp
    static_assert(SMEM_BYTES <= 232448 && N <= 256 && N % 16 == 0 && RAW % 128 == 0, "merged product");
};

__device__ __forceinline__ void mbar_expect_tx_w(const uint32_t mbar, const uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_w(const uint32_t dst_smem, const void* src, const uint32_t bytes, const uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

template <int NB, bool FIRST>
__global__ void __launch_bounds__(MCfg<NB, FIRST>::NT, 1)
wgrad_tc_merged_kernel(const float* __restrict__ dproj, const float* __restrict__ H_in, const float* __restrict__ X4,
                       const float* __restrict__ dz, const float* __restrict__ h1_prev, const int n_nodes, const int n_chunks,
                       float* __restrict__ part, const int accumulate) {
    using C = MCfg<NB, FIRST>;
    using NP = NodePartW<32>;
    constexpr int H = 32, SN = C::SN, BI = C::BI, S = C::STAGES, RS = C::RAW_STAGES;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bars[2 * S + 2 * RS + 1];
    __shared__ uint32_t tmem_slot;
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full = [&](const int s) { return bar0 + 8u * s; };
    auto empty = [&](const int s) { return bar0 + 8u * (S + s); };
    auto raw_full = [&](const int s) { return bar0 + 8u * (2 * S + s); };
    auto raw_empty = [&](const int s) { return bar0 + 8u * (2 * S + RS + s); };
    const uint32_t done = bar0 + 8u * (2 * S + 2 * RS);
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full(s), 8); mbar_init(empty(s), 1); }
        for (int s = 0; s < RS; ++s) { mbar_init(raw_full(s), 1); mbar_init(raw_empty(s), 8); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int n_mine = (n_chunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // >= 1: grid <= n_chunks

    if (warp < 8) {
        // ---------------- loaders: raw rows -> tf32 hi / lo block images ----------------
        // item i = 128 a + r: array a (dproj block a < PB, then H, dz, h1), node r >> 3 of the stage, 16-byte chunk r & 7
        DT_DECL(2);
        for (int it = 0; it < n_mine; ++it) {
            const int node0 = ((int)blockIdx.x + it * (int)gridDim.x) * SN;
            const int sr = it % RS, s = it % S;
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            const unsigned char* raw = smem + C::OFF_RAW + sr * C::RAW;
            DT_WAIT(1, mbar_wait(raw_full(sr), (it / RS) & 1));
            float4 v[C::ITEMS], vx = zero;
#pragma unroll
            for (int j = 0; j < C::ITEMS; ++j) {
                const int i = tid + 256 * j, a = i >> 7, r = i & 127, nl = r >> 3, c = 16 * (r & 7);
                v[j] = zero;
                if (a < C::ARRAYS && node0 + nl < n_nodes) {
                    const int off = a < C::PB ? C::R_P + nl * C::ROW_P + 128 * a
                                  : (a == C::PB ? C::R_H : a == C::PB + 1 ? C::R_Z : C::R_1) + nl * C::ROW_H;
                    v[j] = *reinterpret_cast<const float4*>(raw + off + c);
                }
            }
            if (tid < SN * 4 && node0 + (tid >> 2) < n_nodes) {            // the X block: [x0 x1 x2 x3 | 1 0 0 0 | 0 ...]
                if ((tid & 3) == 0) vx = *reinterpret_cast<const float4*>(raw + C::R_X + 16 * (tid >> 2));
                else if ((tid & 3) == 1) vx.x = 1.f;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_w(raw_empty(sr));     // the rows are in registers: the copy warp may refill the slot
            DT_WAIT(0, mbar_wait(empty(s), ((it / S) & 1) ^ 1));
            unsigned char* st = smem + s * C::STAGE;
            auto put = [&](const float4 x, const int o) {
                float4 hi, lo;
                split3(x.x, hi.x, lo.x); split3(x.y, hi.y, lo.y); split3(x.z, hi.z, lo.z); split3(x.w, hi.w, lo.w);
                *reinterpret_cast<float4*>(st + o) = hi;
                *reinterpret_cast<float4*>(st + C::G_LO + o) = lo;
            };
#pragma unroll
            for (int j = 0; j < C::ITEMS; ++j) {
                const int i = tid + 256 * j, a = i >> 7, r = i & 127;
                if (a < C::ARRAYS) {
                    const int blk = a < C::PB ? C::O_P + a * BI : a == C::PB ? C::O_HXH : a == C::PB + 1 ? C::O_DZ : C::O_H1;
                    put(v[j], blk + mn32_off(r >> 3, r & 7));
                }
            }
            if (tid < SN * 4) {
                const int o = mn32_off(tid >> 2, tid & 3);
                put(vx, C::O_XA + o);
                put(vx, C::O_XB + o);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_w(full(s));
        }
#ifdef GNNSEG_DTRACE
        if (blockIdx.x == 0 && tid == 0) printf("wgradM<%d,%d> loader : total %lld, wait empty %lld, raw %lld (%d stages)\n", NB, (int)FIRST, clock64() - dt_start, dt_acc[0], dt_acc[1], n_mine);
#endif
    } else if (warp == 9) {
      if (lane == 0) {
        // ---------------- copy warp: global rows -> raw ring ----------------
        const uint32_t sraw = smem_u32(smem) + C::OFF_RAW;
        for (int it = 0; it < n_mine; ++it) {
            const int node0 = ((int)blockIdx.x + it * (int)gridDim.x) * SN, sr = it % RS;
            const int nv = n_nodes - node0 < SN ? n_nodes - node0 : SN;
            mbar_wait(raw_empty(sr), ((it / RS) & 1) ^ 1);
            const uint32_t dst = sraw + sr * C::RAW, bar = raw_full(sr);
            mbar_expect_tx_w(bar, (uint32_t)nv * (C::ROW_P + (FIRST ? 2 : 3) * C::ROW_H + 16));
            bulk_g2s_w(dst + C::R_P, dproj + (size_t)node0 * (NB * H), nv * C::ROW_P, bar);
            bulk_g2s_w(dst + C::R_H, H_in + (size_t)node0 * H, nv * C::ROW_H, bar);
            bulk_g2s_w(dst + C::R_Z, dz + (size_t)node0 * H, nv * C::ROW_H, bar);
            if (!FIRST) bulk_g2s_w(dst + C::R_1, h1_prev + (size_t)node0 * H, nv * C::ROW_H, bar);
            bulk_g2s_w(dst + C::R_X, X4 + (size_t)node0 * 4, nv * 16, bar);
        }
      }
      __syncwarp();
    } else {
      if (lane == 0) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t ID = idesc_tf32_mn(128, C::N);
        const uint32_t sbase = smem_u32(smem);
        DT_DECL(1);
        for (int it = 0; it < n_mine; ++it) {
            const int s = it % S, round = it / S;
            DT_WAIT(0, mbar_wait(full(s), round & 1));
            tc_fence_after();
            const uint32_t st = sbase + s * C::STAGE;
            const uint64_t d_a = smem_desc_mn32(st + C::O_HXH, BI, 512), d_b = smem_desc_mn32(st + C::O_P, BI, 512);
#pragma unroll
            for (int ks = 0; ks < C::KS; ++ks) {
                const uint64_t k = (uint64_t)(ks * 1024 >> 4), lo = (uint64_t)(C::G_LO >> 4);
                umma_ss(tmem, d_a + k, d_b + k, ID, (it > 0 || ks > 0) ? 1u : 0u);       // hi . hi
                umma_ss(tmem, d_a + k + lo, d_b + k, ID, 1u);                             // lo . hi
                umma_ss(tmem, d_a + k, d_b + k + lo, ID, 1u);                             // hi . lo
            }
            umma_commit(empty(s));
        }
        umma_commit(done);
#ifdef GNNSEG_DTRACE
        if (blockIdx.x == 0) printf("wgradM<%d,%d> mma    : total %lld, wait full %lld\n", NB, (int)FIRST, clock64() - dt_start, dt_acc[0]);
#endif
      }
      __syncwarp();
    }

    // ---------------- epilogue: row r of D is TMEM lane r ----------------
    if (warp < 3) {
        mbar_wait(done, 0);
        tc_fence_after();
        float* mine = part + (size_t)blockIdx.x * NP::SIZE;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        const bool add_p = accumulate != 0, add_4 = accumulate > 1;
        if (warp < 2) {                                       // rows of [H | X 1]: a row of the projections' partial each
            float* dst = nullptr;
            if (warp == 0) dst = mine + NP::WP + lane * 5 * H;
            else if (lane < 4) dst = mine + NP::WP + (H + lane) * 5 * H;
            else if (lane == 4) dst = mine + NP::BP;
#pragma unroll 1
            for (int c0 = 0; c0 < 32 * C::PB; c0 += 16) {
                float v[16];
                tmem_ld16(lane_base + c0, v);
                if (dst) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        float4 w = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        float* p = dst + c0 + i;
                        if (add_p) { const float4 o = *reinterpret_cast<const float4*>(p); w.x += o.x; w.y += o.y; w.z += o.z; w.w += o.w; }
                        st4(p, w);
                    }
                }
            }
        } else {                                              // rows of dz: lane = column o of dz
            const int o = lane;
            if (!FIRST) {
#pragma unroll 1
                for (int c0 = 0; c0 < H; c0 += 16) {
                    float v[16];
                    tmem_ld16(lane_base + C::COL_H1 + c0, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float* p = mine + NP::W4 + o * H + c0 + i;
                        *p = add_4 ? *p + v[i] : v[i];
                    }
                }
            }
            float v[16];
            tmem_ld16(lane_base + C::COL_X, v);
            if (FIRST) {
#pragma unroll
                for (int f = 0; f < 4; ++f) mine[NP::WIN + f * H + o] = v[f];
                mine[NP::BIN + o] = v[4];
            } else {
                float* p = mine + NP::B4 + o;
                *p = add_4 ? *p + v[4] : v[4];
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS) : "memory");
}

}  // namespace

bool wgrad_tc_width(const int h) { return h == 32 || h == 64; }

// grid of the weight-gradient kernel: a function of (n_nodes, SM count) only, the same on every launch of a backward pass
// GNNSEG_WGRAD=split: the two-product kernel at hidden_dim 32 as well (A/B runs)
static bool wgrad_merged(const int h) {
    static const bool split = [] { const char* v = getenv("GNNSEG_WGRAD"); return v && v[0] == 's'; }();
    return h == 32 && !split;
}
int wgrad_tc_grid(const int n_nodes, const int h, const int sms) {
    const int sn = (h == 32 && !wgrad_merged(h)) ? 32 : 16, n_chunks = (n_nodes + sn - 1) / sn;
    return n_chunks < sms ? n_chunks : sms;
}

template <int H, int NB, bool FIRST>
static int launch_wgrad(const float* dproj, const float* H_in, const float* X4, const float* dz, const float* h1_prev, const int n_nodes,
                        float* part, const int accumulate, const int grid, cudaStream_t st) {
    using C = WCfg<H, NB>;
    if (!ensure_dynamic_smem<wgrad_tc_kernel<H, NB, FIRST>>(C::SMEM_BYTES)) return GNNSEG_ECUDA;
    const int n_chunks = (n_nodes + C::SN - 1) / C::SN;
    wgrad_tc_kernel<H, NB, FIRST><<<grid, C::NT, C::SMEM_BYTES, st>>>(dproj, H_in, X4, dz, h1_prev, n_nodes, n_chunks, part, accumulate);
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

template <int NB, bool FIRST>
static int launch_wgrad_merged(const float* dproj, const float* H_in, const float* X4, const float* dz, const float* h1_prev, const int n_nodes,
                               float* part, const int accumulate, const int grid, cudaStream_t st) {
    using C = MCfg<NB, FIRST>;
    if (!ensure_dynamic_smem<wgrad_tc_merged_kernel<NB, FIRST>>(C::SMEM_BYTES)) return GNNSEG_ECUDA;
    const int n_chunks = (n_nodes + C::SN - 1) / C::SN;
    wgrad_tc_merged_kernel<NB, FIRST><<<grid, C::NT, C::SMEM_BYTES, st>>>(dproj, H_in, X4, dz, h1_prev, n_nodes, n_chunks, part, accumulate);
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

int wgrad_tc(const int h, const int nb, const bool first, const float* dproj, const float* H_in, const float* X4, const float* dz,
             const float* h1_prev, const int n_nodes, float* part, const int accumulate, const int grid, cudaStream_t st) {
    if (n_nodes <= 0) return GNNSEG_OK;
    if (wgrad_merged(h)) {
        if (nb == 2) return first ? launch_wgrad_merged<2, true>(dproj, H_in, X4, dz, h1_prev, n_nodes, part, accumulate, grid, st)
                                  : launch_wgrad_merged<2, false>(dproj, H_in, X4, dz, h1_prev, n_nodes, part, accumulate, grid, st);
        return first ? launch_wgrad_merged<5, true>(dproj, H_in, X4, dz, h1_prev, n_nodes, part, accumulate, grid, st)
                     : launch_wgrad_merged<5, false>(dproj, H_in, X4, dz, h1_prev, n_nodes, part, accumulate, grid, st);
    }
#define GNNSEG_W(HH) \
    if (h == HH) { \
        if (nb == 2) return first ? launch_wgrad<HH, 2, true>(dproj, H_in, X4, dz, h1_prev, n_nodes, part, accumulate, grid, st) \
                                  : launch_wgrad<HH, 2, false>(dproj, H_in, X4, dz, h1_prev, n_nodes, part, accumulate, grid, st); \
        return first ? launch_wgrad<HH, 5, true>(dproj, H_in, X4, dz, h1_prev, n_nodes, part, accumulate, grid, st) \
                     : launch_wgrad<HH, 5, false>(dproj, H_in, X4, dz, h1_prev, n_nodes, part, accumulate, grid, st); \
    }
    GNNSEG_W(32)
    GNNSEG_W(64)
#undef GNNSEG_W
    return GNNSEG_EUNSUPPORTED;
}

}  // namespace gnnseg
