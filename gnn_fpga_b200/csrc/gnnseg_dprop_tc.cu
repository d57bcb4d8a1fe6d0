// The propagating half of the dense backward step on tcgen05 (hidden_dim 32): per 128-node tile
//     dH  = dproj . WP^T   (hidden columns)        dz = dH * (1 - H_t^2)                 -> dz_out (read by wgrad_tc_kernel)
//     dg' = dz . W4        (t > 0)                 dg_{t-1} = dg' * (1 - h1_{t-1}^2)     -> dg_out
// (gnn/model.py:113-125,140-156 backwards; the weight gradients of the same step are gnnseg_wgrad_tc.cu.)  Both products
// contract over COLUMNS of node-major rows, i.e. K-major operands in the canonical no-swizzle layout of the forward's
// MLP kernels; 3xTF32 (hi.hi + lo.hi + hi.lo), fp32 accumulation in tensor memory.
//
// One CTA per SM, 13 warps on mbarriers:
//   4 loader warps   dproj, one 32-column block of the tile per stage (ring of two): coalesced reads, hi / lo split,
//                    conflict-free 16-byte stores into the K-major image; the next block's loads are in flight meanwhile
//   1 MMA warp       dH of tile t + 1 is issued before dz . W4 of tile t, so the tensor pipe has work while the epilogue
//                    warps turn dH into dz; D_B and D_D are double buffered in tensor memory
//   8 epilogue warps two sets of four that alternate tiles (the serial chain of a tile's epilogue is what bounded the
//                    kernel); lane = node: D_B -> dz -> global + K-major hi / lo image for the second product; D_D -> dg -> global.
//                    Row arrays pass through swizzled staging tiles per warp so that global memory is read and written in
//                    full 128-byte rows; the H and h1 rows of the next tile arrive there by cp.async a tile ahead
#include <cstdio>
#include <cstdlib>
#include "gnnseg_tc.cuh"

namespace gnnseg {

namespace {

// -DGNNSEG_DTRACE: CTA 0 prints, per role, the cycles its lane 0 spent in every kind of barrier wait (make dtrace)
#ifdef GNNSEG_DTRACE
#define DT_DECL(n) long long dt_acc[n] = {}; const long long dt_start = clock64()
#define DT_WAIT(i, ...) do { const long long dt_t0 = clock64(); __VA_ARGS__; dt_acc[i] += clock64() - dt_t0; } while (0)
#else
#define DT_DECL(n)
#define DT_WAIT(i, ...) do { __VA_ARGS__; } while (0)
#endif

__device__ __forceinline__ void mbar_arrive_d(const uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}

template <int H, int NB>
struct DCfg {
    static constexpr int TN = 128;                             // nodes per tile
    static constexpr int NBLK = NB * H / 32;                   // 32-column blocks of dproj
    static constexpr int A_HALF = TN * 32 * 4;                 // one precision of a stage: 128 x 32 floats
    static constexpr int A_STAGE = 2 * A_HALF;
    static constexpr int STAGES = 2;
    static constexpr int SBO_A = 8 * 128;                      // 8-row group stride of a 32-column K-major image
    static constexpr int SBO_W = (NB * H / 4) * 128;           // of the WP image (NB*H columns)
    static constexpr int SBO_Z = (H / 4) * 128;                // of the dz and W4 images (H columns)
    static constexpr int WP_HALF = H * NB * H * 4, W4_HALF = H * H * 4, DZ_HALF = TN * H * 4;
    static constexpr int OFF_RING = 0;
    static constexpr int OFF_WP = OFF_RING + STAGES * A_STAGE;
    static constexpr int OFF_W4 = OFF_WP + 2 * WP_HALF;
    static constexpr int OFF_DZ = OFF_W4 + 2 * W4_HALF;
    static constexpr int STG_TILE = 32 * H * 4;              // a warp's 32 x H staging tile
    static constexpr int OFF_STG = OFF_DZ + 2 * DZ_HALF;     // per epilogue warp: H rows | h1 rows (each doubles as the staging of the outgoing rows)
    static constexpr int SMEM_BYTES = OFF_STG + 8 * 2 * STG_TILE + 128;
    // the hi and lo halves of a weight image are consecutive ROWS of one image: [W_hi ; W_lo] is an N = 2H operand, so that
    // A_hi . [W_hi ; W_lo]^T gives hi.hi and hi.lo in one instruction (columns 0..H-1 and H..2H-1 of D; the epilogue adds them)
    static constexpr int COL_B = 0, COL_D = 4 * H, TMEM_COLS = 8 * H;      // D_B x 2, D_D x 2, 2H columns each
    static constexpr int NT = 13 * 32;
    // barriers
    static constexpr int FULL = 0, EMPTY = STAGES, DB_FULL = 2 * STAGES, DB_EMPTY = DB_FULL + 2, DD_FULL = DB_EMPTY + 2,
                         DD_EMPTY = DD_FULL + 2, DZ_FULL = DD_EMPTY + 2, DZ_EMPTY = DZ_FULL + 1, NBAR = DZ_EMPTY + 1;
    static_assert(SMEM_BYTES <= 232448, "shared memory");
    static_assert(H == 32, "weight images of wider layers do not fit next to the ring");
};

template <int H, int NB, bool FIRST>
__global__ void __launch_bounds__(DCfg<H, NB>::NT, 1)
dprop_tc_kernel(const float* __restrict__ blob, const float* __restrict__ dproj, const float* __restrict__ H_in,
                const float* __restrict__ h1_prev, const int n_nodes, const int n_tiles, float* __restrict__ dz_out,
                float* __restrict__ dg_out, const int mode) {
    using C = DCfg<H, NB>;
    using B = Blob<H>;
    constexpr int S = C::STAGES, TN = C::TN;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[C::NBAR];
    __shared__ uint32_t tmem_slot;
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](const int i) { return bar0 + 8u * i; };
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(bar(C::FULL + s), 4); mbar_init(bar(C::EMPTY + s), 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(C::DB_FULL + i), 1); mbar_init(bar(C::DB_EMPTY + i), 4);
            mbar_init(bar(C::DD_FULL + i), 1); mbar_init(bar(C::DD_EMPTY + i), 4);
        }
        mbar_init(bar(C::DZ_FULL), 4); mbar_init(bar(C::DZ_EMPTY), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // weight images: WP rows k < H as [N = k][K = o] (blob row k of WP^T is already that), W4^T likewise; hi / lo
    for (int i = tid; i < H * (NB * H / 4); i += C::NT) {
        const int k = i / (NB * H / 4), c = i % (NB * H / 4);
        const float4 v = ldg4(blob + B::WP + k * 5 * H + 4 * c);
        float4 hi, lo;
        split3(v.x, hi.x, lo.x); split3(v.y, hi.y, lo.y); split3(v.z, hi.z, lo.z); split3(v.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(smem + C::OFF_WP + canon_off(k, 4 * c, C::SBO_W)) = hi;
        *reinterpret_cast<float4*>(smem + C::OFF_WP + canon_off(k + H, 4 * c, C::SBO_W)) = lo;
    }
    if (!FIRST) {
        for (int i = tid; i < H * (H / 4); i += C::NT) {
            const int k = i / (H / 4), c = i % (H / 4);
            const float4 v = ldg4(blob + B::W4 + k * H + 4 * c);
            float4 hi, lo;
            split3(v.x, hi.x, lo.x); split3(v.y, hi.y, lo.y); split3(v.z, hi.z, lo.z); split3(v.w, hi.w, lo.w);
            *reinterpret_cast<float4*>(smem + C::OFF_W4 + canon_off(k, 4 * c, C::SBO_Z)) = hi;
            *reinterpret_cast<float4*>(smem + C::OFF_W4 + canon_off(k + H, 4 * c, C::SBO_Z)) = lo;
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int n_mine = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;       // >= 1: grid <= n_tiles
    auto tile_of = [&](const int t) { return (int)blockIdx.x + t * (int)gridDim.x; };

    if (warp < 4) {
        // ---------------- loaders: stage g = (tile t, block b) ----------------
        // item i of a block: row = 8 (i >> 6) + (i & 7), 16-byte chunk = 4 ((i >> 5) & 1) + ((i >> 3) & 3): a warp stores
        // 8 rows x 4 chunks = 512 contiguous bytes of the image and reads 8 x 64 contiguous bytes of dproj
        const int n_stage = n_mine * C::NBLK;
        DT_DECL(1);
        auto load = [&](float4 (&v)[8], const int g) {
            const int node0 = tile_of(g / C::NBLK) * TN, b = g % C::NBLK;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = tid + 128 * j, row = ((i >> 6) << 3) + (i & 7), ch = (((i >> 5) & 1) << 2) + ((i >> 3) & 3);
                const int n = node0 + row;
                v[j] = n < n_nodes ? ldg4(dproj + (size_t)n * (NB * H) + 32 * b + 4 * ch) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto store = [&](const float4 (&v)[8], const int g) {
            const int s = g % S, round = g / S;
            DT_WAIT(0, mbar_wait(bar(C::EMPTY + s), (round & 1) ^ 1));
            unsigned char* st = smem + C::OFF_RING + s * C::A_STAGE;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = tid + 128 * j, row = ((i >> 6) << 3) + (i & 7), ch = (((i >> 5) & 1) << 2) + ((i >> 3) & 3);
                float4 hi, lo;
                split3(v[j].x, hi.x, lo.x); split3(v[j].y, hi.y, lo.y); split3(v[j].z, hi.z, lo.z); split3(v[j].w, hi.w, lo.w);
                const int o = canon_off(row, 4 * ch, C::SBO_A);
                *reinterpret_cast<float4*>(st + o) = hi;
                *reinterpret_cast<float4*>(st + C::A_HALF + o) = lo;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_d(bar(C::FULL + s));
        };
        float4 v0[8], v1[8];                                  // the next block's loads are in flight behind the one being stored
        load(v0, 0);
        for (int g = 0; g < n_stage; g += 2) {
            if (g + 1 < n_stage) load(v1, g + 1);
            store(v0, g);
            if (g + 1 < n_stage) { if (g + 2 < n_stage) load(v0, g + 2); store(v1, g + 1); }
        }
#ifdef GNNSEG_DTRACE
        if (blockIdx.x == 0 && tid == 0) printf("dprop<%d,%d> loader : total %lld, wait empty %lld (%d stages)\n", NB, (int)FIRST, clock64() - dt_start, dt_acc[0], n_stage);
#endif
    } else if (warp == 4) {
      if (lane == 0) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t ID1 = idesc_tf32(128, H), ID2 = idesc_tf32(128, 2 * H);
        const uint32_t sb = smem_u32(smem);
        DT_DECL(4);
        // descriptors: everything but the start address is fixed, and the address field (bits 0..13, 16-byte units) takes the
        // k-step / block / precision offsets by plain addition (no carry out of the field below 256 KB)
        const uint64_t d_wp = smem_desc(sb + C::OFF_WP, 128, C::SBO_W), d_w4 = smem_desc(sb + C::OFF_W4, 128, C::SBO_Z);
        const uint64_t d_dz = smem_desc(sb + C::OFF_DZ, 128, C::SBO_Z);
        auto gemm_b = [&](const int t) {                      // D_B[t & 1] = dproj tile . [WP_hi ; WP_lo]^T
            DT_WAIT(0, mbar_wait(bar(C::DB_EMPTY + (t & 1)), ((t >> 1) & 1) ^ 1));
            tc_fence_after();
            const uint32_t d = tmem + C::COL_B + (t & 1) * 2 * H;
            for (int b = 0; b < C::NBLK; ++b) {
                const int g = t * C::NBLK + b, s = g % S, round = g / S;
                DT_WAIT(1, mbar_wait(bar(C::FULL + s), round & 1));
                tc_fence_after();
                const uint64_t d_a = smem_desc(sb + C::OFF_RING + s * C::A_STAGE, 128, C::SBO_A);
                const uint64_t d_w = d_wp + (uint64_t)(b * 8 * 128 >> 4);
#pragma unroll
                for (int kq = 0; kq < 4; ++kq) {
                    umma_ss(d, d_a + (kq * 256 >> 4), d_w + (kq * 256 >> 4), ID2, (b > 0 || kq > 0) ? 1u : 0u);             // hi . [hi ; lo]
                    umma_ss(d, d_a + ((C::A_HALF + kq * 256) >> 4), d_w + (kq * 256 >> 4), ID1, 1u);                          // lo . hi
                }
                umma_commit(bar(C::EMPTY + s));
            }
            umma_commit(bar(C::DB_FULL + (t & 1)));
        };
        auto gemm_d = [&](const int t) {                      // D_D[t & 1] = dz tile . [W4_hi ; W4_lo]
            DT_WAIT(2, mbar_wait(bar(C::DZ_FULL), t & 1));
            DT_WAIT(3, mbar_wait(bar(C::DD_EMPTY + (t & 1)), ((t >> 1) & 1) ^ 1));
            tc_fence_after();
            const uint32_t d = tmem + C::COL_D + (t & 1) * 2 * H;
#pragma unroll
            for (int kq = 0; kq < H / 8; ++kq) {
                umma_ss(d, d_dz + (kq * 256 >> 4), d_w4 + (kq * 256 >> 4), ID2, kq > 0 ? 1u : 0u);
                umma_ss(d, d_dz + ((C::DZ_HALF + kq * 256) >> 4), d_w4 + (kq * 256 >> 4), ID1, 1u);
            }
            umma_commit(bar(C::DZ_EMPTY));
            umma_commit(bar(C::DD_FULL + (t & 1)));
        };
        gemm_b(0);
        for (int t = 0; t < n_mine; ++t) {
            if (mode == 0 && t + 1 < n_mine) gemm_b(t + 1);   // dH of the next tile first: the tensor pipe works under the epilogue
            if (!FIRST) gemm_d(t);
            if (mode != 0 && t + 1 < n_mine) gemm_b(t + 1);   // dz . W4 first: the epilogue's second half is not held up
        }
#ifdef GNNSEG_DTRACE
        if (blockIdx.x == 0) printf("dprop<%d,%d> mma    : total %lld, wait db_empty %lld, full %lld, dz_full %lld, dd_empty %lld (%d tiles)\n", NB, (int)FIRST, clock64() - dt_start, dt_acc[0], dt_acc[1], dt_acc[2], dt_acc[3], n_mine);
#endif
      }
      __syncwarp();
    } else {
        // ---------------- epilogue: lane = node ----------------
        // A thread owns a node (the TMEM lane), but 32 own-row accesses per instruction touch 32 lines: every row
        // array goes through the warp's swizzled 32 x 32 staging tile (16-byte chunk c of row r at chunk c ^ (r & 7)),
        // so that global memory sees four full 128-byte rows per instruction and shared memory no conflicts either way.
        const int q = warp & 3;                               // the TMEM lane quarter this warp may read
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        const int row = q * 32 + lane;
        const int eset = (warp - 5) >> 2;                     // two sets of four warps: set e takes the CTA's tiles t = e, e + 2, ...
        float* s_h = reinterpret_cast<float*>(smem + C::OFF_STG + (warp - 5) * 2 * C::STG_TILE);
        float* s_1 = s_h + 32 * H;
        const int rq = lane >> 3, jj = lane & 7;
        DT_DECL(3);
        // rows of the warp's 32 nodes -> staging tile, asynchronously (no registers held while they are in flight); one group
        auto prefetch = [&](float* tile, const float* src, const int nw0, const bool any) {
            if (any) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = 4 * i + rq, n = nw0 + r;
                    if (n < n_nodes) cp_async16(tile + r * 32 + ((jj ^ (r & 7)) << 2), src + (size_t)n * H + 4 * jj);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto own_row = [&](const float* tile, float (&own)[H]) {       // staged rows -> the lane's own row
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 y = lds4(tile + lane * 32 + ((c ^ (lane & 7)) << 2));
                own[4 * c] = y.x; own[4 * c + 1] = y.y; own[4 * c + 2] = y.z; own[4 * c + 3] = y.w;
            }
            __syncwarp();
        };
        auto from_own = [&](const float (&own)[H], float* s_o, float* dst, const int nw0) {   // the lane's own row -> global, coalesced, staged in s_o
#pragma unroll
            for (int c = 0; c < 8; ++c) st4(s_o + lane * 32 + ((c ^ (lane & 7)) << 2), make_float4(own[4 * c], own[4 * c + 1], own[4 * c + 2], own[4 * c + 3]));
            __syncwarp();
            float4 y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { const int r = 4 * i + rq; y[i] = lds4(s_o + r * 32 + ((jj ^ (r & 7)) << 2)); }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int n = nw0 + 4 * i + rq;
                if (n < n_nodes) st4(dst + (size_t)n * H + 4 * jj, y[i]);
            }
            __syncwarp();
        };
        // D[0..H-1] + D[H..2H-1] of this lane: the hi.hi + lo.hi and the hi.lo parts of a product
        auto tmem_sum = [&](const uint32_t col, float (&v)[H]) {
            float w[H];
            tmem_ld32(lane_base + col, v);
            tmem_ld32(lane_base + col + H, w);
#pragma unroll
            for (int k = 0; k < H; ++k) v[k] += w[k];
        };
        prefetch(s_h, H_in, tile_of(eset) * TN + q * 32, eset < n_mine);
        if (!FIRST) prefetch(s_1, h1_prev, tile_of(eset) * TN + q * 32, eset < n_mine);
        for (int t = eset; t < n_mine; t += 2) {
            const int nw0 = tile_of(t) * TN + q * 32;
            const bool more = t + 2 < n_mine;
            float hv[H], v[H];
            if (FIRST) asm volatile("cp.async.wait_group 0;" ::: "memory");
            else asm volatile("cp.async.wait_group 1;" ::: "memory");          // all but the newest group: the H rows are there
            __syncwarp();
            own_row(s_h, hv);
            DT_WAIT(0, mbar_wait(bar(C::DB_FULL + (t & 1)), (t >> 1) & 1));
            tc_fence_after();
            tmem_sum(C::COL_B + (t & 1) * 2 * H, v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_d(bar(C::DB_EMPTY + (t & 1)));
#pragma unroll
            for (int k = 0; k < H; ++k) v[k] *= fmaf(-hv[k], hv[k], 1.f);        // dz
            if (!FIRST) {
                DT_WAIT(1, mbar_wait(bar(C::DZ_EMPTY), (t & 1) ^ 1));     // dz . W4 of the previous tile has read the image
#pragma unroll
                for (int c = 0; c < H / 4; ++c) {
                    float4 hi, lo;
                    split3(v[4 * c], hi.x, lo.x); split3(v[4 * c + 1], hi.y, lo.y); split3(v[4 * c + 2], hi.z, lo.z); split3(v[4 * c + 3], hi.w, lo.w);
                    const int o = canon_off(row, 4 * c, C::SBO_Z);
                    *reinterpret_cast<float4*>(smem + C::OFF_DZ + o) = hi;
                    *reinterpret_cast<float4*>(smem + C::OFF_DZ + C::DZ_HALF + o) = lo;
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_d(bar(C::DZ_FULL));
            }
            from_own(v, s_h, dz_out, nw0);
            prefetch(s_h, H_in, tile_of(t + 2) * TN + q * 32, more);           // the H rows of this set's next tile
            if (!FIRST) {
                asm volatile("cp.async.wait_group 1;" ::: "memory");           // the h1 rows (older than the H rows just requested)
                __syncwarp();
                own_row(s_1, hv);
                DT_WAIT(2, mbar_wait(bar(C::DD_FULL + (t & 1)), (t >> 1) & 1));
                tc_fence_after();
                tmem_sum(C::COL_D + (t & 1) * 2 * H, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_d(bar(C::DD_EMPTY + (t & 1)));
#pragma unroll
                for (int k = 0; k < H; ++k) v[k] *= fmaf(-hv[k], hv[k], 1.f);    // dg
                from_own(v, s_1, dg_out, nw0);
                prefetch(s_1, h1_prev, tile_of(t + 2) * TN + q * 32, more);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#ifdef GNNSEG_DTRACE
        if (blockIdx.x == 0 && (warp == 5 || warp == 9) && lane == 0) printf("dprop<%d,%d> epilog : total %lld, wait db_full %lld, dz_empty %lld, dd_full %lld\n", NB, (int)FIRST, clock64() - dt_start, dt_acc[0], dt_acc[1], dt_acc[2]);
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS) : "memory");
}

template <int H, int NB, bool FIRST>
int launch_dprop(const float* blob, const float* dproj, const float* H_in, const float* h1_prev, const int n_nodes, float* dz_out,
                 float* dg_out, const int sms, cudaStream_t st) {
    using C = DCfg<H, NB>;
    if (!ensure_dynamic_smem<dprop_tc_kernel<H, NB, FIRST>>(C::SMEM_BYTES)) return GNNSEG_ECUDA;
    const int n_tiles = (n_nodes + C::TN - 1) / C::TN;
    const int grid = n_tiles < sms ? n_tiles : sms;
    static const int mode = [] { const char* v = getenv("GNNSEG_DPROP_MODE"); return v ? atoi(v) : 0; }();
    dprop_tc_kernel<H, NB, FIRST><<<grid, C::NT, C::SMEM_BYTES, st>>>(blob, dproj, H_in, h1_prev, n_nodes, n_tiles, dz_out, dg_out, mode);
    return cudaGetLastError() == cudaSuccess ? GNNSEG_OK : GNNSEG_ECUDA;
}

}  // namespace

bool dprop_tc_width(const int h) { return h == 32; }

int dprop_tc(const int h, const int nb, const bool first, const float* blob, const float* dproj, const float* H_in, const float* h1_prev,
             const int n_nodes, float* dz_out, float* dg_out, const int sms, cudaStream_t st) {
    if (n_nodes <= 0) return GNNSEG_OK;
    if (h != 32) return GNNSEG_EUNSUPPORTED;
    if (nb == 2) return first ? launch_dprop<32, 2, true>(blob, dproj, H_in, h1_prev, n_nodes, dz_out, dg_out, sms, st)
                              : launch_dprop<32, 2, false>(blob, dproj, H_in, h1_prev, n_nodes, dz_out, dg_out, sms, st);
    return first ? launch_dprop<32, 5, true>(blob, dproj, H_in, h1_prev, n_nodes, dz_out, dg_out, sms, st)
                 : launch_dprop<32, 5, false>(blob, dproj, H_in, h1_prev, n_nodes, dz_out, dg_out, sms, st);
}

}  // namespace gnnseg
