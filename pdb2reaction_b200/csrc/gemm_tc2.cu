// tcgen05 tensor-core GEMM, TMA-fed on BOTH operands:  C[M,N] = A[M,K] . W[N,K]^T (+bias), fp32-grade accuracy.
//
// Same bf16 "x3" arithmetic as gemm_tc.cu (a.w ~= a_hi w_hi + a_hi w_lo + a_lo w_hi, fp32 accumulation in TMEM),
// but the activation operand arrives PRE-SPLIT: the kernel that produces an activation (gather/rotate, gate,
// LayerNorm+SiLU and the adjoint kernels) writes it as two bf16 planes (hi, lo; 2 + 2 bytes per element, the
// same HBM bytes as one fp32) instead of fp32.  The GEMM main loop is then nothing but TMA loads and
// tcgen05.mma -- no producer warps, no LDG/convert/STS traffic through l1tex (the limiter of gemm_tc.cu: ncu
// l1tex 45-65 %, tensor pipe 28-63 %, profiles/r01_ncu_gemm_tc_v3_in_bench_raw.csv).
//
// Persistent, one CTA per SM, 320 threads:
//   warp 0      TMEM allocator; lane 0 = TMA producer (A_hi, A_lo, W_hi, W_lo tiles of one k block per stage)
//   warp 1      barrier init; lane 0 = tcgen05.mma issuer (3 MMAs per 16-wide k step), tcgen05.commit
//   warps 2-9   epilogue: tcgen05.ld (lane quadrant = warp % 4, 32-column chunks interleaved between the two
//               warps of a quadrant) -> (+bias) -> 128B-swizzled staging tile in smem -> TMA tensor store
//               (cp.async.bulk.tensor global <- shared; rows beyond M are clipped by the tensor map).  It overlaps
//               the next tile's main loop through the double-buffered TMEM accumulator.
// BK = 64 bf16 (SWIZZLE_128B rows).
#include <map>
#include <mutex>
#include <tuple>
#include <unordered_map>

#include "tc_ptx.cuh"

namespace umab {

namespace {

using namespace tcp;

constexpr int BM = 128;
constexpr int EPI_WARPS = 8;
constexpr int T2_THREADS = 64 + EPI_WARPS * 32;
constexpr int STG_BYTES = 32 * 128;            // one 32 x 32 fp32 staging tile per epilogue warp
constexpr int SMEM_LIMIT = 227 * 1024;

struct Tc2Params {
    const float* bias;
    int M, N, K, BN, stages, tmem_cols;
    // fused gate epilogue (pair kernel; see GemmArgs)
    const float* gate; int gate_ld, gate_mode;
    __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
};

template <int BK>
__global__ void __launch_bounds__(T2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tm_ah, const __grid_constant__ CUtensorMap tm_al,
                const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_wl,
                const __grid_constant__ CUtensorMap tm_c, Tc2Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int ROW_BYTES = BK * 2;
    constexpr uint32_t A_PLANE = BM * ROW_BYTES;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int S = p.stages;
    const int BN = p.BN;
    const uint32_t w_plane = (uint32_t)BN * ROW_BYTES;
    const uint32_t stage_bytes = 2u * A_PLANE + 2u * w_plane;
    const uint32_t stg_base = base + (uint32_t)S * stage_bytes;
    const uint32_t bar_base = stg_base + EPI_WARPS * STG_BYTES;
    auto a_hi = [&](int s) { return base + (uint32_t)s * stage_bytes; };
    auto a_lo = [&](int s) { return a_hi(s) + A_PLANE; };
    auto w_hi = [&](int s) { return a_hi(s) + 2u * A_PLANE; };
    auto w_lo = [&](int s) { return w_hi(s) + w_plane; };
    auto full = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto empty = [&](int s) { return bar_base + 8u * (uint32_t)(S + s); };
    auto tmem_full = [&](int b) { return bar_base + 8u * (uint32_t)(2 * S + b); };
    auto tmem_empty = [&](int b) { return bar_base + 8u * (uint32_t)(2 * S + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(2 * S + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int nkb = p.K / BK;
    const int ntn = p.N / BN;                                   // tiles along N (fastest: concurrent CTAs share A through L2)
    const int num_tiles = ntn * ((p.M + BM - 1) / BM);
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(2 * p.tmem_cols)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer: one elected lane streams the four operand tiles of every k block
        if (lane == 0) {
            int g = 0;
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int tile = (int)blockIdx.x + lt * (int)gridDim.x;
                const int n0 = (tile % ntn) * BN;
                const int m0 = (tile / ntn) * BM;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % S;
                    const uint32_t ph = (uint32_t)(g / S) & 1u;
                    mbar_wait(empty(s), ph ^ 1u);
                    mbar_arrive_expect_tx(full(s), stage_bytes);
                    tma_load_2d(a_hi(s), &tm_ah, full(s), kb * BK, m0);
                    tma_load_2d(w_hi(s), &tm_wh, full(s), kb * BK, n0);
                    tma_load_2d(a_lo(s), &tm_al, full(s), kb * BK, m0);
                    tma_load_2d(w_lo(s), &tm_wl, full(s), kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (double-buffered TMEM accumulators)
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(BN);
            int g = 0;
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int ab = lt & 1;
                mbar_wait(tmem_empty(ab), (((uint32_t)lt >> 1) & 1u) ^ 1u);   // epilogue drained this buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(ab * p.tmem_cols);
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % S;
                    const uint32_t ph = (uint32_t)(g / S) & 1u;
                    mbar_wait(full(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t dah = make_smem_desc<ROW_BYTES>(a_hi(s)), dal = make_smem_desc<ROW_BYTES>(a_lo(s));
                    const uint64_t dwh = make_smem_desc<ROW_BYTES>(w_hi(s)), dwl = make_smem_desc<ROW_BYTES>(w_lo(s));
#pragma unroll
                    for (int j = 0; j < BK / 16; ++j) {
                        const uint64_t adv = (uint64_t)(j * 2);      // 32 B per k step, in 16 B units
                        umma_bf16(d_tmem, dah + adv, dwh + adv, idesc, (kb > 0 || j > 0) ? 1u : 0u);
                        umma_bf16(d_tmem, dah + adv, dwl + adv, idesc, 1u);
                        umma_bf16(d_tmem, dal + adv, dwh + adv, idesc, 1u);
                    }
                    umma_commit(empty(s));       // frees the stage when these MMAs have read it
                }
                umma_commit(tmem_full(ab));      // accumulator of this tile complete
            }
        }
    } else {
        // ===================== epilogue warps: TMEM -> registers (+bias) -> swizzled staging tile -> TMA store
        const int ew = warp - 2;
        const int q = warp & 3;                                   // TMEM lane quadrant this warp may read
        const int half = ew >> 2;                                 // which of the two warps of the quadrant
        const uint32_t stg = stg_base + (uint32_t)ew * STG_BYTES;
        const uint32_t row_addr = stg + (uint32_t)lane * 128u;
        const uint32_t sw = (uint32_t)(lane & 7);
        const int nch = BN / 32;
        for (int lt = 0; lt < my_tiles; ++lt) {
            const int tile = (int)blockIdx.x + lt * (int)gridDim.x;
            const int n0 = (tile % ntn) * BN;
            const int mrow0 = (tile / ntn) * BM + q * 32;
            const int ab = lt & 1;
            mbar_wait(tmem_full(ab), ((uint32_t)lt >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + (uint32_t)(ab * p.tmem_cols) + ((uint32_t)(q * 32) << 16);
            for (int ch = half; ch < nch; ch += 2) {
                uint32_t rr[32];
                tmem_ld32_async(taddr + (uint32_t)(ch * 32), rr);
                float4 bv[8];
                if (p.bias) {
                    const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + ch * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) bv[j] = __ldg(bp + j);
                }
                tmem_ld_wait();
                // the previous TMA store of this warp must have finished reading the staging tile
                if (lane == 0) tma_store_wait_read0();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]),
                                           __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
                    if (p.bias) o = f4add(o, bv[j]);
                    const uint32_t addr = row_addr + ((((uint32_t)j) ^ sw) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && mrow0 < p.M) {
                    tma_store_2d(&tm_c, stg, n0 + ch * 32, mrow0);
                    tma_store_commit();
                }
            }
            // all TMEM reads of this buffer are complete (tcgen05.wait::ld above)
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(ab));
        }
        if (lane == 0) tma_store_wait_all();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * p.tmem_cols)) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2).  ncu on the single-CTA kernel above: the big conv shapes pull 11.4 TB/s through
// xbar -> l1tex (l1tex__m_xbar2l1tex_read_bytes), i.e. the L2 -> SM fabric cap, at 55 % tensor-pipe activity -- with a
// 128 x 256 tile every k block of 64 moves 32 KB of A planes + 64 KB of W planes per CTA for 12 MMAs.  Here the two
// SMs of a TPC compute one 256 x BN tile: each CTA loads its own 128 rows of A but only HALF of the W tile (BN / 2
// rows); the MMA, issued by the leader CTA (cluster rank 0), reads the B operand from both CTAs' shared memory and
// accumulates rows 0-127 into the leader's TMEM and rows 128-255 into the peer's.  L2 -> SM bytes per MMA drop by a
// third (64 KB instead of 96 KB per k block and CTA) and the smaller stage gives a third pipeline stage.
//   full[s]       leader only: 1 arrival (leader's producer, expect_tx of BOTH CTAs' bytes); the peer's TMA loads
//                 complete_tx on it through the cta_group::2 load form
//   empty[s]      one per CTA, armed by the leader's tcgen05.commit multicast
//   tmem_full[b]  one per CTA, same multicast commit;  tmem_empty[b]  leader only, 2 x EPI_WARPS (remote) arrivals
template <int BK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
gemm_tc2_pair_kernel(const __grid_constant__ CUtensorMap tm_ah, const __grid_constant__ CUtensorMap tm_al,
                     const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_wl,
                     const __grid_constant__ CUtensorMap tm_c, Tc2Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int ROW_BYTES = BK * 2;
    constexpr uint32_t A_PLANE = BM * ROW_BYTES;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int S = p.stages;
    const int BN = p.BN;
    const int HN = BN / 2;                                       // W rows held by each CTA of the pair
    const uint32_t w_plane = (uint32_t)HN * ROW_BYTES;
    const uint32_t stage_bytes = 2u * A_PLANE + 2u * w_plane;
    const uint32_t stg_base = base + (uint32_t)S * stage_bytes;
    const uint32_t bar_base = stg_base + EPI_WARPS * STG_BYTES;
    auto a_hi = [&](int s) { return base + (uint32_t)s * stage_bytes; };
    auto a_lo = [&](int s) { return a_hi(s) + A_PLANE; };
    auto w_hi = [&](int s) { return a_hi(s) + 2u * A_PLANE; };
    auto w_lo = [&](int s) { return w_hi(s) + w_plane; };
    auto full = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto empty = [&](int s) { return bar_base + 8u * (uint32_t)(S + s); };
    auto tmem_full = [&](int b) { return bar_base + 8u * (uint32_t)(2 * S + b); };
    auto tmem_empty = [&](int b) { return bar_base + 8u * (uint32_t)(2 * S + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(2 * S + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = cluster_ctarank();
    const int pair = (int)blockIdx.x / 2, npairs = (int)gridDim.x / 2;
    const int nkb = p.K / BK;
    const int ntn = p.N / BN;
    const int num_tiles = ntn * ((p.M + 2 * BM - 1) / (2 * BM));        // 256-row tiles
    const int my_tiles = (num_tiles - pair + npairs - 1) / npairs;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 2 * EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(2 * p.tmem_cols)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                    // both CTAs' barriers are initialised before any remote arrive / TMA complete_tx
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs): own A rows, own half of the W tile; bytes land on the leader's barrier
        if (lane == 0) {
            int g = 0;
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int tile = pair + lt * npairs;
                const int n0 = (tile % ntn) * BN + (int)rank * HN;
                const int m0 = (tile / ntn) * (2 * BM) + (int)rank * BM;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % S;
                    const uint32_t ph = (uint32_t)(g / S) & 1u;
                    mbar_wait(empty(s), ph ^ 1u);
                    if (rank == 0) mbar_arrive_expect_tx(full(s), 2u * stage_bytes);
                    const uint32_t fb = mapa_cluster(full(s), 0);
                    tma_load_2d_pair(a_hi(s), &tm_ah, fb, kb * BK, m0);
                    tma_load_2d_pair(w_hi(s), &tm_wh, fb, kb * BK, n0);
                    tma_load_2d_pair(a_lo(s), &tm_al, fb, kb * BK, m0);
                    tma_load_2d_pair(w_lo(s), &tm_wl, fb, kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: leader CTA only
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = make_idesc_bf16_pair(BN);
            int g = 0;
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int ab = lt & 1;
                mbar_wait(tmem_empty(ab), (((uint32_t)lt >> 1) & 1u) ^ 1u);   // both CTAs' epilogues drained this buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(ab * p.tmem_cols);
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % S;
                    const uint32_t ph = (uint32_t)(g / S) & 1u;
                    mbar_wait(full(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t dah = make_smem_desc<ROW_BYTES>(a_hi(s)), dal = make_smem_desc<ROW_BYTES>(a_lo(s));
                    const uint64_t dwh = make_smem_desc<ROW_BYTES>(w_hi(s)), dwl = make_smem_desc<ROW_BYTES>(w_lo(s));
#pragma unroll
                    for (int j = 0; j < BK / 16; ++j) {
                        const uint64_t adv = (uint64_t)(j * 2);
                        umma_bf16_pair(d_tmem, dah + adv, dwh + adv, idesc, (kb > 0 || j > 0) ? 1u : 0u);
                        umma_bf16_pair(d_tmem, dah + adv, dwl + adv, idesc, 1u);
                        umma_bf16_pair(d_tmem, dal + adv, dwh + adv, idesc, 1u);
                    }
                    umma_commit_pair(empty(s));       // frees the stage in both CTAs
                }
                umma_commit_pair(tmem_full(ab));      // accumulator halves of both CTAs complete
            }
        }
    } else {
        // ===================== epilogue warps (both CTAs): own 128 rows of the pair tile
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        const uint32_t stg = stg_base + (uint32_t)ew * STG_BYTES;
        const uint32_t row_addr = stg + (uint32_t)lane * 128u;
        const uint32_t sw = (uint32_t)(lane & 7);
        const int nch = BN / 32;
        for (int lt = 0; lt < my_tiles; ++lt) {
            const int tile = pair + lt * npairs;
            const int n0 = (tile % ntn) * BN;
            const int mrow0 = (tile / ntn) * (2 * BM) + (int)rank * BM + q * 32;
            const int ab = lt & 1;
            mbar_wait(tmem_full(ab), ((uint32_t)lt >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + (uint32_t)(ab * p.tmem_cols) + ((uint32_t)(q * 32) << 16);
            for (int ch = half; ch < nch; ch += 2) {
                uint32_t rr[32];
                tmem_ld32_async(taddr + (uint32_t)(ch * 32), rr);
                float4 bv[8];
                if (p.bias) {
                    const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + ch * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) bv[j] = __ldg(bp + j);
                }
                // fused gate: the 32 x 32 tile of this chunk is re-read from the staging tile below with 8 lanes per row, so
                // that gate loads (128 B per row) and plane stores (64 B per row and plane) are whole sectors; the gate
                // values are requested here, while the TMEM load is in flight
                const int col0 = n0 + ch * 32;
                const int sub_r = lane >> 3, sub_j = lane & 7;
                float4 gv[8];
                if (p.gate) {
                    const int gcol = (p.gate_mode == 1 ? ((col0 >> 7) & 1) * 128 : 128) + (col0 & 127) + sub_j * 4;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int grow = mrow0 + it * 4 + sub_r;
                        gv[it] = grow < p.M ? __ldg(reinterpret_cast<const float4*>(p.gate + (long long)grow * p.gate_ld + gcol))
                                            : f4zero();
                    }
                }
                tmem_ld_wait();
                if (lane == 0) tma_store_wait_read0();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]),
                                           __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
                    if (p.bias) o = f4add(o, bv[j]);
                    const uint32_t addr = row_addr + ((((uint32_t)j) ^ sw) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && mrow0 < p.M) {
                    tma_store_2d(&tm_c, stg, n0 + ch * 32, mrow0);
                    tma_store_commit();
                }
                if (p.gate) {
                    // gated bf16 planes of conv-2: the same fp32 product and hi / lo split as combine_gate_fwd_kernel
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = it * 4 + sub_r;
                        const int grow = mrow0 + r;
                        float4 o;
                        const uint32_t addr = stg + (uint32_t)r * 128u + ((((uint32_t)sub_j) ^ (uint32_t)(r & 7)) << 4);
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(addr));
                        if (grow < p.M) {
                            uint2 h, l;
                            split4(f4mul(o, gv[it]), h, l);
                            const long long oi = (long long)grow * p.N + col0 + sub_j * 4;
                            *reinterpret_cast<uint2*>(p.out_hi + oi) = h;
                            *reinterpret_cast<uint2*>(p.out_lo + oi) = l;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(mapa_cluster(tmem_empty(ab), 0));
        }
        if (lane == 0) tma_store_wait_all();
    }
    // neither CTA may leave (or free TMEM) while its partner can still read its shared memory / arrive on its barriers
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * p.tmem_cols)) : "memory");
    }
}

// fp32 [rows, K] -> bf16 hi / lo planes (weights once per engine; activations only in the unit-test entry)
__global__ void split_planes2_kernel(const float* __restrict__ w, long long n, __nv_bfloat16* __restrict__ hi,
                                     __nv_bfloat16* __restrict__ lo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = w[i];
    __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
}

// ------------------------------------------------------------------ host side
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode2() {
    static EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    });
    if (!fn) throw CudaError("cuTensorMapEncodeTiled is not available from the driver");
    return fn;
}

int pick_bn2(int N) {
    if (N % 256 == 0) return 256;
    if (N <= 256) return N % 32 == 0 ? N : 0;
    for (int bn = 256; bn >= 32; bn -= 32)
        if (N % bn == 0) return bn;
    return 0;
}

// row-major [rows, cols] matrix, row pitch `pitch_bytes`, box [box_rows, box_cols]
void make_map2(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* ptr, long long rows, long long cols,
               long long pitch_bytes, int box_rows, int box_cols, CUtensorMapSwizzle sw) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    (void)esize;
    CUresult r = get_encode2()(tm, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[256];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld pitch=%lld box=%dx%d", (int)r, rows,
                 cols, pitch_bytes, box_rows, box_cols);
        throw CudaError(buf);
    }
}

struct Planes2 {
    __nv_bfloat16* hi = nullptr; __nv_bfloat16* lo = nullptr;
    CUtensorMap tm_hi, tm_lo;             // box = [bn, bk]: the whole W tile (single-CTA kernel)
    CUtensorMap tm_hi_half, tm_lo_half;   // box = [bn / 2, bk]: the half each CTA of a pair loads
    int bn = 0;
};

struct MapKey {
    const void* p; long long rows, cols, pitch; int box_rows, box_cols;
    bool operator==(const MapKey& o) const {
        return p == o.p && rows == o.rows && cols == o.cols && pitch == o.pitch && box_rows == o.box_rows && box_cols == o.box_cols;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.p);
        auto mix = [&](long long v) { h ^= std::hash<long long>()(v) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
        mix(k.rows); mix(k.cols); mix(k.pitch); mix(k.box_rows); mix(k.box_cols);
        return h;
    }
};

}  // namespace

// bf16 weight planes + tensor maps of one engine, and the activation / output tensor maps of its (stable)
// workspace buffers
struct Tc2Cache {
    std::mutex mu;
    std::map<std::tuple<const float*, int, int, int>, Planes2> planes;
    std::unordered_map<MapKey, CUtensorMap, MapKeyHash> maps;
    ~Tc2Cache() { clear(); }
    void clear() {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& kv : planes) { cudaFree(kv.second.hi); cudaFree(kv.second.lo); }
        planes.clear();
        maps.clear();
    }
    // activations / outputs: the pointers are workspace buffers that may be re-allocated between calls, but a map
    // only encodes (address, shape), so a stale entry is still a correct description of that address range
    const CUtensorMap& map(CUtensorMapDataType dt, int esize, const void* ptr, long long rows, long long cols,
                           long long pitch, int box_rows, int box_cols, CUtensorMapSwizzle sw) {
        MapKey k{ptr, rows, cols, pitch, box_rows, box_cols};
        auto it = maps.find(k);
        if (it != maps.end()) return it->second;
        if (maps.size() > 8192) maps.clear();
        CUtensorMap tm;
        make_map2(&tm, dt, esize, ptr, rows, cols, pitch, box_rows, box_cols, sw);
        return maps.emplace(k, tm).first->second;
    }
};
Tc2Cache* tc2_cache_create() { return new Tc2Cache(); }
void tc2_cache_destroy(Tc2Cache* c) { delete c; }
void tc2_cache_clear(Tc2Cache* c) { if (c) c->clear(); }

void tc2_split(const float* x, long long n, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st) {
    if (n <= 0) return;
    split_planes2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n, hi, lo);
    UMAB_LAUNCH_CHECK();
}

bool gemm_tc2_supported(const GemmArgs& a, int bk) {
    return a.batch == 1 && a.M >= 1 && a.K % bk == 0 && a.K >= bk && pick_bn2(a.N) >= 32 && a.ldc % 4 == 0 &&
           a.ldw == a.K && !a.accumulate && (a.A_hi != nullptr || a.lda == a.K);
}

namespace {

Planes2 build_planes2(const float* W, int N, int K, int bk, cudaStream_t st) {
    Planes2 p;
    const long long n = (long long)N * K;
    UMAB_CUDA(cudaMalloc(&p.hi, n * 2));
    UMAB_CUDA(cudaMalloc(&p.lo, n * 2));
    split_planes2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, n, p.hi, p.lo);
    UMAB_LAUNCH_CHECK();
    p.bn = pick_bn2(N);
    const CUtensorMapSwizzle sw = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    make_map2(&p.tm_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.hi, N, K, (long long)K * 2, p.bn, bk, sw);
    make_map2(&p.tm_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.lo, N, K, (long long)K * 2, p.bn, bk, sw);
    make_map2(&p.tm_hi_half, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.hi, N, K, (long long)K * 2, p.bn / 2, bk, sw);
    make_map2(&p.tm_lo_half, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.lo, N, K, (long long)K * 2, p.bn / 2, bk, sw);
    return p;
}

template <int BK>
void launch_tc2(const CUtensorMap& ah, const CUtensorMap& al, const Planes2& pl, const CUtensorMap& cm, const Tc2Params& p,
                size_t smem, dim3 grid, cudaStream_t st) {
    static std::once_flag attr_once[16];
    int dev = 0;
    UMAB_CUDA(cudaGetDevice(&dev));
    std::call_once(attr_once[dev & 15], [] {
        UMAB_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    });
    gemm_tc2_kernel<BK><<<grid, T2_THREADS, smem, st>>>(ah, al, pl.tm_hi, pl.tm_lo, cm, p);
    UMAB_LAUNCH_CHECK();
}

// CTAs of the pair kernel that can be resident at once (2 x the co-resident clusters; 0 = cluster launch unsupported)
template <int BK>
int pair_max_ctas(int dev) {
    static int cached[16] = {0};
    static std::once_flag once[16];
    std::call_once(once[dev & 15], [dev] {
        UMAB_CUDA(cudaFuncSetAttribute(gemm_tc2_pair_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        int n_sm = 0;
        UMAB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(n_sm & ~1)); cfg.blockDim = dim3(T2_THREADS); cfg.dynamicSmemBytes = SMEM_LIMIT;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, gemm_tc2_pair_kernel<BK>, &cfg) != cudaSuccess) { cudaGetLastError(); ncl = 0; }
        cached[dev & 15] = std::min(2 * ncl, n_sm & ~1);
    });
    return cached[dev & 15];
}

template <int BK>
void launch_tc2_pair(const CUtensorMap& ah, const CUtensorMap& al, const Planes2& pl, const CUtensorMap& cm,
                     const Tc2Params& p, size_t smem, long long pair_tiles, cudaStream_t st) {
    int dev = 0;
    UMAB_CUDA(cudaGetDevice(&dev));
    const int cap = pair_max_ctas<BK>(dev);
    if (cap < 2) throw CudaError("gemm_tc2: the CTA-pair kernel cannot be launched on this device");
    dim3 grid((unsigned)std::min<long long>(2 * pair_tiles, cap));
    gemm_tc2_pair_kernel<BK><<<grid, T2_THREADS, smem, st>>>(ah, al, pl.tm_hi_half, pl.tm_lo_half, cm, p);
    UMAB_LAUNCH_CHECK();
}

}  // namespace

// k-block width: 64 bf16 (128 B TMA rows, SWIZZLE_128B).  A 32-wide block (twice the stages) was measured in round 1
// (profiles/r01_gemm_pair_shapes.txt, r01_gemm_tc2_shapes.txt): never faster on the CTA-pair kernel, +1 % on two shapes
// of the single-CTA kernel that the pair kernel beats by 15 %; its instantiations were removed in round 2.
// CTA-pair kernel by default (UMAB_TC2_PAIR=0 selects the single-CTA kernel)
bool tc2_pair_default();
// the fused gate epilogue lives in the pair kernel only
bool gemm_tc2_gate_epilogue_available() { return tc2_pair_default(); }
bool tc2_pair_default() {
    static const bool on = [] {
        const char* e = getenv("UMAB_TC2_PAIR");
        return !(e && atoi(e) == 0);
    }();
    return on;
}

// bk = 64 or 32 (0: per-shape choice); pair = 1 / 0: CTA-pair / single-CTA kernel (-1: default).
// cache == nullptr: everything is built for this call only (unit-test entry).
void gemm_tc2(const GemmArgs& a, cudaStream_t st, Tc2Cache* cache, int bk, int pair) {
    // default: the CTA-pair kernel with 64-wide k blocks; the two shapes it does not win (measured,
    // profiles/r01_gemm_pair_shapes.txt) stay on the single-CTA kernel
    if (pair < 0) pair = (tc2_pair_default() && !((a.N == 128 && a.K == 64) || (a.N == 384 && a.K == 128))) ? 1 : 0;
    if (bk == 0) bk = 64;
    if (bk != 64) throw CudaError("gemm_tc2: the k block is 64");
    if (!gemm_tc2_supported(a, bk)) throw CudaError("gemm_tc2: unsupported shape");
    int dev = 0;
    UMAB_CUDA(cudaGetDevice(&dev));
    const CUtensorMapSwizzle sw = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    Planes2 pl;
    __nv_bfloat16 *tmp_hi = nullptr, *tmp_lo = nullptr;
    const __nv_bfloat16 *A_hi = a.A_hi, *A_lo = a.A_lo;
    if (!A_hi) {
        // fp32 activation (unit-test / benchmark entry): split it here
        const long long n = (long long)a.M * a.K;
        UMAB_CUDA(cudaMalloc(&tmp_hi, n * 2));
        UMAB_CUDA(cudaMalloc(&tmp_lo, n * 2));
        split_planes2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.A, n, tmp_hi, tmp_lo);
        UMAB_LAUNCH_CHECK();
        A_hi = tmp_hi; A_lo = tmp_lo;
    }
    CUtensorMap ah, al, cm;
    if (cache) {
        std::lock_guard<std::mutex> lk(cache->mu);
        auto key = std::make_tuple(a.W, a.N, a.K, bk);
        auto it = cache->planes.find(key);
        if (it == cache->planes.end()) it = cache->planes.emplace(key, build_planes2(a.W, a.N, a.K, bk, st)).first;
        pl = it->second;
        ah = cache->map(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A_hi, a.M, a.K, (long long)a.K * 2, BM, bk, sw);
        al = cache->map(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A_lo, a.M, a.K, (long long)a.K * 2, BM, bk, sw);
        cm = cache->map(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.Cmat, a.M, a.N, a.ldc * 4, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    } else {
        pl = build_planes2(a.W, a.N, a.K, bk, st);
        make_map2(&ah, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A_hi, a.M, a.K, (long long)a.K * 2, BM, bk, sw);
        make_map2(&al, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A_lo, a.M, a.K, (long long)a.K * 2, BM, bk, sw);
        make_map2(&cm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.Cmat, a.M, a.N, a.ldc * 4, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    Tc2Params p;
    p.bias = a.bias; p.M = a.M; p.N = a.N; p.K = a.K; p.BN = pl.bn;
    p.gate = a.gate; p.gate_ld = a.gate_ld; p.gate_mode = a.gate_mode; p.out_hi = a.out_hi; p.out_lo = a.out_lo;
    if (a.gate && (!pair || !a.out_hi || !a.out_lo || a.N % 128 != 0 || (a.gate_mode != 1 && a.gate_mode != 2)))
        throw CudaError("gemm_tc2: the fused gate epilogue needs the CTA-pair kernel, output planes and N % 128 == 0");
    const int stage_bytes = 2 * BM * bk * 2 + 2 * (pair ? pl.bn / 2 : pl.bn) * bk * 2;
    const int fixed = EPI_WARPS * STG_BYTES + 1024 /*alignment slack*/ + 8 * (2 * 8 + 5) + 64;
    p.stages = std::max(2, std::min(8, (SMEM_LIMIT - fixed) / stage_bytes));
    int cols = 32;
    while (cols < pl.bn) cols *= 2;
    p.tmem_cols = cols;
    const size_t smem = (size_t)p.stages * stage_bytes + fixed;
    if (smem > (size_t)SMEM_LIMIT) throw CudaError("gemm_tc2: shared memory budget exceeded");
    static int n_sm[16] = {0};
    if (!n_sm[dev & 15]) UMAB_CUDA(cudaDeviceGetAttribute(&n_sm[dev & 15], cudaDevAttrMultiProcessorCount, dev));
    const long long tiles = (long long)(a.N / pl.bn) * ((a.M + BM - 1) / BM);
    dim3 grid((unsigned)std::min<long long>(tiles, n_sm[dev & 15]));
    if (pair) {
        const long long pair_tiles = (long long)(a.N / pl.bn) * ((a.M + 2 * BM - 1) / (2 * BM));
        launch_tc2_pair<64>(ah, al, pl, cm, p, smem, pair_tiles, st);
    } else launch_tc2<64>(ah, al, pl, cm, p, smem, grid, st);
    if (!cache || tmp_hi) {
        UMAB_CUDA(cudaStreamSynchronize(st));
        if (!cache) { cudaFree(pl.hi); cudaFree(pl.lo); }
        if (tmp_hi) { cudaFree(tmp_hi); cudaFree(tmp_lo); }
    }
}

}  // namespace umab
