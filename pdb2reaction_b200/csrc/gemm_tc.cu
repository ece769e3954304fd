// tcgen05 tensor-core GEMM with fp32-grade accuracy:  C[M,N] = A[M,K] . W[N,K]^T (+bias) (+C)
//
// The SO(2) convolutions, the radial MLPs and their adjoints are the only dense contractions of
// the model.  Parity is fp32-level (1e-5 eV/atom, 1e-4 eV/A), which a single
// bf16 or tf32 pass cannot deliver, so every product is evaluated as a bf16 "x3" split with fp32
// accumulation in TMEM:
//        a = a_hi + a_lo,  w = w_hi + w_lo   (bf16 each, |err| <= 2^-18 |x|)
//        a.w ~= a_hi w_hi + a_hi w_lo + a_lo w_hi        (dropped term ~2^-18)
// Weights are split once on the host side of the launch (cached bf16 planes + TMA tensor maps);
// activations stay fp32 in HBM and are split IN THE KERNEL by the producer warps, which write the
// bf16 hi/lo tiles straight into the 128B-swizzled K-major shared-memory layout tcgen05 expects,
// so no extra HBM pass exists.
//
// Persistent kernel: one CTA per SM loops over 128 x BN output tiles (N fastest, so concurrent CTAs
// share an A tile through L2); 704 threads:
//   warps 0-15  A producers: coalesced LDG.128 of the fp32 tile (2 rows x 256 B per warp instruction, one
//               k block ahead, across tile boundaries), split, st.shared swizzled (8 B per plane),
//               fence.proxy.async, mbarrier arrive
//   warp 16     TMEM allocator + TMA producer of the W_hi / W_lo tiles (cp.async.bulk.tensor)
//   warp 17     single-thread tcgen05.mma issuer (3 MMAs per 16-wide k step), tcgen05.commit
//   warps 18-21 epilogue: tcgen05.ld (lane quadrant warp%4) -> smem transpose -> (+bias,+C) ->
//               fully coalesced STG.128; overlaps the next tile's main loop through the
//               double-buffered TMEM accumulator (2 x tmem_cols columns)
// smem ring of S stages {A_hi, A_lo, W_hi, W_lo}; mbarriers full_a / full_b / empty per stage and
// tmem_full / tmem_empty per accumulator buffer.
#include <cuda.h>
#include <cuda_bf16.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace umab {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // bf16 elements per k block = one 128 B swizzle row
constexpr int PRODUCER_WARPS = 16;
constexpr int PRODUCER_THREADS = PRODUCER_WARPS * 32;
constexpr int ROWS_PER_WARP = 128 / PRODUCER_WARPS;        // tile rows owned by one producer warp
constexpr int LD_PER_KB = ROWS_PER_WARP / 2;               // LDG.128 per thread and k block (2 rows each)
constexpr int TC_THREADS = PRODUCER_THREADS + 2 * 32 + 4 * 32;   // + TMA warp + MMA warp + 4 epilogue warps
constexpr int EPI_PITCH = 36;                       // floats per staged row (16 B aligned, conflict-free)
constexpr int EPI_STAGE_BYTES = 4 * 32 * EPI_PITCH * 4;   // per-warp transpose buffers of the epilogue
constexpr int A_TILE_BYTES = BM * 128; // one bf16 plane of the A tile

struct TcParams {
    const float* A; long long lda;
    float* Cm; long long ldc;
    const float* bias;
    int M, N, K, BN, accumulate, stages, tmem_cols;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// K-major, SWIZZLE_128B, rows packed at 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// split 8 floats into bf16 hi / lo (lo = bf16(x - hi)); returns two 16-byte chunks
__device__ __forceinline__ void split8(float4 v0, float4 v1, uint4& hi, uint4& lo) {
    const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    float r[8];
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
        r[2 * i] = x[2 * i] - __bfloat162float(h0);
        r[2 * i + 1] = x[2 * i + 1] - __bfloat162float(h1);
        __nv_bfloat162 hv; hv.x = h0; hv.y = h1;
        h[i] = *reinterpret_cast<uint32_t*>(&hv);
        l[i] = pack_bf16(r[2 * i], r[2 * i + 1]);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: the dynamic smem base is only guaranteed 16 B aligned -> round up to 1024 B
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int S = p.stages;
    const int BN = p.BN;
    const uint32_t w_tile_bytes = (uint32_t)BN * 128u;
    const uint32_t stage_bytes = 2u * A_TILE_BYTES + 2u * w_tile_bytes;
    const uint32_t epi_base = base + (uint32_t)S * stage_bytes;           // 4 warps x 32 rows x EPI_PITCH floats
    const uint32_t bar_base = epi_base + EPI_STAGE_BYTES;                 // 8 B each
    auto a_hi = [&](int s) { return base + (uint32_t)s * stage_bytes; };
    auto a_lo = [&](int s) { return a_hi(s) + A_TILE_BYTES; };
    auto w_hi = [&](int s) { return a_hi(s) + 2u * A_TILE_BYTES; };
    auto w_lo = [&](int s) { return w_hi(s) + w_tile_bytes; };
    auto full_a = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto full_b = [&](int s) { return bar_base + 8u * (uint32_t)(S + s); };
    auto empty = [&](int s) { return bar_base + 8u * (uint32_t)(2 * S + s); };
    auto tmem_full = [&](int b) { return bar_base + 8u * (uint32_t)(3 * S + b); };
    auto tmem_empty = [&](int b) { return bar_base + 8u * (uint32_t)(3 * S + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(3 * S + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int nkb = p.K / BK;
    const int ntn = p.N / BN;                                   // tiles along N (fastest)
    const int num_tiles = ntn * ((p.M + BM - 1) / BM);
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == PRODUCER_WARPS + 1 && lane == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full_a(s), PRODUCER_THREADS); mbar_init(full_b(s), 1); mbar_init(empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == PRODUCER_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(2 * p.tmem_cols)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < PRODUCER_WARPS) {
        // ===================== A producers: fp32 -> bf16 hi/lo, swizzled K-major tiles.
        // Coalesced mapping: a warp owns ROWS_PER_WARP tile rows; each LDG.128 instruction covers 2 full
        // rows of the k block (16 lanes x 16 B = 256 B per row -> 4 cache lines per instruction).  A lane
        // converts its 4 floats and stores 8 B of the hi and of the lo plane (2 smem wavefronts per
        // 256 B instruction = the minimum).  One flat loop over (tile, k block) with the loads of the
        // next TWO k blocks in flight, so the prefetch runs across tile boundaries.
        const int l16 = lane & 15;
        const int rbase = warp * ROWS_PER_WARP + (lane >> 4);      // + 2*i
        const int total = my_tiles * nkb;
        float4 v[LD_PER_KB], n1[LD_PER_KB], n2[LD_PER_KB];
        auto load_iter = [&](int g, float4* dst) {
            const int lt = g / nkb, kb = g - lt * nkb;
            const int tile = (int)blockIdx.x + lt * (int)gridDim.x;
            const long long mb = (long long)(tile / ntn) * BM;
            const float* src = p.A + (long long)kb * BK + l16 * 4;
#pragma unroll
            for (int i = 0; i < LD_PER_KB; ++i) {
                const long long m = mb + rbase + 2 * i;
                dst[i] = (g < total && m < p.M) ? __ldg(reinterpret_cast<const float4*>(src + m * p.lda)) : f4zero();
            }
        };
        load_iter(0, v);
        load_iter(1, n1);
        for (int g = 0; g < total; ++g) {
            const int s = g % S;
            const uint32_t ph = (uint32_t)(g / S) & 1u;
            load_iter(g + 2, n2);
            mbar_wait(empty(s), ph ^ 1u);
            const uint32_t ah = a_hi(s), al = a_lo(s);
#pragma unroll
            for (int i = 0; i < LD_PER_KB; ++i) {
                const int row = rbase + 2 * i;
                const float4 x = v[i];
                const __nv_bfloat16 h0 = __float2bfloat16_rn(x.x), h1 = __float2bfloat16_rn(x.y),
                                    h2 = __float2bfloat16_rn(x.z), h3 = __float2bfloat16_rn(x.w);
                __nv_bfloat162 ha, hb;
                ha.x = h0; ha.y = h1; hb.x = h2; hb.y = h3;
                const uint32_t l0 = pack_bf16(x.x - __bfloat162float(h0), x.y - __bfloat162float(h1));
                const uint32_t l1 = pack_bf16(x.z - __bfloat162float(h2), x.w - __bfloat162float(h3));
                const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)(l16 >> 1)) ^ (uint32_t)(row & 7)) << 4) +
                                     (uint32_t)(l16 & 1) * 8u;
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(ah + off), "r"(*reinterpret_cast<uint32_t*>(&ha)),
                             "r"(*reinterpret_cast<uint32_t*>(&hb)) : "memory");
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(al + off), "r"(l0), "r"(l1) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(full_a(s));
#pragma unroll
            for (int i = 0; i < LD_PER_KB; ++i) { v[i] = n1[i]; n1[i] = n2[i]; }
        }
    } else if (warp == PRODUCER_WARPS) {
        // ===================== TMA producer of the weight planes
        if (lane == 0) {
            int g = 0;
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int tile = (int)blockIdx.x + lt * (int)gridDim.x;
                const int n0 = (tile % ntn) * BN;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % S;
                    const uint32_t ph = (uint32_t)(g / S) & 1u;
                    mbar_wait(empty(s), ph ^ 1u);
                    mbar_arrive_expect_tx(full_b(s), 2u * w_tile_bytes);
                    tma_load_2d(w_hi(s), &tm_hi, full_b(s), kb * BK, n0);
                    tma_load_2d(w_lo(s), &tm_lo, full_b(s), kb * BK, n0);
                }
            }
        }
    } else if (warp == PRODUCER_WARPS + 1) {
        // ===================== MMA issuer (double-buffered TMEM accumulators)
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N, M=128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int g = 0;
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int ab = lt & 1;
                mbar_wait(tmem_empty(ab), (((uint32_t)lt >> 1) & 1u) ^ 1u);   // epilogue drained this buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(ab * p.tmem_cols);
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % S;
                    const uint32_t ph = (uint32_t)(g / S) & 1u;
                    mbar_wait(full_a(s), ph);
                    mbar_wait(full_b(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t dah = make_smem_desc(a_hi(s)), dal = make_smem_desc(a_lo(s));
                    const uint64_t dwh = make_smem_desc(w_hi(s)), dwl = make_smem_desc(w_lo(s));
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
        // ===================== epilogue warps (the last four; lane quadrant = warp % 4): TMEM -> regs -> smem transpose -> coalesced STG
        const int q = warp & 3;                                   // TMEM lane quadrant this warp may read
        const uint32_t stg = epi_base + (uint32_t)q * (32u * EPI_PITCH * 4u);
        float* stg_ptr = reinterpret_cast<float*>(smem_raw + (stg - smem_u32(smem_raw)));
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;
        for (int lt = 0; lt < my_tiles; ++lt) {
            const int tile = (int)blockIdx.x + lt * (int)gridDim.x;
            const int n0 = (tile % ntn) * BN;
            const long long mrow0 = (long long)(tile / ntn) * BM + q * 32;
            const int ab = lt & 1;
            mbar_wait(tmem_full(ab), ((uint32_t)lt >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + (uint32_t)(ab * p.tmem_cols) + ((uint32_t)(q * 32) << 16);
            for (int ch = 0; ch < BN / 32; ++ch) {
                uint32_t rr[32];
                tmem_ld32(taddr + (uint32_t)(ch * 32), rr);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    st4(stg_ptr + lane * EPI_PITCH + 4 * j,
                        make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]),
                                    __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3])));
                __syncwarp();
                float4 bv = f4zero();
                if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + ch * 32 + c4));
#pragma unroll
                for (int itr = 0; itr < 8; ++itr) {
                    const int row = itr * 4 + rsub;
                    const long long m = mrow0 + row;
                    float4 o = f4add(ld4(stg_ptr + row * EPI_PITCH + c4), bv);
                    if (m < p.M) {
                        float* dst = p.Cm + m * p.ldc + n0 + ch * 32 + c4;
                        if (p.accumulate) o = f4add(o, ld4(dst));
                        st4(dst, o);
                    }
                }
            }
            // all TMEM reads of this buffer are complete (tcgen05.wait::ld inside tmem_ld32)
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(ab));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == PRODUCER_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * p.tmem_cols)) : "memory");
    }
}

// fp32 [rows, K] -> bf16 hi / lo planes
__global__ void split_planes_kernel(const float* __restrict__ w, long long n, __nv_bfloat16* __restrict__ hi,
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

EncodeFn get_encode() {
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

int pick_bn(int N) {
    if (N % 256 == 0) return 256;
    if (N <= 256) return N;
    for (int bn = 256; bn >= 32; bn -= 32)
        if (N % bn == 0) return bn;
    return 0;
}

struct Planes {
    __nv_bfloat16* hi = nullptr; __nv_bfloat16* lo = nullptr;
    CUtensorMap tm_hi, tm_lo;
    int bn = 0;
};

void make_map(CUtensorMap* tm, void* ptr, int N, int K, int bn) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled failed");
}

Planes build_planes(const float* W, int N, int K, cudaStream_t st) {
    Planes p;
    const long long n = (long long)N * K;
    UMAB_CUDA(cudaMalloc(&p.hi, n * 2));
    UMAB_CUDA(cudaMalloc(&p.lo, n * 2));
    split_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, n, p.hi, p.lo);
    UMAB_LAUNCH_CHECK();
    p.bn = pick_bn(N);
    make_map(&p.tm_hi, p.hi, N, K, p.bn);
    make_map(&p.tm_lo, p.lo, N, K, p.bn);
    return p;
}

}  // namespace

// bf16 weight planes + tensor maps, owned by whoever owns the fp32 weights (an engine)
struct TcPlaneCache {
    std::mutex mu;
    std::map<std::tuple<const float*, int, int>, Planes> planes;
    ~TcPlaneCache() { clear(); }
    void clear() {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& kv : planes) { cudaFree(kv.second.hi); cudaFree(kv.second.lo); }
        planes.clear();
    }
};
TcPlaneCache* tc_cache_create() { return new TcPlaneCache(); }
void tc_cache_destroy(TcPlaneCache* c) { delete c; }
void tc_cache_clear(TcPlaneCache* c) { if (c) c->clear(); }

bool gemm_tc_supported(const GemmArgs& a) {
    return a.batch == 1 && a.M >= 1 && a.K % BK == 0 && a.K >= BK && a.N % 32 == 0 && pick_bn(a.N) >= 32 &&
           a.lda % 4 == 0 && a.ldc % 4 == 0 && a.ldw == a.K;
}

// cache == nullptr: planes are built for this call only (unit-test entry)
void gemm_tc(const GemmArgs& a, cudaStream_t st, TcPlaneCache* cache) {
    if (!gemm_tc_supported(a)) throw CudaError("gemm_tc: unsupported shape");
    int dev = 0;
    UMAB_CUDA(cudaGetDevice(&dev));
    Planes pl;
    if (cache) {
        std::lock_guard<std::mutex> lk(cache->mu);
        auto key = std::make_tuple(a.W, a.N, a.K);
        auto it = cache->planes.find(key);
        if (it == cache->planes.end()) it = cache->planes.emplace(key, build_planes(a.W, a.N, a.K, st)).first;
        pl = it->second;
    } else {
        pl = build_planes(a.W, a.N, a.K, st);
    }
    TcParams p;
    p.A = a.A; p.lda = a.lda; p.Cm = a.Cmat; p.ldc = a.ldc; p.bias = a.bias;
    p.M = a.M; p.N = a.N; p.K = a.K; p.BN = pl.bn; p.accumulate = a.accumulate;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * pl.bn * 128;
    p.stages = std::max(2, std::min(4, (200 * 1024) / stage_bytes));
    int cols = 32;
    while (cols < pl.bn) cols *= 2;
    p.tmem_cols = cols;
    const size_t smem = (size_t)p.stages * stage_bytes + EPI_STAGE_BYTES + 1024 /*alignment slack*/ +
                        8 * (3 * p.stages + 5) + 16;
    static std::once_flag attr_once[16];
    std::call_once(attr_once[dev & 15], [] {
        UMAB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    });
    static int n_sm[16] = {0};
    if (!n_sm[dev & 15]) UMAB_CUDA(cudaDeviceGetAttribute(&n_sm[dev & 15], cudaDevAttrMultiProcessorCount, dev));
    const long long tiles = (long long)(a.N / pl.bn) * ((a.M + BM - 1) / BM);
    dim3 grid((unsigned)std::min<long long>(tiles, n_sm[dev & 15]));
    gemm_tc_kernel<<<grid, TC_THREADS, smem, st>>>(pl.tm_hi, pl.tm_lo, p);
    UMAB_LAUNCH_CHECK();
    if (!cache) {
        UMAB_CUDA(cudaStreamSynchronize(st));
        cudaFree(pl.hi);
        cudaFree(pl.lo);
    }
}


}  // namespace umab
