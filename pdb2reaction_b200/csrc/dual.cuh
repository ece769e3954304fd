// Dual numbers (value + one tangent) for forward-over-reverse differentiation of the SAME kernels.
//
// Every elementwise / gather / reduce kernel of the pipeline is a template on the scalar type S:
//   S = float : the energy + force path (identical arithmetic to a plain float kernel)
//   S = D1    : value and directional derivative d/d(eps) along a position tangent; running the
//               forward AND the hand-written backward on duals yields d(forces)/d(eps) = -H.v, i.e.
//               one analytic Hessian column per image (reference: torch.autograd.functional.hessian,
//               pdb2reaction/uma_pysis.py:402-409).
// Dual tensors are stored as two separate fp32 arrays (value plane, tangent plane) so the GEMMs --
// linear in their activation operand -- simply run once per plane with the same weights.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace umab {

struct D1 { float v, d; };
struct D4 { float4 v, d; };

template <class S> struct VecOf;
template <> struct VecOf<float> { using type = float4; };
template <> struct VecOf<D1> { using type = D4; };

// ------------------------------------------------------------------ scalar ops
__device__ __forceinline__ float val(float a) { return a; }
__device__ __forceinline__ float val(D1 a) { return a.v; }
template <class S> __device__ __forceinline__ S cst(float c);
template <> __device__ __forceinline__ float cst<float>(float c) { return c; }
template <> __device__ __forceinline__ D1 cst<D1>(float c) { return D1{c, 0.f}; }

__device__ __forceinline__ D1 operator+(D1 a, D1 b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ D1 operator-(D1 a, D1 b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ D1 operator-(D1 a) { return {-a.v, -a.d}; }
__device__ __forceinline__ D1 operator*(D1 a, D1 b) { return {a.v * b.v, fmaf(a.v, b.d, a.d * b.v)}; }
__device__ __forceinline__ D1 operator*(D1 a, float b) { return {a.v * b, a.d * b}; }
__device__ __forceinline__ D1 operator*(float a, D1 b) { return {a * b.v, a * b.d}; }
__device__ __forceinline__ D1 operator+(D1 a, float b) { return {a.v + b, a.d}; }
__device__ __forceinline__ D1 operator+(float a, D1 b) { return {a + b.v, b.d}; }
__device__ __forceinline__ D1 operator-(D1 a, float b) { return {a.v - b, a.d}; }
__device__ __forceinline__ D1 operator-(float a, D1 b) { return {a - b.v, -b.d}; }
__device__ __forceinline__ D1& operator+=(D1& a, D1 b) { a.v += b.v; a.d += b.d; return a; }
__device__ __forceinline__ D1 operator/(D1 a, D1 b) {
    float q = a.v / b.v;
    return {q, (a.d - q * b.d) / b.v};
}
__device__ __forceinline__ D1 operator/(D1 a, float b) { return {a.v / b, a.d / b}; }
__device__ __forceinline__ D1 operator/(float a, D1 b) {
    float q = a / b.v;
    return {q, -q * b.d / b.v};
}

__device__ __forceinline__ float s_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ D1 s_sqrt(D1 a) { float q = sqrtf(a.v); return {q, a.d / (2.0f * q)}; }
__device__ __forceinline__ float s_rsqrt(float a) { return rsqrtf(a); }
__device__ __forceinline__ D1 s_rsqrt(D1 a) { float r = rsqrtf(a.v); return {r, -0.5f * r * r * r * a.d}; }
__device__ __forceinline__ float s_exp(float a) { return expf(a); }
__device__ __forceinline__ D1 s_exp(D1 a) { float e = expf(a.v); return {e, e * a.d}; }
__device__ __forceinline__ float s_sigmoid(float a) { return sigmoidf_(a); }
__device__ __forceinline__ D1 s_sigmoid(D1 a) { float s = sigmoidf_(a.v); return {s, s * (1.f - s) * a.d}; }
__device__ __forceinline__ float s_silu(float a) { return siluf_(a); }
__device__ __forceinline__ D1 s_silu(D1 a) {
    float s = sigmoidf_(a.v);
    return {a.v * s, s * (1.f + a.v * (1.f - s)) * a.d};
}
__device__ __forceinline__ float s_dsilu(float a) { return dsiluf_(a); }
__device__ __forceinline__ D1 s_dsilu(D1 a) {      // silu'(x) and silu''(x) = s(1-s)(2 + x(1-2s))
    float s = sigmoidf_(a.v);
    return {s * (1.f + a.v * (1.f - s)), s * (1.f - s) * (2.f + a.v * (1.f - 2.f * s)) * a.d};
}
__device__ __forceinline__ float s_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ D1 s_fma(D1 a, D1 b, D1 c) { return a * b + c; }

__device__ __forceinline__ D1 warp_sum(D1 a) { return {warp_sum(a.v), warp_sum(a.d)}; }
__device__ __forceinline__ float s_shfl(float a, int lane) { return __shfl_sync(0xffffffffu, a, lane); }
__device__ __forceinline__ D1 s_shfl(D1 a, int lane) {
    return {__shfl_sync(0xffffffffu, a.v, lane), __shfl_sync(0xffffffffu, a.d, lane)};
}
__device__ __forceinline__ float s_shfl_xor(float a, int m) { return __shfl_xor_sync(0xffffffffu, a, m); }
__device__ __forceinline__ D1 s_shfl_xor(D1 a, int m) {
    return {__shfl_xor_sync(0xffffffffu, a.v, m), __shfl_xor_sync(0xffffffffu, a.d, m)};
}

// ------------------------------------------------------------------ 4-vector ops
template <class V> __device__ __forceinline__ V vzero();
template <> __device__ __forceinline__ float4 vzero<float4>() { return f4zero(); }
template <> __device__ __forceinline__ D4 vzero<D4>() { return D4{f4zero(), f4zero()}; }

__device__ __forceinline__ float4 vadd(float4 a, float4 b) { return f4add(a, b); }
__device__ __forceinline__ D4 vadd(D4 a, D4 b) { return {f4add(a.v, b.v), f4add(a.d, b.d)}; }
__device__ __forceinline__ float4 vsub(float4 a, float4 b) { return f4sub(a, b); }
__device__ __forceinline__ D4 vsub(D4 a, D4 b) { return {f4sub(a.v, b.v), f4sub(a.d, b.d)}; }
__device__ __forceinline__ float4 vmul(float4 a, float4 b) { return f4mul(a, b); }
__device__ __forceinline__ D4 vmul(D4 a, D4 b) { return {f4mul(a.v, b.v), f4add(f4mul(a.v, b.d), f4mul(a.d, b.v))}; }
__device__ __forceinline__ float4 vscale(float4 a, float s) { return f4scale(a, s); }
__device__ __forceinline__ D4 vscale(D4 a, D1 s) { return {f4scale(a.v, s.v), f4add(f4scale(a.d, s.v), f4scale(a.v, s.d))}; }
__device__ __forceinline__ D4 vscale(D4 a, float s) { return {f4scale(a.v, s), f4scale(a.d, s)}; }
__device__ __forceinline__ float4 vneg(float4 a) { return f4scale(a, -1.f); }
__device__ __forceinline__ D4 vneg(D4 a) { return {f4scale(a.v, -1.f), f4scale(a.d, -1.f)}; }
// acc += s * b
__device__ __forceinline__ void vfma(float4& acc, float s, float4 b) { f4fma(acc, s, b); }
__device__ __forceinline__ void vfma(D4& acc, D1 s, D4 b) {
    f4fma(acc.v, s.v, b.v);
    f4fma(acc.d, s.v, b.d);
    f4fma(acc.d, s.d, b.v);
}
// Product that must ROUND AS A PRODUCT even when the next operation adds it to something: ptxas contracts a packed
// mul.rn.f32x2 with a following add.rn.f32x2 into FFMA2 (the .rn of the packed forms does not stop it, an empty asm
// between them does not either -- the contraction happens in ptxas, seen in SASS as 18 -> 16 FMUL2 / FADD2), and it
// does so in some copies of a loop body and not in others.  The scalar mul.rn.f32 is never contracted (PTX ISA).  Used
// where one kernel STORES the product and another ACCUMULATES it and both must give the same bits (the l = 0 row of
// the gather adjoint: open-chunk kernel + source_reduce against the closed-chunk half kernels).
__device__ __forceinline__ float4 vmul_unfused(float4 a, float4 b) {
    return make_float4(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z), __fmul_rn(a.w, b.w));
}
// acc += a * b,  acc -= a * b  (elementwise; packed FFMA2)
__device__ __forceinline__ void f4fmav(float4& acc, float4 a, float4 b) {
    const float2 lo = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(acc.x, acc.y));
    const float2 hi = __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), make_float2(acc.z, acc.w));
    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void vfmav(float4& acc, float4 a, float4 b) { f4fmav(acc, a, b); }
__device__ __forceinline__ D4 vmul_unfused(D4 a, D4 b) {
    float4 t = vmul_unfused(a.d, b.v);
    f4fmav(t, a.v, b.d);
    return D4{vmul_unfused(a.v, b.v), t};
}
__device__ __forceinline__ void vfmav(D4& acc, D4 a, D4 b) {
    f4fmav(acc.v, a.v, b.v);
    f4fmav(acc.d, a.v, b.d);
    f4fmav(acc.d, a.d, b.v);
}
__device__ __forceinline__ void vfnmav(float4& acc, float4 a, float4 b) {
    f4fmav(acc, make_float4(-a.x, -a.y, -a.z, -a.w), b);
}
__device__ __forceinline__ void vfnmav(D4& acc, D4 a, D4 b) {
    const float4 na = make_float4(-a.v.x, -a.v.y, -a.v.z, -a.v.w), nd = make_float4(-a.d.x, -a.d.y, -a.d.z, -a.d.w);
    f4fmav(acc.v, na, b.v);
    f4fmav(acc.d, na, b.d);
    f4fmav(acc.d, nd, b.v);
}
__device__ __forceinline__ float vdot(float4 a, float4 b) { return f4dot(a, b); }
__device__ __forceinline__ D1 vdot(D4 a, D4 b) { return {f4dot(a.v, b.v), f4dot(a.v, b.d) + f4dot(a.d, b.v)}; }
__device__ __forceinline__ float vhsum(float4 a) { return f4hsum(a); }
__device__ __forceinline__ D1 vhsum(D4 a) { return {f4hsum(a.v), f4hsum(a.d)}; }
// a - s (broadcast scalar)
__device__ __forceinline__ float4 vsubs(float4 a, float s) { return make_float4(a.x - s, a.y - s, a.z - s, a.w - s); }
__device__ __forceinline__ D4 vsubs(D4 a, D1 s) { return {vsubs(a.v, s.v), vsubs(a.d, s.d)}; }

#define UMAB_MAP4(name, fn)                                                                     \
    __device__ __forceinline__ float4 name(float4 a) { return make_float4(fn(a.x), fn(a.y), fn(a.z), fn(a.w)); } \
    __device__ __forceinline__ D4 name(D4 a) {                                                  \
        D1 x = fn(D1{a.v.x, a.d.x}), y = fn(D1{a.v.y, a.d.y}), z = fn(D1{a.v.z, a.d.z}), w = fn(D1{a.v.w, a.d.w}); \
        return {make_float4(x.v, y.v, z.v, w.v), make_float4(x.d, y.d, z.d, w.d)};              \
    }
UMAB_MAP4(vsilu, s_silu)
UMAB_MAP4(vsigmoid, s_sigmoid)
UMAB_MAP4(vdsilu, s_dsilu)
#undef UMAB_MAP4

// L2 prefetch of the 16 bytes at p (a warp issuing it with consecutive lanes pulls 512 B = 4 lines): overlaps the
// DRAM latency of data a later phase / loop iteration will load, without holding destination registers
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ------------------------------------------------------------------ global pointers (value plane [+ tangent plane])
template <class S> struct GP;
template <> struct GP<float> {
    float* p;
    __host__ __device__ GP operator+(long long o) const { return GP{p + o}; }
    __host__ __device__ explicit operator bool() const { return p != nullptr; }
    __device__ __forceinline__ float4 ld4(long long i) const { return *reinterpret_cast<const float4*>(p + i); }
    // read-only path: only for tensors the kernel never writes (not for in-place operands)
    __device__ __forceinline__ float4 ldg4(long long i) const { return __ldg(reinterpret_cast<const float4*>(p + i)); }
    __device__ __forceinline__ float ldg(long long i) const { return __ldg(p + i); }
    __device__ __forceinline__ void prefetch(long long i) const { prefetch_l2(p + i); }
    __device__ __forceinline__ void st4(long long i, float4 x) const { *reinterpret_cast<float4*>(p + i) = x; }
    __device__ __forceinline__ float ld(long long i) const { return p[i]; }
    __device__ __forceinline__ void st(long long i, float x) const { p[i] = x; }
    __device__ __forceinline__ GP vback(long long) const { return *this; }
};
template <> struct GP<D1> {
    float* v;
    float* d;
    __host__ __device__ GP operator+(long long o) const { return GP{v + o, d + o}; }
    __host__ __device__ explicit operator bool() const { return v != nullptr; }
    __device__ __forceinline__ D4 ld4(long long i) const {
        return D4{*reinterpret_cast<const float4*>(v + i), *reinterpret_cast<const float4*>(d + i)};
    }
    __device__ __forceinline__ D4 ldg4(long long i) const {
        return D4{__ldg(reinterpret_cast<const float4*>(v + i)), __ldg(reinterpret_cast<const float4*>(d + i))};
    }
    __device__ __forceinline__ D1 ldg(long long i) const { return D1{__ldg(v + i), __ldg(d + i)}; }
    __device__ __forceinline__ void prefetch(long long i) const { prefetch_l2(v + i); prefetch_l2(d + i); }
    __device__ __forceinline__ void st4(long long i, D4 x) const {
        *reinterpret_cast<float4*>(v + i) = x.v;
        *reinterpret_cast<float4*>(d + i) = x.d;
    }
    __device__ __forceinline__ D1 ld(long long i) const { return D1{v[i], d[i]}; }
    __device__ __forceinline__ void st(long long i, D1 x) const { v[i] = x.v; d[i] = x.d; }
    // value plane stepped back by o elements (ImgShare: read the value of the FIRST image of the chunk); read-only use
    __device__ __forceinline__ GP vback(long long o) const { return GP{v - o, d}; }
};

// ------------------------------------------------------------------ Hessian columns of one base geometry
// All images of the launch hold the same VALUE planes (only the tangents differ), and the tensors that come out of the
// value-plane GEMMs exist for the first image of the chunk only (engine: dedupe, share).  A kernel launched with
// n_img > 1 then (a) orders its work so that consecutive CTAs take the SAME rows of consecutive images -- they read the
// same value rows at the same time, once from HBM and otherwise from L2 -- and (b) reads those tensors through
// GP::vback(img * rows * width).  n_img <= 1: the plain mapping.  `rows` = rows (edges or nodes) per image.
struct ImgShare { int n_img; int rows; };
// row of this warp within the launch (false: nothing to do) and the image it belongs to; wpc warps per CTA
template <class S>
__device__ __forceinline__ bool share_row(const ImgShare& sh, int wpc, long long n_rows, long long& row, int& img) {
    const int w = threadIdx.x >> 5;
    if (std::is_same<S, float>::value || sh.n_img <= 1) {      // the float kernels keep exactly their plain mapping
        row = (long long)blockIdx.x * wpc + w;
        img = 0;
        return row < n_rows;
    }
    const int tile = blockIdx.x / sh.n_img;
    img = blockIdx.x - tile * sh.n_img;
    const int r = tile * wpc + w;
    row = (long long)img * sh.rows + r;
    return r < sh.rows;
}
inline unsigned share_grid(const ImgShare& sh, int wpc, long long n_rows) {
    if (sh.n_img <= 1) return (unsigned)((n_rows + wpc - 1) / wpc);
    return (unsigned)(((long long)sh.rows + wpc - 1) / wpc) * (unsigned)sh.n_img;
}
// ------------------------------------------------------------------ A-operand outputs
// An activation that is ONLY consumed as the A operand of a GEMM is written either as plain fp32 (exact SIMT
// GEMMs, small systems) or directly in the operand format of the tensor-core GEMM: two bf16 planes hi / lo with
// x ~= hi + lo (gemm_tc2.cu) -- the same 4 bytes per element, so the split costs no HBM traffic.
// Element offsets are identical in both formats.
template <class S> struct AP;
template <> struct AP<float> {
    float* p;
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
    __host__ __device__ AP operator+(long long o) const {
        return AP{p ? p + o : nullptr, hi ? hi + o : nullptr, lo ? lo + o : nullptr};
    }
    // format known at compile time (kernels instantiated per format: no uniform branch per store, and the address
    // arithmetic of neighbouring stores folds into immediates)
    template <bool PL> __device__ __forceinline__ void st4t(long long i, float4 x) const {
        if constexpr (PL) {
            uint2 h, l;
            split4(x, h, l);
            *reinterpret_cast<uint2*>(hi + i) = h;
            *reinterpret_cast<uint2*>(lo + i) = l;
        } else {
            *reinterpret_cast<float4*>(p + i) = x;
        }
    }
    __host__ __device__ bool planes() const { return hi != nullptr; }
    __device__ __forceinline__ void st4(long long i, float4 x) const {
        if (hi) {
            // (pairing lanes for one 16-byte store per lane instead of two 8-byte ones was measured: no gain)
            uint2 h, l;
            split4(x, h, l);
            *reinterpret_cast<uint2*>(hi + i) = h;
            *reinterpret_cast<uint2*>(lo + i) = l;
        } else {
            *reinterpret_cast<float4*>(p + i) = x;
        }
    }
};
template <> struct AP<D1> {
    AP<float> v, d;
    // Value elements at offsets >= vlim are NOT stored.  Hessian columns of one base geometry (engine: dedupe): the
    // value-plane GEMM reads the rows of the FIRST image of the chunk only, so the value plane of an A operand is a
    // dead store for every other image (the tangent plane is always written).  Default: store everything.
    long long vlim = 0x7fffffffffffffffLL;
    __host__ __device__ AP operator+(long long o) const {
        return AP{v + o, d + o, vlim == 0x7fffffffffffffffLL ? vlim : vlim - o};
    }
    __device__ __forceinline__ void st4(long long i, D4 x) const {
        if (i < vlim) v.st4(i, x.v);
        d.st4(i, x.d);
    }
    template <bool PL> __device__ __forceinline__ void st4t(long long i, D4 x) const {
        if (i < vlim) v.template st4t<PL>(i, x.v);
        d.template st4t<PL>(i, x.d);
    }
    __host__ __device__ bool planes() const { return v.hi != nullptr; }
};

// occupancy hint: the float instantiations keep the register budgets of the hand-tuned float kernels
#ifndef UMAB_D1_MINB
#define UMAB_D1_MINB 1
#endif
template <class S> constexpr int min_blocks(int for_float) { return std::is_same<S, float>::value ? for_float : UMAB_D1_MINB; }

inline GP<float> gpf(const float* p) { return GP<float>{const_cast<float*>(p)}; }

}  // namespace umab
