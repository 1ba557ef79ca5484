// Shared helpers for the tag_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define TAG_OK 0
#define TAG_ERR_BAD_ARG 10001
#define TAG_ERR_UNSUPPORTED 10002

#define TAG_RETURN_IF_LAUNCH_FAILED()                         \
    do {                                                      \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess) return (int)e__;              \
    } while (0)

#define TAG_DTYPE_F32 0
#define TAG_DTYPE_BF16 1

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// round-trip through the storage type (what a later kernel will read back)
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f<T>(from_f<T>(v)); }

// 4-wide vector load/store of storage type T <-> float[4]
template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float* o) {
        float4 v = *reinterpret_cast<const float4*>(p);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
    static __device__ __forceinline__ void store(float* p, const float* o) {
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    }
};
template <> struct Vec4<bf16> {
    static __device__ __forceinline__ void load(const bf16* p, float* o) {
        uint2 v = *reinterpret_cast<const uint2*>(p);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&v.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&v.y);
        float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        o[0] = fa.x; o[1] = fa.y; o[2] = fb.x; o[3] = fb.y;
    }
    static __device__ __forceinline__ void store(bf16* p, const float* o) {
        __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(o[2], o[3]);
        uint2 v;
        v.x = *reinterpret_cast<uint32_t*>(&a);
        v.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = v;
    }
};

// 8-wide vector (16 B of bf16 / 32 B of fp32)
template <typename T> __device__ __forceinline__ void load8(const T* p, float* o) {
    Vec4<T>::load(p, o);
    Vec4<T>::load(p + 4, o + 4);
}
template <> __device__ __forceinline__ void load8<bf16>(const bf16* p, float* o) {   // one LDG.128
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xFFFF0000u);
    o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xFFFF0000u);
    o[4] = __uint_as_float(r.z << 16); o[5] = __uint_as_float(r.z & 0xFFFF0000u);
    o[6] = __uint_as_float(r.w << 16); o[7] = __uint_as_float(r.w & 0xFFFF0000u);
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float* o) {
    Vec4<T>::store(p, o);
    Vec4<T>::store(p + 4, o + 4);
}
template <> __device__ __forceinline__ void store8<bf16>(bf16* p, const float* o) {  // one STG.128
    __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(o[4], o[5]), d = __floats2bfloat162_rn(o[6], o[7]);
    uint4 r;
    r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
    r.z = *reinterpret_cast<uint32_t*>(&c); r.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(p) = r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Counter-based dropout RNG: one 64-bit draw per (seed, group of 4 consecutive elements), 16 bits per
// element.  The same function regenerates the mask in the backward pass, so no mask is ever stored.
__device__ __forceinline__ uint64_t tag_hash64(uint64_t seed, uint64_t idx) {
    uint64_t z = idx * 0x9E3779B97F4A7C15ull + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// 16-byte raw vector access: 8 bf16 or 4 fp32 channels per load
template <typename T> struct Raw16 { static constexpr int VEC = 16 / (int)sizeof(T); static constexpr int NSUB = VEC / 4; };
// unpack 4 consecutive channels (sub-chunk `sub`) of a raw 16-byte vector
template <typename T> __device__ __forceinline__ void unpack4(const uint4& raw, int sub, float* o);
template <> __device__ __forceinline__ void unpack4<float>(const uint4& raw, int, float* o) {
    o[0] = __uint_as_float(raw.x); o[1] = __uint_as_float(raw.y);
    o[2] = __uint_as_float(raw.z); o[3] = __uint_as_float(raw.w);
}
template <> __device__ __forceinline__ void unpack4<bf16>(const uint4& raw, int sub, float* o) {
    const uint32_t lo = sub == 0 ? raw.x : raw.z, hi = sub == 0 ? raw.y : raw.w;
    o[0] = __uint_as_float(lo << 16); o[1] = __uint_as_float(lo & 0xFFFF0000u);
    o[2] = __uint_as_float(hi << 16); o[3] = __uint_as_float(hi & 0xFFFF0000u);
}
template <typename T> __device__ __forceinline__ void pack4(uint4& raw, int sub, const float* o);
template <> __device__ __forceinline__ void pack4<float>(uint4& raw, int, const float* o) {
    raw.x = __float_as_uint(o[0]); raw.y = __float_as_uint(o[1]);
    raw.z = __float_as_uint(o[2]); raw.w = __float_as_uint(o[3]);
}
template <> __device__ __forceinline__ void pack4<bf16>(uint4& raw, int sub, const float* o) {
    __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(o[2], o[3]);
    const uint32_t lo = *reinterpret_cast<uint32_t*>(&a), hi = *reinterpret_cast<uint32_t*>(&b);
    if (sub == 0) { raw.x = lo; raw.y = hi; } else { raw.z = lo; raw.w = hi; }
}
__device__ __forceinline__ uint4 ld16(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st16(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
// 256-bit global accesses (sm_100: LDG / STG .256): one instruction moves a whole 32-byte sector per lane, where two
// 16-byte stores of a strided row pattern reach L2 as two half-filled sector writes
__device__ __forceinline__ void st32(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void ld32(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}

inline void tag_dropout_params(float p, uint32_t* thresh, float* keep_scale) {
    if (p <= 0.f) { *thresh = 0u; *keep_scale = 1.f; return; }
    double t = (double)p * 4294967296.0;
    if (t > 4294967295.0) t = 4294967295.0;
    *thresh = (uint32_t)t;
    *keep_scale = 1.0f / (1.0f - p);
}

// keep-scale: 0 when dropped, 1/(1-p) when kept.  thresh = p * 2^32 (its top 16 bits are compared).
__device__ __forceinline__ float tag_dropout_scale(uint64_t seed, uint64_t idx, uint32_t thresh,
                                                   float keep_scale) {
    const uint64_t z = tag_hash64(seed, idx >> 2);
    const uint32_t u = (uint32_t)(z >> (16 * (idx & 3))) & 0xFFFFu;
    return u >= (thresh >> 16) ? keep_scale : 0.0f;
}
// four consecutive elements starting at idx4 (a multiple of 4): one hash
__device__ __forceinline__ void tag_dropout_scale4(uint64_t seed, uint64_t idx4, uint32_t thresh,
                                                    float keep_scale, float* s4) {
    const uint64_t z = tag_hash64(seed, idx4 >> 2);
    const uint32_t t16 = thresh >> 16;
#pragma unroll
    for (int k = 0; k < 4; ++k) s4[k] = ((uint32_t)(z >> (16 * k)) & 0xFFFFu) >= t16 ? keep_scale : 0.0f;
}
