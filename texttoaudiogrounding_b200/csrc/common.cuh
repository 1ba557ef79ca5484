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
template <typename T> __device__ __forceinline__ void store8(T* p, const float* o) {
    Vec4<T>::store(p, o);
    Vec4<T>::store(p + 4, o + 4);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Counter-based dropout RNG: one 32-bit draw per (seed, element index).  The same
// function regenerates the mask in the backward pass, so no mask is ever stored.
__device__ __forceinline__ uint32_t tag_hash32(uint64_t seed, uint64_t idx) {
    uint64_t z = idx * 0x9E3779B97F4A7C15ull + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}
// keep-scale: 0 when dropped, 1/(1-p) when kept.  thresh = p * 2^32.
__device__ __forceinline__ float tag_dropout_scale(uint64_t seed, uint64_t idx, uint32_t thresh,
                                                   float keep_scale) {
    return tag_hash32(seed, idx) >= thresh ? keep_scale : 0.0f;
}
