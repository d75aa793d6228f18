// Shared device helpers for libmslesseg (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mslesseg.h"

namespace msl {

void set_error(const char* fmt, ...);

#define MSL_CUDA_CHECK(expr)                                                        \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            msl::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));          \
            return MSL_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define MSL_LAUNCH_CHECK(name)                                                      \
    do {                                                                            \
        cudaError_t _e = cudaGetLastError();                                        \
        if (_e != cudaSuccess) {                                                    \
            msl::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
            return MSL_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

constexpr unsigned FULL = 0xffffffffu;

// Packed FP32 pairs (Blackwell FADD2 / FMUL2 / FFMA2): two IEEE operations per issue slot, each half rounded exactly like
// the scalar instruction.  CAUTION: ptxas contracts a mul2 that feeds an add2 into one FFMA2 even under -fmad=false (one
// rounding instead of two).  Where the reference rounds the product and the sum separately, keep the sum scalar or make
// sure (cuobjdump -sass) that the pattern did not fuse.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 pk2(float2 v) { return pk2(v.x, v.y); }
__device__ __forceinline__ float2 upk2(f32x2 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ f32x2 mul2_rn(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2_rn(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2_rn(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2_rn(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// Order-preserving float <-> uint32 key (so unsigned atomicMin/atomicMax order floats).
__device__ __forceinline__ unsigned f2key(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
    unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// E1 (reference utils/utils.py:400-405): float32 sub, then div, then mul by 255, truncation.
// Explicit _rn intrinsics: no FMA contraction, IEEE division.
__device__ __forceinline__ uint8_t normalise_px(float f, float mn, float p) {
    float g = __fsub_rn(f, mn);
    if (p > 0.0f) g = __fmul_rn(255.0f, __fdiv_rn(g, p));
    return (uint8_t)__float2int_rz(g);
}

// cv::saturate_cast<uchar>(float): cvRound (round-half-even) then clamp.
__device__ __forceinline__ uint8_t sat_u8_rn(float v) {
    int i = __float2int_rn(v);
    return (uint8_t)min(max(i, 0), 255);
}

template <typename T> __device__ __forceinline__ float load_as_float(const T* p);
template <> __device__ __forceinline__ float load_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_as_float<uint8_t>(const uint8_t* p) { return (float)__ldg(p); }

}  // namespace msl
