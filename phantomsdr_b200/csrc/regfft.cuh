// In-register forward DFTs of 2/4/8/16/32 complex points (sign -1, unnormalised, natural order in
// and out). Every index is a compile-time constant so the arrays live in registers and the
// twiddles become FFMA immediates. Building block of the shared-memory four-step passes in
// fft_fwd.cuh.
#pragma once
#include <cuda_runtime.h>
#include <utility>
#include "w32_table.cuh"

namespace b200 {

template <int I> using IC = std::integral_constant<int, I>;

template <typename F, int... I> __device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>) {
    (f(IC<I>{}), ...);
}
template <int N, typename F> __device__ __forceinline__ void static_for(F &&f) {
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// a * exp(-2*pi*i*M/32), M a compile-time constant
template <int M> __device__ __forceinline__ float2 mul_w32(float2 a) {
    constexpr int m = ((M % 32) + 32) % 32;
    if constexpr (m == 0) {
        return a;
    } else if constexpr (m == 8) {  // -i
        return make_float2(a.y, -a.x);
    } else if constexpr (m == 16) {
        return make_float2(-a.x, -a.y);
    } else if constexpr (m == 24) {  // +i
        return make_float2(-a.y, a.x);
    } else if constexpr (m % 8 == 4) {
        constexpr float h = 0.70710678118654752440f;
        if constexpr (m == 4) return make_float2((a.x + a.y) * h, (a.y - a.x) * h);          // (1-i)/sqrt2
        else if constexpr (m == 12) return make_float2((a.y - a.x) * h, -(a.x + a.y) * h);   // (-1-i)/sqrt2
        else if constexpr (m == 20) return make_float2(-(a.x + a.y) * h, (a.x - a.y) * h);   // (-1+i)/sqrt2
        else return make_float2((a.x - a.y) * h, (a.x + a.y) * h);                            // (1+i)/sqrt2
    } else {
        constexpr float c = W32<m>::c, s = W32<m>::s;  // w = c - i s
        return make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
    }
}

template <int R> struct RegDft;

template <> struct RegDft<1> {
    __device__ __forceinline__ static void run(float2 (&)[1]) {}
};
template <> struct RegDft<2> {
    __device__ __forceinline__ static void run(float2 (&v)[2]) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};
template <> struct RegDft<4> {
    __device__ __forceinline__ static void run(float2 (&v)[4]) {
        float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        float2 t2 = cadd(v[1], v[3]), d = csub(v[1], v[3]);
        float2 t3 = make_float2(d.y, -d.x);  // -i * d
        v[0] = cadd(t0, t2);
        v[1] = cadd(t1, t3);
        v[2] = csub(t0, t2);
        v[3] = csub(t1, t3);
    }
};

// R = 4 * R2:  n = R2*n1 + n2,  k = k1 + 4*k2
template <int R> struct RegDft {
    static_assert(R == 8 || R == 16 || R == 32, "unsupported register DFT size");
    static constexpr int R2 = R / 4;
    __device__ __forceinline__ static void run(float2 (&v)[R]) {
        float2 a[4][R2];
        static_for<R2>([&](auto n2c) {
            constexpr int n2 = decltype(n2c)::value;
            float2 t[4] = {v[n2], v[R2 + n2], v[2 * R2 + n2], v[3 * R2 + n2]};
            RegDft<4>::run(t);
            static_for<4>([&](auto k1c) {
                constexpr int k1 = decltype(k1c)::value;
                a[k1][n2] = mul_w32<n2 * k1 *(32 / R)>(t[k1]);
            });
        });
        static_for<4>([&](auto k1c) {
            constexpr int k1 = decltype(k1c)::value;
            RegDft<R2>::run(a[k1]);
            static_for<R2>([&](auto k2c) {
                constexpr int k2 = decltype(k2c)::value;
                v[k1 + 4 * k2] = a[k1][k2];
            });
        });
    }
};

}  // namespace b200
