// In-register radix-R complex FFT butterflies with compile-time twiddles (sm_100a).
//
// RegFFT<R, INV>::run(v) replaces v[0..R) by its length-R DFT (forward: exp(-2 pi i nk/R),
// INV: exp(+2 pi i nk/R), unnormalised), natural order in, natural order out.  R is a power of
// two up to 64; composite sizes are split as 4 x (R/4) Cooley-Tukey steps whose indices and
// twiddle factors are all compile-time constants, so the whole butterfly is straight-line code
// on registers with the twiddles as FFMA immediates.
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace bdof {

// ---------------------------------------------------------------- compile-time trigonometry
constexpr double kPiD = 3.141592653589793238462643383279502884;

constexpr double cx_sin_small(double x) {   // |x| <= pi/4, Taylor to ~1e-19
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 12; ++i) { term *= -x2 / double((2 * i) * (2 * i + 1)); sum += term; }
    return sum;
}
constexpr double cx_cos_small(double x) {
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 12; ++i) { term *= -x2 / double((2 * i - 1) * (2 * i)); sum += term; }
    return sum;
}
struct cx_cs { double c, s; };
// cos/sin of 2*pi*m/R, exact at multiples of a quarter turn
constexpr cx_cs cx_cossin(int m, int R) {
    m %= R; if (m < 0) m += R;
    int q = (4 * m) / R;
    int rem = 4 * m - q * R;                     // angle within the quadrant = (pi/2) * rem / R
    double c = 1.0, s = 0.0;
    if (rem != 0) {
        if (2 * rem > R) { double a = kPiD * double(R - rem) / double(2 * R); c = cx_sin_small(a); s = cx_cos_small(a); }
        else             { double a = kPiD * double(rem) / double(2 * R);     c = cx_cos_small(a); s = cx_sin_small(a); }
    }
    switch (q) {
        case 0: return {c, s};
        case 1: return {-s, c};
        case 2: return {-c, -s};
        default: return {s, -c};
    }
}

// ---------------------------------------------------------------- static loop helper
template <int... Is, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F&& f) {
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

// ---------------------------------------------------------------- complex helpers
// Complex values are float2 in an aligned register pair.  Arithmetic uses the sm_100 packed FP32
// instructions (add/mul/fma .f32x2 -> SASS FADD2/FMUL2/FFMA2): one issue slot per complex add, two per
// complex multiply.  ptxas folds the (re,im) swap, per-half negation and scalar broadcast of the
// operands below into operand modifiers (R.F32x2.LO_HI.NP, R.F32, immediates), so multiplying by
// +-i or by a compile-time twiddle needs no extra instructions or registers.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float x, float y) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ float2 upk(u64 r) { float2 c; asm("mov.b64 {%0,%1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(r)); return c; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return upk(add2(pk(a.x, a.y), pk(b.x, b.y))); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return upk(sub2(pk(a.x, a.y), pk(b.x, b.y))); }
// a * b
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return upk(fma2(pk(-a.y, a.x), pk(b.y, b.y), mul2(pk(a.x, a.y), pk(b.x, b.x))));
}
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
    return upk(fma2(pk(a.y, -a.x), pk(b.y, b.y), mul2(pk(a.x, a.y), pk(b.x, b.x))));
}
// conj(a * b)
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {
    return upk(fma2(pk(-a.y, -a.x), pk(b.y, b.y), mul2(pk(a.x, -a.y), pk(b.x, b.x))));
}
__device__ __forceinline__ float2 conjf2(float2 a) { return make_float2(a.x, -a.y); }
// a * (1 + tau) = a + a * tau, a * conj(1 + tau), conj(a * (1 + tau)): products with a transmission given as tau = t - 1
// (common.h); two packed FMAs each, like a plain complex multiply
__device__ __forceinline__ float2 cmul1p(float2 a, float2 tau) {
    return upk(fma2(pk(-a.y, a.x), pk(tau.y, tau.y), fma2(pk(a.x, a.y), pk(tau.x, tau.x), pk(a.x, a.y))));
}
__device__ __forceinline__ float2 cmulc1p(float2 a, float2 tau) {
    return upk(fma2(pk(a.y, -a.x), pk(tau.y, tau.y), fma2(pk(a.x, a.y), pk(tau.x, tau.x), pk(a.x, a.y))));
}
__device__ __forceinline__ float2 cmul_conj1p(float2 a, float2 tau) {
    return upk(fma2(pk(-a.y, -a.x), pk(tau.y, tau.y), fma2(pk(a.x, -a.y), pk(tau.x, tau.x), pk(a.x, -a.y))));
}
// forward: a * (-i); inverse: a * (+i)
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// a * W_R^M, W_R = exp(-2 pi i / R) (forward) or its conjugate (INV); M, R compile-time
template <int M, int R, bool INV>
__device__ __forceinline__ float2 mul_tw(float2 a) {
    constexpr cx_cs w = cx_cossin(M, R);
    constexpr float c = float(w.c);
    constexpr float s = INV ? float(-w.s) : float(w.s);      // multiply by (c - i s)
    if constexpr (w.c == 1.0 && w.s == 0.0) return a;
    else if constexpr (w.c == -1.0 && w.s == 0.0) return make_float2(-a.x, -a.y);
    else if constexpr (w.c == 0.0 && w.s == 1.0) return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
    else if constexpr (w.c == 0.0 && w.s == -1.0) return INV ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
    else return upk(fma2(pk(a.y, -a.x), pk(s, s), mul2(pk(a.x, a.y), pk(c, c))));   // (xc + ys, yc - xs)
}

// ---------------------------------------------------------------- butterflies
template <int R, bool INV> struct RegFFT;

template <bool INV> struct RegFFT<1, INV> {
    static __device__ __forceinline__ void run(float2 (&)[1]) {}
};

template <bool INV> struct RegFFT<2, INV> {
    static __device__ __forceinline__ void run(float2 (&v)[2]) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <bool INV> struct RegFFT<4, INV> {
    static __device__ __forceinline__ void run(float2 (&v)[4]) {
        float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        float2 t2 = cadd(v[1], v[3]), t3 = mul_mi<INV>(csub(v[1], v[3]));
        v[0] = cadd(t0, t2);
        v[2] = csub(t0, t2);
        v[1] = cadd(t1, t3);
        v[3] = csub(t1, t3);
    }
};

// R = A * B with A = (R == 8 ? 2 : 4):  n = B*n1 + n2,  k = k1 + A*k2
//   X[k1 + A k2] = sum_{n2} W_B^{n2 k2} [ W_R^{n2 k1} sum_{n1} W_A^{n1 k1} x[B n1 + n2] ]
template <int R, bool INV> struct RegFFT {
    static constexpr int A = (R == 8) ? 2 : 4;
    static constexpr int B = R / A;
    static __device__ __forceinline__ void run(float2 (&v)[R]) {
        static_for<B>([&](auto N2) __attribute__((always_inline)) {
            constexpr int n2 = decltype(N2)::value;
            float2 t[A];
            static_for<A>([&](auto N1) __attribute__((always_inline)) { constexpr int n1 = decltype(N1)::value; t[n1] = v[B * n1 + n2]; });
            RegFFT<A, INV>::run(t);
            static_for<A>([&](auto K1) __attribute__((always_inline)) {
                constexpr int k1 = decltype(K1)::value;
                v[B * k1 + n2] = mul_tw<(n2 * k1) % R, R, INV>(t[k1]);
            });
        });
        float2 out[R];
        static_for<A>([&](auto K1) __attribute__((always_inline)) {
            constexpr int k1 = decltype(K1)::value;
            float2 u[B];
            static_for<B>([&](auto N2) __attribute__((always_inline)) { constexpr int n2 = decltype(N2)::value; u[n2] = v[B * k1 + n2]; });
            RegFFT<B, INV>::run(u);
            static_for<B>([&](auto K2) __attribute__((always_inline)) { constexpr int k2 = decltype(K2)::value; out[k1 + A * k2] = u[k2]; });
        });
        static_for<R>([&](auto I) __attribute__((always_inline)) { constexpr int i = decltype(I)::value; v[i] = out[i]; });
    }
};

}  // namespace bdof
