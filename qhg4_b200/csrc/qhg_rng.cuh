// qhg_rng.cuh -- counter-based per-agent random streams and bit-reproducible double math.
//
// The reference draws every random number from one WELL512 generator per OpenMP thread
// (utils/WELL512.cpp:70-86, core/SPopulation.cpp:168-176), which makes results depend on the
// thread count.  Here every draw is Philox4x32-10(counter = {agent id lo, agent id hi, step,
// stream}, key = seed), so a draw depends only on WHO draws WHEN, never on where the agent is
// stored or which thread processes it.  Like the reference's wrandd() (utils/WELL512.h:33) a
// real-valued draw is a 32-bit integer scaled by 2^-32.
//
// All double arithmetic that feeds a decision uses the explicitly rounded intrinsics
// (__dadd_rn/__dmul_rn/__ddiv_rn, never contracted to FMA) so the CPU oracle
// (oracle/qhg_oracle.cpp, built with -ffp-contract=off) reproduces it bit for bit.
#pragma once
#include <cstdint>

namespace qhg {

// draw streams (counter word 3) and lanes
enum : uint32_t { STREAM_ACT0 = 0, STREAM_ACT1 = 1, STREAM_PAIR = 2, STREAM_BABY = 3 };
enum { L0_DEATH = 0, L0_MOVE = 1, L0_BIRTH = 2, L0_DEATH2 = 3 };   // lanes of STREAM_ACT0
enum { L1_MOVE2 = 0, L1_NAV = 1, L1_SIGDEATH = 2, L1_OLDAGE = 3 };  // lanes of STREAM_ACT1 (bridges: streams 0x04000000|b/4)

struct RngKey { uint32_t k0, k1; };

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, RngKey key) {
    uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// the same function with the ten round keys (warp-uniform) computed once per kernel, and the two 32x32->64
// products of a round written so that each becomes one IMAD.WIDE
struct RoundKeys { uint32_t a[10], b[10]; };
__device__ __forceinline__ RoundKeys round_keys(RngKey key) {
    RoundKeys K;
#pragma unroll
    for (int r = 0; r < 10; r++) { K.a[r] = key.k0 + (uint32_t)r * 0x9E3779B9u; K.b[r] = key.k1 + (uint32_t)r * 0xBB67AE85u; }
    return K;
}
__device__ __forceinline__ uint4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const RoundKeys &K) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.a[r], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.b[r];
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint4 agent_draws_rk(int64_t id, uint32_t step, uint32_t stream, const RoundKeys &K) {
    return philox4x32_10_rk((uint32_t)((uint64_t)id & 0xffffffffu), (uint32_t)((uint64_t)id >> 32), step, stream, K);
}

// u2d(u) < p  <=>  u < ceil(p * 2^32): the double comparison of the reference as an exact integer threshold
// (p * 2^32 is exact; p <= 0 or NaN gives 0 = never, p >= 1 gives >= 2^32 = always)
__device__ __forceinline__ unsigned long long prob_threshold(double p) {
    return (p > 0.0) ? __double2ull_ru(__dmul_rn(p, 4294967296.0)) : 0ull;
}

__device__ __forceinline__ uint4 agent_draws(int64_t id, uint32_t step, uint32_t stream, RngKey key) {
    return philox4x32_10((uint32_t)((uint64_t)id & 0xffffffffu), (uint32_t)((uint64_t)id >> 32), step, stream, key);
}

// wrandd(): x / 2^32, exact in double
__device__ __forceinline__ double u2d(uint32_t x) { return __dmul_rn((double)x, 2.3283064365386962890625e-10); }
// wrandr(a,b): a + ((b-a)*x)/2^32   (utils/WELL512.h:36)
__device__ __forceinline__ double u2range(uint32_t x, double a, double b) {
    return __dadd_rn(a, __dmul_rn(__dmul_rn(__dadd_rn(b, -a), (double)x), 2.3283064365386962890625e-10));
}
// wrandi(a,b): a + (uint)((1.0*(b-a)*x)/2^32)   (utils/WELL512.h:39, s = 1)
__device__ __forceinline__ uint32_t u2int(uint32_t x, uint32_t a, uint32_t b) {
    return a + (uint32_t)__dmul_rn(__dmul_rn((double)(b - a), (double)x), 2.3283064365386962890625e-10);
}

// wrandi(a,b,s): multiples of s, a + s*(uint)((1.0*r*x)/(s*2^32)) with r = b-a rounded up to a multiple of s (s = 1 or 2: the
// divisor is a power of two, the division exact)
__device__ __forceinline__ uint32_t u2int_s(uint32_t x, uint32_t a, uint32_t b, uint32_t s) {
    uint32_t r = b - a;
    r += r % s;
    return a + s * (uint32_t)__dmul_rn(__dmul_rn((double)r, (double)x), (s == 2) ? 1.16415321826934814453125e-10 : 2.3283064365386962890625e-10);
}

// atan in double, argument reduction + odd polynomial (the classic fdlibm scheme), written with
// explicitly rounded operations so that the oracle's counter mode (same operation order on the CPU)
// gives the identical bits.  The resulting p(age) is within 1e-6 relative of the reference's libm value
// (tests/test_parity_gpu.py::test_deterministic_substeps_vs_reference) and equal to the golden curve generated from the
// reference through the oracle (tests/test_oracle_golden.py).
__device__ __forceinline__ double atan_rn(double x) {
    const double aT[11] = {3.33333333333329318027e-01, -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                           -1.11111104054623557880e-01, 9.09088713343650656196e-02, -7.69187620504482999495e-02,
                           6.66107313738753120669e-02, -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                           -3.65315727442169155270e-02, 1.62858201153657823623e-02};
    const double hi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01, 1.57079632679489655800e+00};
    const double lo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17, 6.12323399573676603587e-17};
    const bool neg = x < 0;
    const double ax = fabs(x);
    int idx;
    double xx;
    if (ax >= 73786976294838206464.0) {  // 2^66
        double r = __dadd_rn(hi[3], lo[3]);
        return neg ? -r : r;
    }
    if (ax < 0.4375) {
        if (ax < 7.450580596923828125e-9) return x;  // 2^-27
        idx = -1; xx = ax;
    } else if (ax < 1.1875) {
        if (ax < 0.6875) { idx = 0; xx = __ddiv_rn(__dadd_rn(__dmul_rn(2.0, ax), -1.0), __dadd_rn(2.0, ax)); }
        else             { idx = 1; xx = __ddiv_rn(__dadd_rn(ax, -1.0), __dadd_rn(ax, 1.0)); }
    } else {
        if (ax < 2.4375) { idx = 2; xx = __ddiv_rn(__dadd_rn(ax, -1.5), __dadd_rn(1.0, __dmul_rn(1.5, ax))); }
        else             { idx = 3; xx = __ddiv_rn(-1.0, ax); }
    }
    const double z = __dmul_rn(xx, xx), w = __dmul_rn(z, z);
    double s1 = aT[10];
    s1 = __dadd_rn(aT[8], __dmul_rn(w, s1));
    s1 = __dadd_rn(aT[6], __dmul_rn(w, s1));
    s1 = __dadd_rn(aT[4], __dmul_rn(w, s1));
    s1 = __dadd_rn(aT[2], __dmul_rn(w, s1));
    s1 = __dadd_rn(aT[0], __dmul_rn(w, s1));
    s1 = __dmul_rn(z, s1);
    double s2 = aT[9];
    s2 = __dadd_rn(aT[7], __dmul_rn(w, s2));
    s2 = __dadd_rn(aT[5], __dmul_rn(w, s2));
    s2 = __dadd_rn(aT[3], __dmul_rn(w, s2));
    s2 = __dadd_rn(aT[1], __dmul_rn(w, s2));
    s2 = __dmul_rn(w, s2);
    double r;
    if (idx < 0) {
        r = __dadd_rn(xx, -__dmul_rn(xx, __dadd_rn(s1, s2)));
    } else {
        double hh = hi[0], ll = lo[0];
        if (idx == 1) { hh = hi[1]; ll = lo[1]; }
        else if (idx == 2) { hh = hi[2]; ll = lo[2]; }
        else if (idx == 3) { hh = hi[3]; ll = lo[3]; }
        r = __dadd_rn(hh, -__dadd_rn(__dadd_rn(__dmul_rn(xx, __dadd_rn(s1, s2)), -ll), -xx));
    }
    return neg ? -r : r;
}

// exp in double for the Miami NPP model (core/NPPCalcMiami.cpp:28-42): reduction by ln2 in two pieces + degree-5
// polynomial in r^2 (fdlibm scheme), explicitly rounded, identical operation order in oracle/qhg_oracle.cpp::exp_portable.
// Valid for |x| < 700 (the model's arguments stay within [-100, 100]).
__device__ __forceinline__ double exp_rn(double x) {
    const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    if (fabs(x) < 3.725290298461914e-9) return __dadd_rn(1.0, x);  // 2^-28
    const int k = (int)__dadd_rn(__dmul_rn(invln2, x), (x < 0 ? -0.5 : 0.5));  // truncation toward zero, as the C cast
    const double t = (double)k;
    const double hi = __dadd_rn(x, -__dmul_rn(t, ln2HI)), lo = __dmul_rn(t, ln2LO);
    const double rr = __dadd_rn(hi, -lo);
    const double tt = __dmul_rn(rr, rr);
    double pp = P5;
    pp = __dadd_rn(P4, __dmul_rn(tt, pp));
    pp = __dadd_rn(P3, __dmul_rn(tt, pp));
    pp = __dadd_rn(P2, __dmul_rn(tt, pp));
    pp = __dadd_rn(P1, __dmul_rn(tt, pp));
    const double c = __dadd_rn(rr, -__dmul_rn(tt, pp));
    if (k == 0) return __dadd_rn(1.0, -__dadd_rn(__ddiv_rn(__dmul_rn(rr, c), __dadd_rn(c, -2.0)), -rr));
    const double y = __dadd_rn(1.0, -__dadd_rn(__dadd_rn(lo, -__ddiv_rn(__dmul_rn(rr, c), __dadd_rn(2.0, -c))), -hi));
    return ldexp(y, k);
}

}  // namespace qhg
