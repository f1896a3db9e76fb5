// oracle/qhg_oracle.cpp -- CPU restatement of QHG4's per-step agent update.
//
// TEST INFRASTRUCTURE ONLY (see qhg_oracle.h).  Single-threaded, plain C++; every function
// cites the reference file:line it follows (paths relative to /root/reference/QHG4/).
// Parity status: PINNED.  QOR_MODE_WELL is checked bit-exactly against the reference itself
// (oracle/_ref, one OpenMP thread) in tests/test_oracle_vs_ref.py and against golden vectors
// generated from it (tests/golden/, tests/make_golden.py).
#include "qhg_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <numbers>
#include <queue>
#include <set>
#include <string>
#include <vector>

namespace {

// life states, core/SPopulation.h:70-74
constexpr uint32_t LIFE_DEAD = 0, LIFE_ALIVE = 1, LIFE_FERTILE = 5, LIFE_MOVING = 8;
// event ids, utils_qhg/EventConsts.h:14-36
constexpr int EVENT_ID_GEO = 2, EVENT_ID_FLUSH = 20;
constexpr double PI = std::numbers::pi;   // utils/qhg_consts.h:71
constexpr double ATAN_EPS = 0.001;        // actions/ATanDeath.h:14

// ---------------------------------------------------------------------------------------------
// WELL512 (utils/WELL512.cpp:70-86; Lomont's public-domain formulation)
struct Well512 {
    uint32_t s[16];
    uint32_t idx = 0;
    void seed(const uint32_t *st) { memcpy(s, st, sizeof(s)); idx = 0; }
    uint32_t next() {
        uint32_t a = s[idx];
        uint32_t c = s[(idx + 13) & 15];
        uint32_t b = a ^ c ^ (a << 16) ^ (c << 15);
        c = s[(idx + 9) & 15];
        c ^= (c >> 11);
        a = s[idx] = b ^ c;
        uint32_t d = a ^ ((a << 5) & 0xDA442D24u);
        idx = (idx + 15) & 15;
        a = s[idx];
        s[idx] = a ^ b ^ d ^ (a << 2) ^ (b << 18) ^ (c << 28);
        return s[idx];
    }
};

// utils/WELL512.h:33-39: every real-valued draw is a 32-bit integer scaled by 2^-32
inline double u2d(uint32_t x) { return (1.0 * x) / 4294967296.0; }
inline double u2range(uint32_t x, double a, double b) { return a + ((b - a) * x) / 4294967296.0; }
inline uint32_t u2int(uint32_t x, uint32_t a, uint32_t b) {  // wrandi(a,b) with s=1
    uint32_t r = b - a;
    return a + (uint32_t)((1.0 * r * x) / 4294967296.0);
}
inline uint32_t u2int_s(uint32_t x, uint32_t a, uint32_t b, uint32_t s) {  // wrandi(a,b,s), utils/WELL512.h:39: multiples of s
    uint32_t r = b - a;
    r += (r % s);
    return a + s * (uint32_t)((1.0 * r * x) / (s * 4294967296.0));
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011, public algorithm) -- the counter-based generator of the CUDA path
inline void philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// atan used by the counter mode: argument reduction + odd polynomial (the classic fdlibm scheme), evaluated in
// exactly the operation order of qhg4_b200/csrc/qhg_rng.cuh::atan_rn (this file is built with
// -ffp-contract=off, the device code with explicitly rounded intrinsics), so both give identical bits.
// WELL mode keeps libm's atan like the reference.
inline double atan_portable(double x) {
    static const double aT[11] = {3.33333333333329318027e-01, -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                                  -1.11111104054623557880e-01, 9.09088713343650656196e-02, -7.69187620504482999495e-02,
                                  6.66107313738753120669e-02, -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                                  -3.65315727442169155270e-02, 1.62858201153657823623e-02};
    static const double hi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01, 1.57079632679489655800e+00};
    static const double lo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17, 6.12323399573676603587e-17};
    const bool neg = x < 0;
    const double ax = std::fabs(x);
    int idx;
    double xx;
    if (ax >= 73786976294838206464.0) { double r = hi[3] + lo[3]; return neg ? -r : r; }
    if (ax < 0.4375) {
        if (ax < 7.450580596923828125e-9) return x;
        idx = -1; xx = ax;
    } else if (ax < 1.1875) {
        if (ax < 0.6875) { idx = 0; xx = (2.0 * ax + -1.0) / (2.0 + ax); }
        else             { idx = 1; xx = (ax + -1.0) / (ax + 1.0); }
    } else {
        if (ax < 2.4375) { idx = 2; xx = (ax + -1.5) / (1.0 + 1.5 * ax); }
        else             { idx = 3; xx = -1.0 / ax; }
    }
    const double z = xx * xx, w = z * z;
    double s1 = aT[10];
    s1 = aT[8] + w * s1; s1 = aT[6] + w * s1; s1 = aT[4] + w * s1; s1 = aT[2] + w * s1; s1 = aT[0] + w * s1;
    s1 = z * s1;
    double s2 = aT[9];
    s2 = aT[7] + w * s2; s2 = aT[5] + w * s2; s2 = aT[3] + w * s2; s2 = aT[1] + w * s2;
    s2 = w * s2;
    double r;
    if (idx < 0) r = xx + -(xx * (s1 + s2));
    else r = hi[idx] + -(((xx * (s1 + s2)) + -lo[idx]) + -xx);
    return neg ? -r : r;
}

// exp used by the counter mode (Miami NPP model), same scheme: reduction by ln2 in two pieces + degree-5 polynomial
// in r^2, operation order identical to qhg_rng.cuh::exp_rn.  Arguments outside [-700, 700] do not occur here.
inline double exp_portable(double x) {
    const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    if (std::fabs(x) < 3.725290298461914e-9) return 1.0 + x;  // 2^-28
    const int k = (int)(invln2 * x + (x < 0 ? -0.5 : 0.5));
    const double t = (double)k;
    const double hi = x + -(t * ln2HI), lo = t * ln2LO;
    const double r = hi + -lo;
    const double tt = r * r;
    double pp = P5;
    pp = P4 + tt * pp; pp = P3 + tt * pp; pp = P2 + tt * pp; pp = P1 + tt * pp;
    const double c = r + -(tt * pp);
    if (k == 0) return 1.0 + -(((r * c) / (c + -2.0)) + -r);
    const double y = 1.0 + -((lo + -((r * c) / (2.0 + -c))) + -hi);
    return std::ldexp(y, k);
}

// draw streams of the counter mode (mirrored in qhg4_b200/csrc/qhg_rng.cuh)
enum { STREAM_ACT0 = 0, STREAM_ACT1 = 1, STREAM_PAIR = 2, STREAM_BABY = 3 };
// lanes of STREAM_ACT0 / STREAM_ACT1
enum { L0_DEATH = 0, L0_MOVE = 1, L0_BIRTH = 2, L0_DEATH2 = 3 };
enum { L1_MOVE2 = 0, L1_NAV = 1, L1_SIGDEATH = 2, L1_OLDAGE = 3 };  // (bridges draw from their own streams 0x04000000|b/4)

// ---------------------------------------------------------------------------------------------
// PolyLine (utils/PolyLine.cpp:60-89 getVal, :92-127 readFromString)
struct PolyLine {
    std::vector<double> x, v, a;
    unsigned nseg = 0;
    bool parse(const char *def) {
        std::vector<double> d;
        const char *p = def;
        char *e;
        while (true) {
            while (*p == ' ' || *p == '\t') p++;
            if (!*p) break;
            double val = strtod(p, &e);
            if (e == p) return false;
            d.push_back(val);
            p = e;
        }
        if (d.size() < 4 || (d.size() % 2)) return false;
        size_t n = d.size() / 2;
        x.resize(n); v.resize(n); a.assign(n, 0.0);
        for (size_t i = 0; i < n; i++) { x[i] = d[2 * i]; v[i] = d[2 * i + 1]; }
        nseg = (unsigned)n - 1;
        for (unsigned i = 0; i < nseg; i++) a[i] = (v[i + 1] - v[i]) / (x[i + 1] - x[i]);  // utils/PolyLine.cpp addPoint
        return true;
    }
    double val(double fx) const {
        if (nseg == 0) return fx;
        if (fx >= x[nseg]) return v[nseg];
        unsigned i = 0;
        while (i <= nseg && fx > x[i]) i++;
        if (i == 0) return v[0];
        if (i <= nseg) return v[i - 1] + a[i - 1] * (fx - x[i - 1]);
        return v[nseg];
    }
};

// ---------------------------------------------------------------------------------------------
// genome primitives for 1-bit nucleotides (genes/BitGeneUtils.cpp), generic over the source of 32-bit draws so that
// WELL mode (draws = consecutive words of one WELL512, the reference's call order) and counter mode share them.
typedef std::function<uint32_t()> NextU32;

// genes/BitGeneUtils.cpp:86-106.  The break list is used in the order given (the reference never sorts it,
// SURVEY.md §7 "crossover break lists per block are not sorted although makeMultiMask assumes sorted").
inline uint64_t makeMultiMask(const std::vector<uint32_t> &br) {
    uint32_t k = 1 - (uint32_t)(br.size() % 2);
    uint64_t out = 0;
    int i = (int)br.size() - 1;
    for (uint32_t j = 0; j < 64; j++) {
        if (i >= 0 && j == 64 - br[i]) { i--; k = 1 - k; }
        out = (out << 1) + k;
    }
    return out;
}

// genes/BitGeneUtils.cpp:116-186: out (2 strands x nBlocks) = in with its two strands crossed over at nCross random bits
// With bpn = 2 (genes/GeneUtils.cpp:184-245) the breaks fall on nucleotide boundaries: wrandi(0, nBits, 2).
inline void bitCrossOver(uint64_t *out, const uint64_t *in, int G, int nCross, const NextU32 &next, int bpn = 1) {
    const uint32_t nBlocks = (uint32_t)((G * bpn + 63) / 64), nBits = nBlocks * 64;
    std::map<uint32_t, std::vector<uint32_t>> blockBreaks;
    for (int i = 0; i < nCross; i++) {
        uint32_t pos = u2int_s(next(), 0, nBits, (uint32_t)bpn);  // wrandi(0, nBits, BITSINNUC)
        blockBreaks[pos / 64].push_back(pos % 64);
    }
    uint32_t last = 0, S = 0;
    int cur = 0;
    for (auto &kv : blockBreaks) {
        const uint32_t b = kv.first;
        cur = S % 2;
        for (uint32_t q = last; q < b; q++) { out[q] = in[cur * nBlocks + q]; out[nBlocks + q] = in[(1 - cur) * nBlocks + q]; }
        const uint64_t L = makeMultiMask(kv.second), R = ~L;
        out[b] = (L & in[cur * nBlocks + b]) | (R & in[(1 - cur) * nBlocks + b]);
        out[b + nBlocks] = (L & in[(1 - cur) * nBlocks + b]) | (R & in[cur * nBlocks + b]);
        S += (uint32_t)kv.second.size();
        last = b + 1;
    }
    if (nBlocks > last) {
        cur = S % 2;
        for (uint32_t q = last; q < nBlocks; q++) { out[q] = in[cur * nBlocks + q]; out[nBlocks + q] = in[(1 - cur) * nBlocks + q]; }
    }
}

// genes/BitGeneUtils.cpp:190-220: every block gets an independent random 64-bit mask (two 32-bit draws, high word first)
// With bpn = 2 (genes/GeneUtils.cpp:322-362) every other bit of the mask is doubled, so that nucleotides stay whole.
inline void bitFreeReco(uint64_t *out, const uint64_t *in, int nBlocks, const NextU32 &next, int bpn = 1) {
    for (int b = 0; b < nBlocks; b++) {
        const uint64_t hi = next(), lo = next();
        uint64_t L = (hi << 32) + lo;
        if (bpn == 2) { L &= 0x5555555555555555ull; L += L << 1; }
        const uint64_t R = ~L;
        out[b] = (L & in[b]) | (R & in[nBlocks + b]);
        out[b + nBlocks] = (L & in[nBlocks + b]) | (R & in[b]);
    }
}

// genes/BitGeneUtils.cpp:57-75: flip nMut random bits of the first nBits bits
inline void bitMutate(uint64_t *g, int nBits, int nMut, const NextU32 &next) {
    for (int i = 0; i < nMut; i++) {
        uint32_t pos = u2int(next(), 0, (uint32_t)nBits);
        g[pos / 64] ^= (uint64_t)1 << (pos % 64);
    }
}
// genes/GeneUtils.cpp:112-146: nMut times, a random nucleotide among the first nNucs is XORed with 01, 10 or 11
inline void nucMutate(uint64_t *g, int nNucs, int nMut, const NextU32 &next) {
    const uint32_t nBits = 2u * (uint32_t)nNucs;
    for (int i = 0; i < nMut; i++) {
        const uint32_t pos = u2int_s(next(), 0, nBits, 2);
        const uint64_t mask = u2int(next(), 1, 4);
        g[pos / 64] ^= mask << (pos % 64);
    }
}

// utils/bino_tools.cpp (the classic log-gamma series and continued fraction of the incomplete beta function)
inline double gammaln_(double xx) {
    static const double co[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091, -1.231739572450155, 0.1208650973866179e-2, -0.5395239384953e-5};
    double ser = 1.000000000190015, x = xx, y = xx + 1, tmp = x + 5.5;
    tmp -= (x + 0.5) * log(tmp);
    for (int k = 0; k <= 5; k++) { ser += co[k] / y; y++; }
    return -tmp + log(2.5066282746310005 * ser / x);
}
inline double betacf_(double a, double b, double x) {
    const double EPSB = 3.0e-7; const float FMIN = 1.0e-30f;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0, c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < FMIN) d = FMIN;
    d = 1.0 / d;
    double h = d;
    int m;
    for (m = 1; m <= 100; m++) {
        int m2 = 2 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (fabs(d) < FMIN) d = FMIN;
        c = 1.0 + aa / c; if (fabs(c) < FMIN) c = FMIN;
        d = 1.0 / d; h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (fabs(d) < FMIN) d = FMIN;
        c = 1.0 + aa / c; if (fabs(c) < FMIN) c = FMIN;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < EPSB) break;
    }
    if (m > 100) h = -1;
    return h;
}
inline double ibeta_(double a, double b, double x) {
    if (x < 0.0 || x > 1.0) return -1;
    double bt = (x == 0.0 || x == 1.0) ? 0.0 : exp(gammaln_(a + b) - gammaln_(a) - gammaln_(b) + a * log(x) + b * log(1.0 - x));
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf_(a, b, x) / a;
    return 1.0 - bt * betacf_(b, a, 1.0 - x) / b;
}
// utils/BinomialDist.cpp:61-87: table[k] = 1 - I_p(k+1, n-k) until the tail drops below eps; getN(r) = first k with r <= table[k]
inline std::vector<double> binomialTable(double prob, int n, double eps) {
    std::vector<double> v;
    int k = 1;
    double d2 = ibeta_(k, n - k + 1, prob);
    while (d2 > eps && k < n) { v.push_back(d2); k++; d2 = ibeta_(k, n - k + 1, prob); }
    v.push_back(d2);
    for (auto &x : v) x = 1 - x;
    return v;
}
inline int binomialGetN(const std::vector<double> &t, double r) {  // utils/BinomialDist.cpp:92-103
    if (!(r >= 0)) return -1;
    size_t i = 0;
    while (i < t.size() && r > t[i]) i++;
    return (int)i;
}

// ---------------------------------------------------------------------------------------------
struct Agent {  // core/SPopulation.h:43-50 + populations/tut_EnvironAltPop.h:16-21
    uint32_t life = 0;
    int32_t cell = -1;
    int64_t id = -1;
    float birth = 0;
    uint8_t gender = 0;
    float age = 0;
    float lastBirth = 0;
    int32_t mate = -3;
    int32_t numBabies = 0;          // populations/OoANavGenPop.h:21-27
    std::vector<uint64_t> genome;   // 2 strands x nBlocks (actions/Genetics.cpp keeps them in a LayerArrBuf beside the agents)
};

enum ActKind { A_GETOLD, A_ATANDEATH, A_WEIGHTEDMOVE, A_SINGLEEVAL, A_FERTILITY, A_RANDOMPAIR, A_VERHULST, A_OLDAGEDEATH,
               A_VERHULSTVARK, A_MULTIEVAL, A_NPPCAP, A_GENETICS, A_NAVIGATE, A_RANDOMMOVE, A_CONFINEDMOVE, A_WEIGHTEDMOVERAND, A_SIGDEATH,
               A_CONDWEIGHTEDMOVE, A_RANDPERMPAIR, A_MOVESTATS};

// one SingleEvaluator inside a MultiEvaluator (actions/SingleEvaluator.cpp:138-167)
struct SubEval {
    std::string input;      // environment array it evaluates ("Altitude") or "" for the capacities array
    std::string weightName; // name of its combination weight attribute
    bool usePoly = false;
    int trigger = 0;        // event id that makes it recompute (EVENT_ID_GEO / EVENT_ID_VEG)
    PolyLine poly;
    std::string polyName;
    bool first = true, needUpdate = false;
    double weight = 0;
    bool cumulate = true;   // the bCumulate argument of its constructor (actions/SingleEvaluator.cpp:216-243)
};

struct Action {
    std::string name;
    ActKind kind;
    int prio = -1;       // -1: no priority assigned -> never initialized/executed (core/Prioritizer.cpp:19-30)
    bool enabled = true;
};

}  // namespace

struct qor_pop {
    std::string popClass;
    int nCells = 0, maxNeigh = 6, mode = QOR_MODE_WELL;
    std::vector<int32_t> nbr, gid;
    std::vector<uint8_t> nNbr;
    std::map<std::string, std::vector<double>> env;  // "Altitude", "Ice", ...
    std::map<std::string, std::vector<double>> envDelta;  // AutoInterpolator::m_mDiff: per-step differences of the interpolated targets
    std::vector<Action> actions;

    // parameters (names as in the reference's XML / QDF attributes)
    double atanMaxAge = 0, atanRange = 0, atanSlope = 0, atanScale = 0;
    double oadMaxAge = 0, oadUncertainty = 0;
    double moveProb = 0;
    double randMoveProb = 0;  // RandomMove_prob (actions/RandomMove.cpp)
    double moveRandProb = 0;  // WeightedMoveRand_prob (actions/WeightedMoveRand.cpp)
    double sigMaxAge = 0, sigRange = 0, sigSlope = 0, sigScale = 0;  // SigDeath (actions/SigDeath.cpp)
    // CondWeightedMove (actions/CondWeightedMove.cpp) with a SimpleCondition over the altitudes (actions/SimpleCondition.cpp):
    // the mode is a constructor argument in the reference; the probe classes carry it in their name (tut_EnvironAltCond<m>Pop)
    double condMoveProb = 0;
    int condMode = 0;
    // MoveStats (actions/MoveStats.cpp): per cell, hops / distance / time of the arrival that counts under MoveStats_Mode
    // (0 first, 1 minimum, 2 last); the Temp arrays are the per-thread arrays of the reference (one thread), never reset
    int msMode = -1;
    std::vector<int> msHops, msHopsTemp;
    std::vector<double> msDist, msTime, msDistTemp, msTimeTemp;
    // ConfinedMove (actions/ConfinedMove.cpp): centre (lon, lat in degrees) and radius (km) of the region agents may enter
    double confX = 0, confY = 0, confR = 0;
    std::vector<uint8_t> confAllowed;
    // tut_ParthenoPop: no pairing action; LinearBirth tests m_iMateIndex >= 0 (actions/LinearBirth.cpp:139-142), which holds for
    // every agent there: newborns get their own index (populations/tut_ParthenoPop.cpp:107-119), agents that were read in keep
    // the zero of LayerBuf's fresh memory.  All newborns are female, whatever the gender draw said (:116)
    bool selfMate = false;
    float fertMinAge = 0, fertMaxAge = 0, fertInterbirth = 0;
    double vB0 = -1024, vD0 = -1024, vTheta = -1024, vK = -1024;
    PolyLine altPref;
    bool havePoly = false;
    // NPPCapacity (actions/NPPCapacity.cpp) + MultiEvaluator (actions/MultiEvaluator.cpp) of tut_EnvironCapAltPop
    double nppWater = 0, nppCoastal = 0, nppCoastMinLat = 0, nppCoastMaxLat = 0, nppMin = 0, nppMax = 0, nppKMax = 0, nppKMin = 0, nppEff = 1;
    bool nppNeedUpdate = true;
    std::vector<double> cap;      // m_adCapacities
    std::vector<SubEval> subs;    // evaluators of the MultiEvaluator, in construction order
    bool multiFirst = true, multiObserves = false;
    int multiMode = 0;                  // MultiEvalModes (actions/MultiEvaluator.h:19-26): ADD_SIMPLE, ADD_BLOCK, MUL_SIMPLE, MAX_SIMPLE, MAX_BLOCK, MIN_SIMPLE
    std::vector<double> singleW;        // m_adSingleEvalWeights: ONE scratch array shared by all evaluators, never cleared between calls
    std::vector<uint8_t> multiAllowed;  // m_acAllowed (findBlockings)
    // Genetics<.., BitGeneUtils> (actions/Genetics.cpp)
    int genomeSize = 0, numCrossOvers = 0, nBlocks = 0;
    int bitsPerNuc = 1;     // 1: Genetics<.., BitGeneUtils>, 2: Genetics<.., GeneUtils> (OoANavGen2bitPop ...)
    Well512 genWell;        // WELL mode: the Genetics action's OWN generator (actions/Genetics.cpp:91-123), thread 0
    bool haveGenWell = false;
    double mutationRate = 0;
    std::vector<double> binoTable;
    // Navigate (actions/Navigate.cpp) over the Navigation group (core/Navigation.h:13-37)
    std::map<int, std::map<int, double>> navDest;          // port cell -> (destination cell -> distance)
    std::vector<std::pair<int, int>> navBridges, curBridges;
    std::map<int, std::pair<int, std::vector<std::pair<int, double>>>> jumpProbs;  // port -> (n, [(dest, cumulated prob)])
    double navDecay = 0, navDist0 = 0, navProb0 = 0, navBridgeProb = 0;
    bool navNeedUpdate = true;

    // evaluator state (actions/SingleEvaluator.cpp:138-167,332-346)
    bool evalFirst = true, evalNeedUpdate = false;
    bool evaluatorObserves = false;  // does the population addObserver() its evaluator?

    // storage: slots with holes, lowest-free-index allocation (utils/LBController.cpp:249-269 + utils/L2List.cpp:355-370:
    // the PASSIVE list is kept in index order, so getFreeIndex returns the lowest free slot)
    std::vector<Agent> slots;
    std::vector<uint8_t> active;  // slot is linked in the ACTIVE list (includes last step's dead until they are unlinked)
    std::priority_queue<int, std::vector<int>, std::greater<int>> freeHoles;
    std::vector<int> prevDead, deathList, birthList, moveList;  // core/SPopulation.cpp:160-181 queues (one thread)
    int64_t numUsed = 0;       // LBController::m_iNumUsed
    int64_t maxID = 0, nextID = 0;
    int64_t birthIdOffset = 0, birthIdTotal = -1;  // sharded runs: births of the lower ranks / of all ranks this step
    std::vector<uint64_t> counts;
    std::vector<double> B, D, W;
    float curTime = -1;
    uint32_t stepIndex = 0;
    std::vector<unsigned> levelsDone;
    uint64_t totBirths = 0, totDeaths = 0, totMoves = 0, stepBirths = 0, stepDeaths = 0, stepMoves = 0;

    // rng
    uint32_t state16[16];
    Well512 well;
    uint32_t key[2] = {0, 0};

    // ---- helpers -------------------------------------------------------------------------
    Action *find(const std::string &n) {
        for (auto &a : actions) if (a.name == n) return &a;
        return nullptr;
    }
    int hi() const { return (int)slots.size(); }
    void draw4(int64_t id, uint32_t stream, uint32_t out[4]) const {
        uint32_t ctr[4] = {(uint32_t)((uint64_t)id & 0xffffffffu), (uint32_t)((uint64_t)id >> 32), stepIndex, stream};
        philox(ctr, key, out);
    }
    // one 32-bit draw: sequential WELL stream, or lane `lane` of the agent's counter stream
    uint32_t draw(int64_t id, uint32_t stream, int lane) {
        if (mode == QOR_MODE_WELL) return well.next();
        uint32_t o[4];
        draw4(id, stream, o);
        return o[lane];
    }
    int allocSlot() {  // LBController::getFreeIndex, utils/LBController.cpp:249-269
        int i;
        if (!freeHoles.empty()) { i = freeHoles.top(); freeHoles.pop(); }
        else { i = hi(); slots.emplace_back(); active.push_back(0); }
        active[i] = 1;
        numUsed++;
        return i;
    }
    void freeSlot(int i) {  // LBController::deleteElement, utils/LBController.cpp:276-290
        active[i] = 0;
        freeHoles.push(i);
        numUsed--;
    }

    // core/SPopulation.cpp:926-938
    void registerDeath(int i) { slots[i].life = LIFE_DEAD; deathList.push_back(i); }
    // core/SPopulation.cpp:732-740
    void registerBirth(int cell, int mother, int father) { birthList.push_back(cell); birthList.push_back(mother); birthList.push_back(father); }
    // core/SPopulation.cpp:1035-1050
    void registerMove(int from, int i, int to) { slots[i].life |= LIFE_MOVING; moveList.push_back(from); moveList.push_back(i); moveList.push_back(to); }

    // core/SPopulation.cpp:509-536
    void updateNumAgentsPerCell() {
        std::fill(counts.begin(), counts.end(), 0);
        for (int i = 0; i < hi(); i++) if (active[i] && slots[i].life > 0) counts[slots[i].cell]++;
    }

    // ---- initialize() of each action ---------------------------------------------------------
    // actions/LinearBirth.cpp:90-115 and actions/LinearDeath.cpp:92-123 (constant K)
    void verhulstInit() {
        for (int c = 0; c < nCells; c++) {
            B[c] = vB0 + (vTheta - vB0) * ((double)counts[c] / vK);
            D[c] = vD0 + (vTheta - vD0) * ((double)counts[c] / vK);
        }
    }
    // actions/VerhulstVarK.cpp:69-87 -> LinearBirth.cpp:97-112 / LinearDeath.cpp:101-119 with a per-cell K
    void verhulstVarKInit() {
        for (int c = 0; c < nCells; c++) {
            if (cap[c] <= 0) { B[c] = 0; D[c] = 1; }
            else {
                B[c] = vB0 + (vTheta - vB0) * ((double)counts[c] / cap[c]);
                D[c] = vD0 + (vTheta - vD0) * ((double)counts[c] / cap[c]);
            }
        }
    }
    // actions/NPPCapacity.cpp:138-217 (recalculate) + core/NPPCalcMiami.cpp:28-42 (calcNPP)
    void nppRecalculate() {
        if (!nppNeedUpdate) return;
        const std::vector<double> &T = env["AnnualMeanTemp"], &P = env["AnnualRainfall"], &Wt = env["Water"], &npp = env["BaseNPP"],
                                  &alt = env["Altitude"], &lon = env["Longitude"], &lat = env["Latitude"], &coast = env["Coastal"];
        for (int i = 0; i < nCells; i++) {
            const double eT = (mode == QOR_MODE_WELL) ? exp(1.315 - 0.119 * T[i]) : exp_portable(1.315 - 0.119 * T[i]);
            const double eP = (mode == QOR_MODE_WELL) ? exp(-0.000664 * P[i]) : exp_portable(-0.000664 * P[i]);
            const double nppT = 3000.0 / (1 + eT);
            const double nppP = 3000.0 * (1 - eP);
            cap[i] = 0.000475 * ((nppT < nppP) ? nppT : nppP);  // GDM_TO_KGC, core/NPPCalcMiami.cpp:12
        }
        for (int i = 0; i < nCells; i++) {
            double tc = -1;
            if (alt[i] > 0) {
                const float A0 = 1500, A1 = 2500;
                double af = (alt[i] < A0) ? 1 : ((A1 - alt[i]) / (A1 - A0));
                if (af < 0) af = 0;
                double tn = npp[i];
                if (lon[i] > 115.0 && lat[i] > -12.0 && lon[i] < 150.0 && lat[i] < 1.0) {  // Oceania box, actions/NPPCapacity.cpp:22-25
                    if (npp[i] < nppMin) tn = cap[i];
                }
                tn *= af;
                if (tn < nppMin) tc = nppKMin;
                else if (tn > nppMax) tc = nppKMax;
                else tc = nppKMin + tn * nppKMax / (nppMax - nppMin);
                if (coast[i] != 0 && lat[i] > nppCoastMinLat && lat[i] < nppCoastMaxLat) tc += nppCoastal * nppKMax;
                tc += Wt[i] * nppWater * nppKMax;
                if (tc > nppKMax) tc = nppKMax;
            } else {
                tc = 0;
            }
            cap[i] = tc * nppEff;
        }
        nppNeedUpdate = false;
    }
    // SingleEvaluator::calcValues + exchangeAndCumulate (bCumulate = true) into a scratch array
    void subEvalCompute(const SubEval &e, std::vector<double> &out) {
        const int stride = maxNeigh + 1;
        const std::vector<double> &in = e.input.empty() ? cap : env[e.input];
        const std::vector<double> *ice = env.count("Ice") ? &env["Ice"] : nullptr;
        for (int c = 0; c < nCells; c++) {
            double v = 0;
            if (!ice || (*ice)[c] == 0) {
                double dv = e.usePoly ? (e.polyName.empty() ? altPref : e.poly).val((float)in[c]) : in[c];
                v = (dv > 0) ? dv : 0;
            }
            out[(size_t)c * stride] = v;
        }
        for (int c = 0; c < nCells; c++) {
            double w = out[(size_t)c * stride];
            for (int k = 0; k < maxNeigh; k++) {
                int n = nbr[(size_t)c * maxNeigh + k];
                double cw = (n >= 0) ? out[(size_t)n * stride] : 0;
                cw = (cw > 0) ? cw : 0;
                w = e.cumulate ? w + cw : cw;
                out[(size_t)c * stride + k + 1] = w;
            }
        }
    }
    // SingleEvaluator::initialize (actions/SingleEvaluator.cpp:138-167): the shared scratch array is only rewritten when the
    // evaluator needs an update (or has never run); otherwise it keeps whatever the previous user left in it
    void subEvalInit(SubEval &e) {
        if (e.needUpdate || e.first) {
            e.first = false;
            std::fill(singleW.begin(), singleW.end(), 0.0);  // calcValues starts with a memset (:175)
            subEvalCompute(e, singleW);
        }
    }
    // MultiEvaluator::findBlockings (actions/MultiEvaluator.cpp:579-598): an entry is blocked if ANY evaluator is <= 0 there.
    // No memset before the evaluators' initialize: an evaluator that does not recompute is judged by the array as it stands.
    void findBlockings() {
        multiAllowed.assign(W.size(), 1);
        for (auto &e : subs) {
            subEvalInit(e);
            for (size_t i = 0; i < W.size(); i++) if (singleW[i] <= 0) multiAllowed[i] = 0;
        }
    }
    // MultiEvaluator::initialize and the six combine modes (actions/MultiEvaluator.cpp:142-182, 221-253 ADD_SIMPLE, 263-298
    // ADD_BLOCK, 308-342 MUL_SIMPLE, 350-381 MAX_SIMPLE (the only one that does not cumulate the rows), 391-432 MAX_BLOCK,
    // 440-478 MIN_SIMPLE).  An evaluator that needs no update contributes the zeros of the memset.
    void multiEvalInit() {
        bool need = false;
        for (auto &e : subs) need |= e.needUpdate;
        if (!(need || multiFirst)) return;
        multiFirst = false;
        if (singleW.size() != W.size()) singleW.assign(W.size(), 0.0);
        const int stride = maxNeigh + 1;
        const bool block = multiMode == 1 || multiMode == 4;
        if (block) findBlockings();
        const double inf = std::numeric_limits<double>::infinity();
        const double init = (multiMode == 2) ? 1.0 : (multiMode == 3 || multiMode == 4) ? -inf : (multiMode == 5) ? inf : 0.0;
        std::fill(W.begin(), W.end(), init);
        for (auto &e : subs) {
            std::fill(singleW.begin(), singleW.end(), 0.0);
            subEvalInit(e);
            for (size_t i = 0; i < W.size(); i++) {
                if (block && !multiAllowed[i]) continue;
                const double v = singleW[i] * e.weight;
                switch (multiMode) {
                case 0: case 1: W[i] += v; break;
                case 2: W[i] *= v; break;
                case 3: case 4: if (W[i] < v) W[i] = v; break;
                case 5: if (W[i] > v) W[i] = v; break;
                }
            }
        }
        if (multiMode != 3)
            for (int c = 0; c < nCells; c++)
                for (int k = 1; k < stride; k++) W[(size_t)c * stride + k] += W[(size_t)c * stride + k - 1];
    }
    // actions/SingleEvaluator.cpp:174-207 (calcValues) and :216-243 (exchangeAndCumulate), bCumulate = true
    void evaluatorCompute() {
        const int stride = maxNeigh + 1;
        const std::vector<double> &in = env["Altitude"];
        const std::vector<double> *ice = env.count("Ice") ? &env["Ice"] : nullptr;
        std::fill(W.begin(), W.end(), 0.0);
        for (int c = 0; c < nCells; c++) {
            if (!ice || (*ice)[c] == 0) {
                double dv = havePoly ? altPref.val((float)in[c]) : in[c];
                W[(size_t)c * stride] = (dv > 0) ? dv : 0;
            }
        }
        for (int c = 0; c < nCells; c++) {
            double w = W[(size_t)c * stride];
            for (int k = 0; k < maxNeigh; k++) {
                double cw = 0;
                int n = nbr[(size_t)c * maxNeigh + k];
                if (n >= 0) cw = W[(size_t)n * stride];
                cw = (cw > 0) ? cw : 0;
                w = w + cw;
                W[(size_t)c * stride + k + 1] = w;
            }
        }
    }
    void evaluatorInit() {
        if (evalNeedUpdate || evalFirst) { evalFirst = false; evaluatorCompute(); }
    }
    // actions/Navigate.cpp:94-144
    int navRecalculate() {
        if (!navNeedUpdate) return 0;
        const double A = navProb0 / exp(navDecay * navDist0);
        jumpProbs.clear();
        curBridges.clear();
        for (auto &pt : navDest) {
            std::vector<std::pair<int, double>> tp(pt.second.size() + 1);
            double sum = 0;
            int i = 1;
            for (auto &dd : pt.second) {
                double pr = A * exp(navDecay * dd.second);
                sum += pr;
                tp[i++] = {dd.first, pr};
            }
            if (!(sum < 1)) return -1;
            tp[0] = {-1, 1 - sum};
            for (size_t j = 1; j < tp.size(); j++) tp[j].second += tp[j - 1].second;
            jumpProbs[pt.first] = {(int)pt.second.size(), tp};
        }
        const std::vector<double> &alt = env["Altitude"];
        for (auto &b : navBridges) if (alt[b.first] > 0 && alt[b.second] > 0) curBridges.push_back(b);
        navNeedUpdate = false;
        return 0;
    }

    // actions/RandomPair.cpp:107-128 (initialize) and :146-279 (findMates)
    void randomPairInit() {
        for (int i = 0; i < hi(); i++) if (active[i]) slots[i].mate = -3;
        std::vector<std::vector<int>> F(nCells), M(nCells);
        for (int i = 0; i < hi(); i++) {
            if (!active[i]) continue;
            const Agent &a = slots[i];
            if (a.life > 0 && a.life == LIFE_FERTILE) {
                if (a.gender == 0) F[a.cell].push_back(i);
                else if (a.gender == 1) M[a.cell].push_back(i);
            }
        }
        for (int c = 0; c < nCells; c++) {
            int nf = (int)F[c].size(), nm = (int)M[c].size();
            if (nf == 0 || nm == 0) continue;
            if (mode == QOR_MODE_WELL) {
                // slot-sorted buckets (already ascending), the smaller sex picks uniformly among the unpaired of the other
                std::vector<int> &S = (nf <= nm) ? F[c] : M[c];   // iterated in order
                std::vector<int> &L = (nf <= nm) ? M[c] : F[c];   // chosen from
                int nl = (int)L.size();
                std::vector<uint8_t> taken(nl, 0);
                for (int s = 0; s < (int)S.size() && nl > 0; s++) {
                    int choose = (int)u2range(well.next(), 0, nl);
                    int seen = -1, j = -1;
                    while (seen < choose) { j++; if (!taken[j]) seen++; }
                    taken[j] = 1;
                    slots[S[s]].mate = L[j];
                    slots[L[j]].mate = S[s];
                    nl--;
                }
            } else {
                // counter mode: rank both sexes by (random key, id); equal ranks pair up -> a uniformly random
                // injection of the smaller sex into the larger one, the same law as the sequential picking above
                auto ranked = [&](std::vector<int> &v) {
                    std::vector<std::pair<std::pair<uint32_t, int64_t>, int>> k;
                    for (int i : v) { uint32_t o[4]; draw4(slots[i].id, STREAM_PAIR, o); k.push_back({{o[0], slots[i].id}, i}); }
                    std::sort(k.begin(), k.end());
                    for (size_t j = 0; j < v.size(); j++) v[j] = k[j].second;
                };
                ranked(F[c]); ranked(M[c]);
                int np = std::min(nf, nm);
                for (int r = 0; r < np; r++) { slots[F[c][r]].mate = M[c][r]; slots[M[c][r]].mate = F[c][r]; }
            }
        }
    }

    // actions/SimpleCondition.cpp:12-19,73-77: the candidate's value enters scaled by 0.2 (and "less" also asks for < 10)
    bool condAllow(int cur, int to) {
        const std::vector<double> &ref = env["Altitude"];
        const double c = ref[cur], n = 0.2 * ref[to];
        switch (condMode) {
        case 1: return true;
        case 2: return n > c;
        case 3: return (n < 10) && (n < c);
        case 4: return n == c;
        case 5: return n >= c;
        case 6: return n <= c;
        case 7: return n != c;
        default: return false;
        }
    }

    // actions/RandPermPair.cpp:67-83 (initialize) and :99-178 (findMates), :187-196 (permute): the fertile agents of a cell in
    // slot order; the larger sex is shuffled (a partial Fisher-Yates over its first min(nF, nM) places), equal places mate.
    // Counter mode: a uniformly random injection of the smaller sex into the larger one is exactly what the rank-by-key
    // pairing of RandomPair produces there -- the two actions share the law, so they share the code.
    void randPermPairInit() {
        if (mode != QOR_MODE_WELL) { randomPairInit(); return; }
        for (int i = 0; i < hi(); i++) slots[i].mate = -3;
        std::vector<std::vector<int>> F(nCells), M(nCells);
        for (int i = 0; i < hi(); i++) {
            if (!active[i]) continue;
            const Agent &a = slots[i];
            if (a.life > 0 && a.life == LIFE_FERTILE) {
                if (a.gender == 0) F[a.cell].push_back(i);
                else if (a.gender == 1) M[a.cell].push_back(i);
            }
        }
        auto permute = [&](std::vector<int> &v, int nSel) {
            for (int i = 0; i < nSel; i++) {
                const int k = (int)u2int(well.next(), i, (int)v.size());
                std::swap(v[k], v[i]);
            }
        };
        for (int c = 0; c < nCells; c++) {
            if (counts[c] <= 1) continue;
            const int nf = (int)F[c].size(), nm = (int)M[c].size();
            if (nf == 0 || nm == 0) continue;
            int np;
            if (nf <= nm) { np = nf; permute(M[c], nf); }
            else { np = nm; permute(F[c], nm); }
            for (int k = 0; k < np; k++) { slots[F[c][k]].mate = M[c][k]; slots[M[c][k]].mate = F[c][k]; }
        }
    }

    // great-circle distance as MoveStats computes it: utils/geomutils.cpp:309-333 with the radius of the Geography
    double msDistance(int from, int to) {
        const std::vector<double> &lon = env["Longitude"], &lat = env["Latitude"];
        const double Q_PI = 3.14159265358979323846;  // (x * Q_PI) / 180 as spherdistDeg writes it: the rounding differs from x * (Q_PI / 180)
        const double lo1 = lon[from] * Q_PI / 180, la1 = lat[from] * Q_PI / 180, lo2 = lon[to] * Q_PI / 180, la2 = lat[to] * Q_PI / 180;
        const double x1 = cos(lo1) * cos(la1), y1 = sin(lo1) * cos(la1), z1 = sin(la1);
        const double x2 = cos(lo2) * cos(la2), y2 = sin(lo2) * cos(la2), z2 = sin(la2);
        double pr = x1 * x2 + y1 * y2 + z1 * z2;
        if (pr > 1) pr = 1; else if (pr < -1) pr = -1;
        return 6371.3 * acos(pr);
    }
    // actions/MoveStats.cpp:107-143 (preLoop) and :148-186 (initializeOccupied: every slot between the first and the last agent)
    void moveStatsPreLoop() {
        msHops.assign(nCells, -1); msDist.assign(nCells, -1.0); msTime.assign(nCells, -1.0);
        msHopsTemp.assign(nCells, -1); msDistTemp.assign(nCells, -1.0); msTimeTemp.assign(nCells, -1.0);
        for (int i = 0; i < hi(); i++) {
            if (!active[i]) continue;
            const int c = slots[i].cell;
            msHops[c] = 0; msDist[c] = 0; msTime[c] = 0;
        }
    }
    // actions/MoveStats.cpp:195-285 (finalize, one thread): the move list is still complete -- ConfinedMove may already have turned
    // some of its entries around, depending on the order of the two in the Prioritizer.  WELL mode follows the list; counter
    // mode does not know an order of the agents: "first" is the move of the agent with the smallest id, "last" the largest.
    void moveStatsFinalize(float t) {
        std::vector<size_t> ord(moveList.size() / 3);
        for (size_t k = 0; k < ord.size(); k++) ord[k] = k;
        if (mode != QOR_MODE_WELL)
            std::stable_sort(ord.begin(), ord.end(), [&](size_t x, size_t y) { return slots[moveList[3 * x + 1]].id < slots[moveList[3 * y + 1]].id; });
        std::set<int> changed;
        for (size_t k : ord) {
            const int from = moveList[3 * k], to = moveList[3 * k + 2];
            const int newHops = msHops[from] + 1;
            const double newDist = msDist[from] + msDistance(from, to), newTime = t;
            if (msHopsTemp[to] < 0 || msMode == 2) {
                msHopsTemp[to] = newHops; msDistTemp[to] = newDist; msTimeTemp[to] = newTime;
                changed.insert(to);
            } else if (msMode == 1) {
                if (newHops < msHopsTemp[to]) msHopsTemp[to] = newHops;
                if (newDist < msDistTemp[to]) msDistTemp[to] = newDist;
                msTimeTemp[to] = newTime;
                changed.insert(to);
            }
        }
        for (int c : changed) {
            if (msHops[c] < 0 || msMode == 2) {
                msHops[c] = msHopsTemp[c]; msDist[c] = msDistTemp[c]; msTime[c] = msTimeTemp[c];
            } else if (msMode == 1) {
                if (msHopsTemp[c] < msHops[c]) msHops[c] = msHopsTemp[c];
                if (msDistTemp[c] < msDist[c]) msDist[c] = msDistTemp[c];
                if (msTimeTemp[c] < msTime[c]) msTime[c] = msTimeTemp[c];
            }
        }
    }

    // ---- execute() of each action -------------------------------------------------------------
    void execute(const Action &act, int i, float t) {
        Agent &a = slots[i];
        switch (act.kind) {
        case A_GETOLD:  // actions/GetOld.cpp:37-48
            if (a.life > 0) a.age = t - a.birth;
            break;
        case A_ATANDEATH: {  // actions/ATanDeath.cpp:66-90
            if (a.life > 0) {
                a.age = t - a.birth;
                double x = atanSlope * (a.age - atanMaxAge);
                double p = 0.5 + atanScale * (mode == QOR_MODE_WELL ? atan(x) : atan_portable(x)) / PI;
                double r = u2d(draw(a.id, STREAM_ACT0, L0_DEATH));
                if (r < p) registerDeath(i);
            }
            break;
        }
        case A_OLDAGEDEATH: {  // actions/OldAgeDeath.cpp:48-67
            if (a.life > 0) {
                a.age = t - a.birth;
                double r = u2range(draw(a.id, STREAM_ACT1, L1_OLDAGE), 1 - oadUncertainty * oadMaxAge, 1 + oadUncertainty * oadMaxAge);
                if (a.age > oadMaxAge + r) registerDeath(i);
            }
            break;
        }
        case A_WEIGHTEDMOVE: {  // actions/WeightedMove.cpp:45-106
            if (a.life > 0) {
                double r = u2d(draw(a.id, STREAM_ACT0, L0_MOVE));
                if (r < moveProb) {
                    int c = a.cell;
                    int nreal = nNbr[c];
                    size_t off = (size_t)c * (maxNeigh + 1);
                    int pick = -1;
                    uint32_t u2 = draw(a.id, STREAM_ACT1, L1_MOVE2);
                    if (W[off] == W[off + nreal]) {
                        pick = (int)u2int(u2, 0, nreal + 1);
                    } else {
                        double r2 = u2d(u2) * W[off + nreal];
                        for (int k = 0; k < nreal + 1; k++) if (r2 < W[off + k]) { pick = k; break; }
                    }
                    if (pick > 0) {
                        int to = nbr[(size_t)c * maxNeigh + pick - 1];
                        if (to >= 0) {
                            bool iced = env.count("Ice") && env["Ice"][to] != 0;
                            if (!iced) registerMove(c, i, to);
                        }
                    }
                }
            }
            break;
        }
        case A_RANDOMMOVE: {  // actions/RandomMove.cpp:65-100: no weights, no ice test; direction 0 = stay
            if (a.life > 0) {
                double r = u2d(draw(a.id, STREAM_ACT0, L0_MOVE));
                if (r < randMoveProb) {
                    int c = a.cell;
                    double r2 = u2d(draw(a.id, STREAM_ACT1, L1_MOVE2));
                    int pick = (int)(r2 * (nNbr[c] + 1));
                    if (pick > 0) {
                        int to = nbr[(size_t)c * maxNeigh + pick - 1];
                        if (to >= 0) registerMove(c, i, to);
                    }
                }
            }
            break;
        }
        case A_WEIGHTEDMOVERAND: {  // actions/WeightedMoveRand.cpp:43-100: the weights decide unless they are all zero, then a uniform pick
            if (a.life > 0) {
                double r = u2d(draw(a.id, STREAM_ACT0, L0_MOVE));
                if (r < moveRandProb) {
                    int c = a.cell;
                    int nreal = nNbr[c];
                    size_t off = (size_t)c * (maxNeigh + 1);
                    int pick = -1;
                    uint32_t u2 = draw(a.id, STREAM_ACT1, L1_MOVE2);
                    if (W[off + maxNeigh] > 0) {
                        double r2 = u2d(u2) * W[off + maxNeigh];
                        for (int k = 0; k < maxNeigh + 1; k++) if (r2 < W[off + k]) { pick = k; break; }
                    } else {
                        pick = (int)u2range(u2, 0, nreal + 1);  // (int) wrandr(0, iNumActualNeigh+1)
                    }
                    if (pick > 0) {
                        int to = nbr[(size_t)c * maxNeigh + pick - 1];
                        if (to >= 0) {
                            bool iced = env.count("Ice") && env["Ice"][to] != 0;
                            if (!iced) registerMove(c, i, to);
                        }
                    }
                }
            }
            break;
        }
        case A_CONDWEIGHTEDMOVE: {  // actions/CondWeightedMove.cpp:41-86: no special case for equal weights, the row is read
            // up to the grid's connectivity, the ice test looks at the cell the agent is IN, the condition decides last
            if (a.life > 0) {
                double r = u2d(draw(a.id, STREAM_ACT0, L0_MOVE));
                if (r < condMoveProb) {
                    int c = a.cell;
                    size_t off = (size_t)c * (maxNeigh + 1);
                    int pick = -1;
                    double r2 = u2d(draw(a.id, STREAM_ACT1, L1_MOVE2)) * W[off + maxNeigh];
                    for (int k = 0; k < maxNeigh + 1; k++) if (r2 < W[off + k]) { pick = k; break; }
                    if (pick > 0) {
                        int to = nbr[(size_t)c * maxNeigh + pick - 1];
                        bool iced = env.count("Ice") && env["Ice"][c] != 0;
                        if (to >= 0 && !iced && condAllow(c, to)) registerMove(c, i, to);
                    }
                }
            }
            break;
        }
        case A_SIGDEATH: {  // actions/SigDeath.cpp:66-90: p = scale / (1 + exp(-slope * (age - maxAge))), one draw per agent
            if (a.life > 0) {
                a.age = t - a.birth;
                const double x = -sigSlope * (a.age - sigMaxAge);
                const double p = sigScale / (1 + (mode == QOR_MODE_WELL ? exp(x) : exp_portable(x)));
                const double r = u2d(draw(a.id, STREAM_ACT1, L1_SIGDEATH));
                if (r < p) registerDeath(i);
            }
            break;
        }
        case A_FERTILITY: {  // actions/Fertility.cpp:49-74 (overwrites the whole life state, incl. the MOVING bit)
            if (a.life > 0) {
                if (a.gender == 0) {
                    a.life = (a.age > fertMinAge && a.age < fertMaxAge && (t - a.lastBirth) > fertInterbirth) ? LIFE_FERTILE : LIFE_ALIVE;
                } else {
                    a.life = (a.age > fertMinAge) ? LIFE_FERTILE : LIFE_ALIVE;
                }
            }
            break;
        }
        case A_VERHULSTVARK:  // actions/VerhulstVarK.cpp:96-112: the same two executes
        case A_VERHULST: {  // actions/Verhulst.cpp:101-115 -> LinearBirth.cpp:122-168 then LinearDeath.cpp:131-153
            if (a.life > 0) {
                int c = a.cell;
                if (B[c] > 0) {
                    if (a.gender == 0 && a.mate >= 0) {
                        double r = u2d(draw(a.id, STREAM_ACT0, L0_BIRTH));
                        if (r < B[c]) registerBirth(c, i, a.mate);
                    }
                } else if (B[c] < 0) {
                    double r = u2d(draw(a.id, STREAM_ACT0, L0_BIRTH));
                    if (r < -B[c]) registerDeath(i);
                }
            }
            if (a.life > 0) {
                double r = u2d(draw(a.id, STREAM_ACT0, L0_DEATH2));
                if (r < D[a.cell]) registerDeath(i);
            }
            break;
        }
        case A_NAVIGATE: {  // actions/Navigate.cpp:181-250
            if (a.life > 0 && (a.life & LIFE_MOVING) == 0) {
                const int c = a.cell;
                auto it = jumpProbs.find(c);
                if (it != jumpProbs.end()) {
                    // the reference bounds the search by the map key = the port's cell index (:194); the table has n+1 entries
                    const int nd = it->second.first, lim = (it->first < nd) ? it->first : nd;
                    const double r = u2d(draw(a.id, STREAM_ACT1, L1_NAV));
                    int i = 0;
                    while (i < lim && r > it->second.second[i].second) i++;
                    if (i > 0) {
                        const int to = it->second.second[i].first;
                        const bool iced = env.count("Ice") && env["Ice"][to] != 0;
                        if (!iced) registerMove(c, (int)(&a - slots.data()), to);
                    }
                }
                for (size_t b = 0; b < curBridges.size(); b++) {
                    const int to = (curBridges[b].first == c) ? curBridges[b].second : ((curBridges[b].second == c) ? curBridges[b].first : -1);
                    if (to >= 0) {
                        uint32_t w;
                        if (mode == QOR_MODE_WELL) w = well.next();
                        else { uint32_t o[4]; draw4(a.id, 0x04000000u | (uint32_t)(b / 4), o); w = o[b % 4]; }
                        if (u2d(w) < navBridgeProb) registerMove(c, (int)(&a - slots.data()), to);
                    }
                }
            }
            break;
        }
        case A_SINGLEEVAL:
        case A_MULTIEVAL:
        case A_NPPCAP:
        case A_GENETICS:
        case A_RANDOMPAIR:
        case A_RANDPERMPAIR:
        case A_MOVESTATS:
        case A_CONFINEDMOVE:
            break;  // execute() is empty for these (actions/Action.h:36 default)
        }
    }

    // action order: priority ascending, same priority in name order (core/SPopulation.cpp:249-257 iterates a
    // std::map<string,int>; core/Prioritizer.cpp:19-30 appends in that order)
    std::vector<const Action *> ordered() const {
        std::vector<const Action *> v;
        for (auto &a : actions) if (a.prio >= 0) v.push_back(&a);
        std::stable_sort(v.begin(), v.end(), [](const Action *x, const Action *y) {
            if (x->prio != y->prio) return x->prio < y->prio;
            return x->name < y->name;
        });
        return v;
    }

    // ---- the step ------------------------------------------------------------------------------
    int initializeStep(float t) {  // core/SPopulation.cpp:394-417
        curTime = t;
        levelsDone.clear();
        for (const Action *a : ordered()) {
            if (!a->enabled) continue;
            switch (a->kind) {
            case A_VERHULST: verhulstInit(); break;
            case A_VERHULSTVARK: verhulstVarKInit(); break;
            case A_MULTIEVAL: multiEvalInit(); break;
            case A_RANDOMPAIR: randomPairInit(); break;
            case A_RANDPERMPAIR: randPermPairInit(); break;
            case A_SINGLEEVAL: evaluatorInit(); break;
            default: break;
            }
        }
        return 0;
    }
    int doActions(unsigned prio, float t) {  // core/SPopulation.cpp:554-577
        for (const Action *a : ordered()) {
            if ((unsigned)a->prio != prio || !a->enabled) continue;
            for (int i = 0; i < hi(); i++) {
                if (active[i] && slots[i].life > LIFE_DEAD) execute(*a, i, t);
            }
        }
        return 0;
    }

    // core/SPopulation.cpp:880-918 (createAgentAtIndex + resetAgent) and populations/tut_EnvironAltPop.cpp:141-149
    void makeBaby(int slot, int cell, int64_t id, uint32_t genderDraw) {
        Agent &b = slots[slot];
        b.id = id;
        b.life = LIFE_ALIVE;
        b.cell = cell;
        b.birth = curTime;
        b.gender = (uint8_t)(2 * u2d(genderDraw));
        if (b.gender == 0) b.life = LIFE_FERTILE;
        b.age = 0.0f;
        b.lastBirth = 0.0f;
        b.mate = -3;
        b.numBabies = 0;
        if (selfMate) { b.gender = 0; b.mate = slot; }  // populations/tut_ParthenoPop.cpp:107-119 (the life state keeps what the draw gave)
    }

    // Genetics::makeOffspring (actions/Genetics.cpp:285-337) under the counter-mode law: every draw is a word of
    // Philox(child id, step, stream) -- stream 4: strand choices and mutation count; 0x01000000|parent<<20|block/2: free
    // recombination masks; 0x02000000|parent<<20|i/4: crossover breaks; 0x03000000|i/4: mutation positions
    void makeGenome(int babySlot, int64_t cid, int mother, int father) {
        const int nb = nBlocks;
        std::vector<uint64_t> out(2 * nb, 0), t1(2 * nb), t2(2 * nb);
        uint32_t g0[4];
        const bool seq = (mode == QOR_MODE_WELL);  // the reference's order of draws from Genetics' generator: strands, mother, father, count, positions
        if (seq) { g0[0] = genWell.next(); g0[1] = genWell.next(); g0[2] = g0[3] = 0; }
        else draw4(cid, 4, g0);
        const int i1 = (int)(2 * 1.0 * u2d(g0[0])), i2 = (int)(2 * 1.0 * u2d(g0[1]));
        const std::vector<uint64_t> &gm = slots[mother].genome, &gf = slots[father].genome;
        auto words = [&](uint32_t base) {
            if (seq) return NextU32([this]() { return genWell.next(); });
            auto state = std::make_shared<std::pair<uint32_t, std::vector<uint32_t>>>(0u, std::vector<uint32_t>());
            return NextU32([this, cid, base, state]() {
                uint32_t i = state->first++;
                uint32_t o[4];
                draw4(cid, base | (i / 4), o);
                return o[i % 4];
            });
        };
        if (numCrossOvers > 0) {
            bitCrossOver(t1.data(), gm.data(), genomeSize, numCrossOvers, words(0x02000000u | (0u << 20)), bitsPerNuc);
            bitCrossOver(t2.data(), gf.data(), genomeSize, numCrossOvers, words(0x02000000u | (1u << 20)), bitsPerNuc);
        } else if (numCrossOvers == -1) {
            bitFreeReco(t1.data(), gm.data(), nb, words(0x01000000u | (0u << 20)), bitsPerNuc);
            bitFreeReco(t2.data(), gf.data(), nb, words(0x01000000u | (1u << 20)), bitsPerNuc);
        } else {
            t1 = gm; t2 = gf;
        }
        for (int q = 0; q < nb; q++) { out[q] = t1[i1 * nb + q]; out[nb + q] = t2[i2 * nb + q]; }
        if (mutationRate > 0) {
            if (seq) g0[2] = genWell.next();
            int nMut = binomialGetN(binoTable, u2d(g0[2]));
            if (nMut > 0) {  // U::mutateNucs(pBabyGenome, m_iNumParents*m_iGenomeSize, ...), actions/Genetics.cpp:332
                if (bitsPerNuc == 2) nucMutate(out.data(), 2 * genomeSize, nMut, words(0x03000000u));
                else bitMutate(out.data(), 2 * genomeSize, nMut, words(0x03000000u));
            }
        }
        slots[babySlot].genome = out;
    }

    // core/SPopulation.cpp:596-724 recycleDeadSpaceNew (+ performBirths :778-815, performDeaths :974-989)
    void recycleDeadSpace() {
        size_t nBirths = birthList.size() / 3, nDeaths = deathList.size();
        stepBirths = nBirths; stepDeaths = nDeaths;
        totBirths += nBirths; totDeaths += nDeaths;
        // counter mode: newborn id = nextID + rank of (cell, mother id) among this step's births
        std::vector<int64_t> babyId(nBirths);
        if (mode == QOR_MODE_COUNTER) {
            std::vector<size_t> ord(nBirths);
            for (size_t k = 0; k < nBirths; k++) ord[k] = k;
            std::sort(ord.begin(), ord.end(), [&](size_t x, size_t y) {
                int cx = birthList[3 * x], cy = birthList[3 * y];
                if (cx != cy) return cx < cy;
                return slots[birthList[3 * x + 1]].id < slots[birthList[3 * y + 1]].id;
            });
            for (size_t r = 0; r < nBirths; r++) babyId[ord[r]] = nextID + birthIdOffset + (int64_t)r;
            nextID += (birthIdTotal >= 0) ? birthIdTotal : (int64_t)nBirths;
            birthIdOffset = 0; birthIdTotal = -1;
        }
        size_t nReuse = std::min(prevDead.size(), nBirths);
        auto born = [&](int slot, size_t k) {
            int64_t id;
            uint32_t g;
            if (mode == QOR_MODE_WELL) {  // IDGen::getID with one thread (core/IDGen.h:28), then the gender draw
                id = nextID++;
                g = well.next();
            } else {
                id = babyId[k];
                uint32_t o[4];
                draw4(id, STREAM_BABY, o);
                g = o[0];
            }
            makeBaby(slot, birthList[3 * k], id, g);
            if (find("Genetics")) {  // populations/OoANavGenPop.cpp:231-245 makePopSpecificOffspring
                makeGenome(slot, id, birthList[3 * k + 1], birthList[3 * k + 2]);
                slots[birthList[3 * k + 1]].numBabies++;
            }
        };
        for (size_t k = 0; k < nReuse; k++) born(prevDead[k], k);          // babies into last step's dead slots
        for (size_t k = nReuse; k < nBirths; k++) born(allocSlot(), k);     // remaining births: lowest free index
        for (size_t k = nReuse; k < prevDead.size(); k++) freeSlot(prevDead[k]);  // leftover dead are unlinked
        prevDead = deathList;
        for (int i : prevDead) slots[i].life = LIFE_DEAD;
    }
    // core/SPopulation.cpp:1014-1027,1058-1092
    void performMoves() {
        stepMoves = moveList.size() / 3;
        totMoves += stepMoves;
        for (size_t k = 0; k < moveList.size(); k += 3) {
            Agent &a = slots[moveList[k + 1]];
            a.cell = moveList[k + 2];
            a.life &= ~LIFE_MOVING;
        }
    }
    // actions/ConfinedMove.cpp:44-78 (preLoop): the cells within m_dR km of (m_dX, m_dY); the grid built by the driver has no
    // type (GRID_TYPE_NONE), which takes the icosahedral branch: great-circle distance (utils/geomutils.cpp:311-326)
    void confinedPreLoop() {
        const std::vector<double> &lon = env["Longitude"], &lat = env["Latitude"];
        confAllowed.assign(nCells, 0);
        const double conv = 3.14159 / 180.0;  // the reference's constant (actions/ConfinedMove.cpp:66)
        for (int i = 0; i < nCells; i++) {
            const double lo1 = lon[i] * conv, la1 = lat[i] * conv, lo2 = confX * conv, la2 = confY * conv;
            const double x1 = cos(lo1) * cos(la1), y1 = sin(lo1) * cos(la1), z1 = sin(la1);
            const double x2 = cos(lo2) * cos(la2), y2 = sin(lo2) * cos(la2), z2 = sin(la2);
            double pr = x1 * x2 + y1 * y2 + z1 * z2;
            if (pr > 1) pr = 1; else if (pr < -1) pr = -1;
            if (6371.3 * acos(pr) < confR) confAllowed[i] = 1;  // RADIUS_EARTH_KM, utils/qhg_consts.h:54-55
        }
    }
    int finalizeStep() {  // core/SPopulation.cpp:439-477
        for (const Action *a : ordered()) {
            if (a->enabled && a->kind == A_SINGLEEVAL) evalNeedUpdate = false;  // actions/SingleEvaluator.cpp:125-130
            if (a->enabled && a->kind == A_MULTIEVAL) for (auto &e : subs) e.needUpdate = false;  // actions/MultiEvaluator.cpp:189-195
            // ConfinedMove::finalize (actions/ConfinedMove.cpp:86-101): every registered move into a cell outside the
            // region is turned into a move to the cell it starts from (it stays in the list and is counted)
            if (a->enabled && a->kind == A_CONFINEDMOVE)
                for (size_t k = 0; k < moveList.size(); k += 3) if (!confAllowed[moveList[k + 2]]) moveList[k + 2] = moveList[k];
            if (a->enabled && a->kind == A_MOVESTATS) moveStatsFinalize(curTime);
        }
        recycleDeadSpace();
        performMoves();
        updateNumAgentsPerCell();
        birthList.clear(); deathList.clear(); moveList.clear();
        stepIndex++;
        return 0;
    }
    int step(float t) {  // core/PopLooper.cpp:166-202 for one population
        initializeStep(t);
        std::vector<unsigned> levels;
        for (const Action *a : ordered()) if (levels.empty() || levels.back() != (unsigned)a->prio) levels.push_back((unsigned)a->prio);
        for (unsigned l : levels) doActions(l, t);
        return finalizeStep();
    }
    // populations/tut_EnvironAltPop.cpp:93-127
    int updateEvent(int ev, float) {
        // only the classes that override updateEvent drown their agents (populations/tut_EnvironAltPop.cpp:93-127,
        // tut_EnvironCapAltPop.cpp, OoANavGenPop.cpp:179-214); tut_SexualPop, tut_MovePop, tut_OldAgeDiePop, tut_ParthenoPop and
        // tut_StaticPop inherit SPopulation::updateEvent, which does nothing (core/SPopulation.h:116)
        const bool drowns = popClass.rfind("tut_Environ", 0) == 0 || popClass.rfind("OoANavGen", 0) == 0;
        if (ev == EVENT_ID_GEO && drowns) {
            const std::vector<double> &alt = env["Altitude"];
            const std::vector<double> *ice = env.count("Ice") ? &env["Ice"] : nullptr;
            for (int i = 0; i < hi(); i++) {
                if (active[i] && slots[i].life > LIFE_DEAD) {
                    int c = slots[i].cell;
                    if (alt[c] < 0 || (ice && (*ice)[c] > 0)) registerDeath(i);
                }
            }
            recycleDeadSpace();
            updateNumAgentsPerCell();
            birthList.clear(); deathList.clear(); moveList.clear();
            // notifyObservers(EVENT_ID_GEO) follows (populations/tut_EnvironAltPop.cpp:120), but tut_EnvironAltPop never
            // registers its evaluator as an observer (no addObserver in populations/tut_EnvironAltPop.cpp:24-53, unlike
            // populations/OoANavGenPop.cpp:59), so SingleEvaluator::notify (actions/SingleEvaluator.cpp:332-346) is never
            // reached: the weights computed at the first step stay in force.  Replicated as is.
            if (evaluatorObserves) evalNeedUpdate = true;
        }
        // NPPCapacity registers itself as an observer (actions/NPPCapacity.cpp:69) and reacts to GEO, CLIMATE and VEG (:121-131);
        // the MultiEvaluator of tut_EnvironCapAltPop is never registered, so its weights stay as first computed
        if (ev == 2 || ev == 3 || ev == 4) nppNeedUpdate = true;
        // OoANavGenPop registers its MultiEvaluator (populations/OoANavGenPop.cpp:59), which forwards the event to its
        // evaluators (actions/MultiEvaluator.cpp:203-214): each one reacts to its own trigger id
        if (multiObserves) for (auto &e : subs) if (e.trigger == ev) e.needUpdate = true;
        if (ev == 2 || ev == 5) navNeedUpdate = true;  // Navigate::notify, actions/Navigate.cpp:79-87
        return 0;
    }
};

// tut_EnvironCapAlt<Mode>Pop -> MultiEvalModes value, -1 for any other name
static int multiProbeMode(const std::string &cls) {
    static const char *const names[] = {nullptr, "tut_EnvironCapAltAddBlockPop", "tut_EnvironCapAltMulPop", "tut_EnvironCapAltMaxPop",
                                        "tut_EnvironCapAltMaxBlockPop", "tut_EnvironCapAltMinPop"};
    for (int m = 1; m <= 5; m++) if (cls == names[m]) return m;
    return -1;
}

// =================================================================================================
extern "C" {

qor_pop *qor_create(const char *pop_class, int n_cells, int max_neigh, int mode) {
    qor_pop *p = new qor_pop;
    p->popClass = pop_class;
    p->nCells = n_cells;
    p->maxNeigh = max_neigh;
    p->mode = mode;
    p->counts.assign(n_cells, 0);
    p->B.assign(n_cells, 0.0);
    p->D.assign(n_cells, 0.0);
    p->W.assign((size_t)n_cells * (max_neigh + 1), 0.0);
    if (p->popClass == "tut_EnvironAltPop") {  // populations/tut_EnvironAltPop.cpp:24-53
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"Verhulst", A_VERHULST},
                      {"RandomPair", A_RANDOMPAIR}, {"SingleEvaluator[Alt]", A_SINGLEEVAL}, {"WeightedMove", A_WEIGHTEDMOVE}};
    } else if (p->popClass == "tut_EnvironAltNavPop") {
        // probe class (no such class ships): tut_EnvironAltPop with Navigate and OldAgeDeath added, the way NavProbePop in
        // oracle/ref_driver.cpp adds the reference's Navigate<T> to the reference's tut_EnvironAltPop -- pins Navigate
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"Verhulst", A_VERHULST},
                      {"RandomPair", A_RANDOMPAIR}, {"SingleEvaluator[Alt]", A_SINGLEEVAL}, {"WeightedMove", A_WEIGHTEDMOVE},
                      {"Navigate", A_NAVIGATE}, {"OldAgeDeath", A_OLDAGEDEATH}};
    } else if (p->popClass == "tut_EnvironAltConfPop") {
        // probe class: tut_EnvironAltPop with ConfinedMove added (ConfProbePop in oracle/ref_driver.cpp) -- pins ConfinedMove
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"Verhulst", A_VERHULST},
                      {"RandomPair", A_RANDOMPAIR}, {"SingleEvaluator[Alt]", A_SINGLEEVAL}, {"WeightedMove", A_WEIGHTEDMOVE},
                      {"ConfinedMove", A_CONFINEDMOVE}};
    } else if (p->popClass == "tut_EnvironAltVarPop") {
        // probe class: tut_EnvironAltPop with WeightedMoveRand and SigDeath added (VarProbePop in oracle/ref_driver.cpp)
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"Verhulst", A_VERHULST},
                      {"RandomPair", A_RANDOMPAIR}, {"SingleEvaluator[Alt]", A_SINGLEEVAL}, {"WeightedMove", A_WEIGHTEDMOVE},
                      {"WeightedMoveRand", A_WEIGHTEDMOVERAND}, {"SigDeath", A_SIGDEATH}};
    } else if (p->popClass.size() == 22 && p->popClass.rfind("tut_EnvironAltCond", 0) == 0 && p->popClass.substr(19) == "Pop" &&
               p->popClass[18] >= '0' && p->popClass[18] <= '7') {
        // probe classes tut_EnvironAltCond<m>Pop: tut_EnvironAltPop with CondWeightedMove (a SimpleCondition of mode m over the
        // altitudes), RandPermPair and MoveStats added (ExtProbePop<m> in oracle/ref_driver.cpp); the <prio> entries decide which
        // of WeightedMove / CondWeightedMove and RandomPair / RandPermPair run
        p->condMode = p->popClass[18] - '0';
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"Verhulst", A_VERHULST},
                      {"RandomPair", A_RANDOMPAIR}, {"SingleEvaluator[Alt]", A_SINGLEEVAL}, {"WeightedMove", A_WEIGHTEDMOVE},
                      {"CondWeightedMove", A_CONDWEIGHTEDMOVE}, {"RandPermPair", A_RANDPERMPAIR}, {"MoveStats", A_MOVESTATS}};
    } else if (p->popClass == "tut_EnvironAltGenPop" || p->popClass == "tut_EnvironAltGen2bitPop") {
        // probe classes: tut_EnvironAltPop with Genetics<.., BitGeneUtils> resp. Genetics<.., GeneUtils> added and called from
        // makePopSpecificOffspring (GenProbePop<U> in oracle/ref_driver.cpp) -- pin the Genetics action with 1- and 2-bit nucleotides
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"Verhulst", A_VERHULST},
                      {"RandomPair", A_RANDOMPAIR}, {"SingleEvaluator[Alt]", A_SINGLEEVAL}, {"WeightedMove", A_WEIGHTEDMOVE},
                      {"Genetics", A_GENETICS}};
        p->bitsPerNuc = (p->popClass == "tut_EnvironAltGen2bitPop") ? 2 : 1;
    } else if (p->popClass == "tut_ParthenoPop") {  // populations/tut_ParthenoPop.cpp:22-45
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"RandomMove", A_RANDOMMOVE}, {"Verhulst", A_VERHULST},
                      {"Fertility", A_FERTILITY}};
        p->selfMate = true;
    } else if (p->popClass == "tut_StaticPop") {  // populations/tut_StaticPop.cpp:16-21: no actions at all
        p->actions = {};
    } else if (p->popClass == "tut_SexualPop") {  // populations/tut_SexualPop.cpp:24-44
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"RandomMove", A_RANDOMMOVE}, {"Fertility", A_FERTILITY},
                      {"Verhulst", A_VERHULST}, {"RandomPair", A_RANDOMPAIR}};
    } else if (p->popClass == "tut_MovePop") {  // populations/tut_MovePop.cpp:19-31
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"RandomMove", A_RANDOMMOVE}};
    } else if (p->popClass == "tut_OldAgeDiePop") {  // populations/tut_OldAgeDiePop.cpp:17-26
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}};
    } else if (p->popClass == "tut_EnvironCapAltPop" || multiProbeMode(p->popClass) >= 0) {  // populations/tut_EnvironCapAltPop.cpp:27-72
        p->actions = {{"GetOld", A_GETOLD}, {"ATanDeath", A_ATANDEATH}, {"Fertility", A_FERTILITY}, {"VerhulstVarK", A_VERHULSTVARK},
                      {"RandomPair", A_RANDOMPAIR}, {"MultiEvaluator[NPP+Alt]", A_MULTIEVAL}, {"WeightedMove", A_WEIGHTEDMOVE},
                      {"NPPCapacity", A_NPPCAP}};
        SubEval ea; ea.input = "Altitude"; ea.weightName = "Multi_weight_alt"; ea.usePoly = true;
        SubEval en; en.input = ""; en.weightName = "Multi_weight_npp"; en.usePoly = false;
        p->subs = {ea, en};
        if (multiProbeMode(p->popClass) >= 0) {
            // probe classes (MultiProbePop<MODE> in oracle/ref_driver.cpp): tut_EnvironCapAltPop whose MultiEvaluator combines in
            // another mode over NON-cumulating evaluators and is registered as an observer, the way populations/OoANavPop.cpp:50-62
            // builds its MODE_MUL_SIMPLE evaluator
            p->multiMode = multiProbeMode(p->popClass);
            for (auto &e : p->subs) e.cumulate = false;
            p->subs[0].trigger = EVENT_ID_GEO; p->subs[1].trigger = 4;
            p->multiObserves = true;
        }
        p->cap.assign(n_cells, 0.0);
        for (const char *nm : {"Water", "Coastal", "Latitude", "Longitude", "AnnualMeanTemp", "AnnualRainfall", "BaseNPP"})
            p->env[nm].assign(n_cells, 0.0);
    } else if (p->popClass == "OoANavGenPop" || p->popClass == "OoANavGen2bitPop") {
        // populations/OoANavGenPop.cpp:33-97; populations/OoANavGen2bitPop.cpp is the same class with Genetics<.., GeneUtils>
        // (2-bit nucleotides) and WITHOUT addObserver(m_pME)
        // WELL mode: OoANavGenPop itself is part of oracle/_ref (tests/test_oracle_vs_ref.py pins the class); its 2-bit sibling is not
        if (mode != QOR_MODE_COUNTER && p->popClass != "OoANavGenPop") { delete p; return nullptr; }
        p->bitsPerNuc = (p->popClass == "OoANavGen2bitPop") ? 2 : 1;
        p->actions = {{"MultiEvaluator[Alt+NPP]", A_MULTIEVAL}, {"WeightedMove", A_WEIGHTEDMOVE}, {"VerhulstVarK", A_VERHULSTVARK},
                      {"RandomPair", A_RANDOMPAIR}, {"GetOld", A_GETOLD}, {"OldAgeDeath", A_OLDAGEDEATH}, {"Fertility", A_FERTILITY},
                      {"NPPCapacity", A_NPPCAP}, {"Genetics", A_GENETICS}, {"Navigate", A_NAVIGATE}};
        SubEval ea; ea.input = "Altitude"; ea.weightName = "Multi_weight_alt"; ea.usePoly = true; ea.polyName = "AltCapPref"; ea.trigger = EVENT_ID_GEO;
        SubEval en; en.input = ""; en.weightName = "Multi_weight_npp"; en.usePoly = true; en.polyName = "NPPPref"; en.trigger = 4;
        p->subs = {ea, en};
        p->multiObserves = (p->bitsPerNuc == 1);  // addObserver(m_pME), populations/OoANavGenPop.cpp:59; absent in OoANavGen2bitPop.cpp
        p->cap.assign(n_cells, 0.0);
        for (const char *nm : {"Water", "Coastal", "Latitude", "Longitude", "AnnualMeanTemp", "AnnualRainfall", "BaseNPP"})
            p->env[nm].assign(n_cells, 0.0);
    } else {
        delete p;
        return nullptr;
    }
    static const uint32_t zero[16] = {0};
    qor_set_seed(p, zero);
    return p;
}

void qor_destroy(qor_pop *p) { delete p; }

int qor_set_cells(qor_pop *p, const int32_t *nbr, const int32_t *global_id) {
    size_t n = (size_t)p->nCells * p->maxNeigh;
    p->nbr.assign(nbr, nbr + n);
    p->nNbr.assign(p->nCells, 0);
    p->gid.resize(p->nCells);
    for (int c = 0; c < p->nCells; c++) {
        int k = 0;
        for (int j = 0; j < p->maxNeigh; j++) if (nbr[(size_t)c * p->maxNeigh + j] >= 0) k++;
        p->nNbr[c] = (uint8_t)k;
        p->gid[c] = global_id ? global_id[c] : c;
    }
    return 0;
}

int qor_set_env_array(qor_pop *p, const char *name, const double *v, int64_t n) {
    if (n != p->nCells) return -1;
    p->env[name].assign(v, v + n);
    return 0;
}

int qor_set_env_delta(qor_pop *p, const char *name, const double *delta, int64_t n) {
    if (!delta) { p->envDelta.erase(name); return 0; }
    if (n != p->nCells || !p->env.count(name)) return -1;
    p->envDelta[name].assign(delta, delta + n);
    return 0;
}

// AutoInterpolator::interpolate (core/AutoInterpolator.cpp:461-483): pTarget[i] += iSteps*pSource[i] (for iSteps == 1 the
// reference writes += pSource[i], the same value)
int qor_interpolate_env(qor_pop *p, int steps) {
    for (auto &kv : p->envDelta) {
        std::vector<double> &t = p->env[kv.first];
        for (int i = 0; i < p->nCells; i++) t[i] += steps * kv.second[i];
    }
    return 0;
}

int qor_get_env_array(qor_pop *p, const char *name, double *out) {
    auto it = p->env.find(name);
    if (it == p->env.end()) return -1;
    memcpy(out, it->second.data(), sizeof(double) * p->nCells);
    return 0;
}

int qor_set_attribute(qor_pop *p, const char *name, double v) {
    std::string s(name);
    if (s == "ATanDeath_max_age") p->atanMaxAge = v;
    else if (s == "ATanDeath_range") p->atanRange = v;
    else if (s == "ATanDeath_slope") p->atanSlope = v;
    else if (s == "OAD_max_age") p->oadMaxAge = v;
    else if (s == "OAD_uncertainty") p->oadUncertainty = v;
    else if (s == "WeightedMove_prob") p->moveProb = v;
    else if (s == "RandomMove_prob") p->randMoveProb = v;
    else if (s == "WeightedMoveRand_prob") p->moveRandProb = v;
    else if (s == "CondWeightedMove_prob") p->condMoveProb = v;
    else if (s == "MoveStats_Mode") p->msMode = (int)v;
    else if (s == "SigDeath_max_age") p->sigMaxAge = v;
    else if (s == "SigDeath_range") p->sigRange = v;
    else if (s == "SigDeath_slope") p->sigSlope = v;
    else if (s == "ConfinedMove_x") p->confX = v;
    else if (s == "ConfinedMove_y") p->confY = v;
    else if (s == "ConfinedMove_r") p->confR = v;
    else if (s == "Fertility_min_age") p->fertMinAge = (float)v;
    else if (s == "Fertility_max_age") p->fertMaxAge = (float)v;
    else if (s == "Fertility_interbirth") p->fertInterbirth = (float)v;
    else if (s == "Verhulst_b0") p->vB0 = v;
    else if (s == "Verhulst_d0") p->vD0 = v;
    else if (s == "Verhulst_theta") p->vTheta = v;
    else if (s == "Verhulst_K") p->vK = v;
    else if (s == "NPPCap_water_factor") p->nppWater = v;
    else if (s == "NPPCap_coastal_factor") p->nppCoastal = v;
    else if (s == "NPPCap_coastal_min_latitude") p->nppCoastMinLat = v;
    else if (s == "NPPCap_coastal_max_latitude") p->nppCoastMaxLat = v;
    else if (s == "NPPCap_NPP_min") p->nppMin = v;
    else if (s == "NPPCap_NPP_max") p->nppMax = v;
    else if (s == "NPPCap_K_max") p->nppKMax = v;
    else if (s == "NPPCap_K_min") p->nppKMin = v;
    else if (s == "NPPCap_efficiency") p->nppEff = v;
    else if (s == "Navigate_decay") p->navDecay = v;
    else if (s == "Navigate_dist0") p->navDist0 = v;
    else if (s == "Navigate_prob0") p->navProb0 = v;
    else if (s == "Navigate_min_dens") {}
    else if (s == "Navigate_bridge_prob") p->navBridgeProb = v;
    else if (s == "Genetics_genome_size") { p->genomeSize = (int)v; p->nBlocks = ((int)v * p->bitsPerNuc + 63) / 64; }
    else if (s == "Genetics_num_crossover") p->numCrossOvers = (int)v;
    else if (s == "Genetics_mutation_rate") p->mutationRate = v;
    else if (s == "Genetics_create_new_genome" || s == "Genetics_bits_per_nuc") { if (s == "Genetics_bits_per_nuc" && (int)v != p->bitsPerNuc) return -1; }
    else if (s == "Multi_weight_alt" || s == "Multi_weight_npp") { for (auto &e : p->subs) if (e.weightName == s) e.weight = v; }
    else return -1;
    return 0;
}

int qor_set_attribute_str(qor_pop *p, const char *name, const char *v) {
    for (auto &e : p->subs) {
        if (!e.polyName.empty() && e.polyName == name) return e.poly.parse(v) ? 0 : -1;
    }
    if (std::string(name) == "Genetics_initial_muts") return 0;  // genomes are supplied by the host (qor_set_genomes)
    if (std::string(name) == "AltCapPref" || std::string(name) == "AltPref") {
        p->havePoly = p->altPref.parse(v);
        return p->havePoly ? 0 : -1;
    }
    char *e;
    double d = strtod(v, &e);
    if (e == v) return -1;
    return qor_set_attribute(p, name, d);
}

int qor_set_prio(qor_pop *p, const char *action, int prio) {
    Action *a = p->find(action);
    if (!a) return -1;
    a->prio = prio;
    return 0;
}

int qor_enable_action(qor_pop *p, const char *action, int enabled) {
    Action *a = p->find(action);
    if (!a) return -1;
    a->enabled = enabled != 0;
    return 0;
}

int qor_set_seed(qor_pop *p, const uint32_t *st) {
    memcpy(p->state16, st, sizeof(p->state16));
    uint32_t tmp[16];
    for (int j = 0; j < 16; j++) tmp[j] = st[(0 + 13 * j) % 16];  // thread 0, core/SPopulation.cpp:171-176
    p->well.seed(tmp);
    p->key[0] = p->key[1] = 0;
    for (int j = 0; j < 16; j += 2) { p->key[0] ^= st[j]; p->key[1] ^= st[j + 1]; }
    return 0;
}

int qor_add_agents(qor_pop *p, int64_t n, const int32_t *cell, const int64_t *id, const float *birth,
                   const uint8_t *gender, const float *age, const float *last_birth, const uint32_t *life) {
    for (int64_t j = 0; j < n; j++) {
        if (cell[j] < 0 || cell[j] >= p->nCells) return -1;
        int i = p->allocSlot();
        Agent &a = p->slots[i];
        a.life = life ? life[j] : LIFE_ALIVE;
        a.cell = cell[j];
        a.id = id[j];
        a.birth = birth[j];
        a.gender = gender[j];
        a.age = age ? age[j] : 0.0f;
        a.lastBirth = last_birth ? last_birth[j] : 0.0f;
        a.mate = p->selfMate ? 0 : -3;
        if (a.id > p->maxID) p->maxID = a.id;
    }
    return 0;
}

int qor_pre_loop(qor_pop *p) {  // core/SPopulation.cpp:273-292 + actions/ATanDeath.cpp:49-59 + app/Simulator.cpp:107-111
    p->atanScale = (PI / 2 - ATAN_EPS) / atan(p->atanSlope * p->atanRange);
    p->sigScale = 1 + exp(-p->sigRange);  // SigDeath::preLoop, actions/SigDeath.cpp:49-60
    p->nextID = p->maxID + 1;
    p->updateNumAgentsPerCell();
    if (p->find("Genetics")) {  // Genetics::init, actions/Genetics.cpp:196-267
        if (p->genomeSize <= 0) return -1;
        if (p->mutationRate > 0) p->binoTable = binomialTable(p->mutationRate, 2 * p->genomeSize, 1e-6);
        for (auto &a : p->slots) if (a.genome.empty()) a.genome.assign(2 * p->nBlocks, 0);
    }
    { Action *nv = p->find("Navigate"); if (nv && nv->prio >= 0 && p->navRecalculate() != 0) return -1; }  // Navigate::preLoop
    if (p->find("NPPCapacity")) p->nppRecalculate();  // NPPCapacity::preLoop, actions/NPPCapacity.cpp:92-115
    { Action *ms = p->find("MoveStats"); if (ms && ms->prio >= 0) { if (!p->env.count("Longitude") || !p->env.count("Latitude")) return -1; p->moveStatsPreLoop(); } }
    { Action *cm = p->find("ConfinedMove"); if (cm && cm->prio >= 0) { if (!p->env.count("Longitude") || !p->env.count("Latitude")) return -1; p->confinedPreLoop(); } }
    return 0;
}

int qor_initialize_step(qor_pop *p, float t) { return p->initializeStep(t); }
int qor_do_actions(qor_pop *p, unsigned prio, float t) { return p->doActions(prio, t); }
int qor_finalize_step(qor_pop *p) { return p->finalizeStep(); }
int qor_step(qor_pop *p, float t) { return p->step(t); }
int qor_update_event(qor_pop *p, int ev, float t) { return p->updateEvent(ev, t); }
int qor_flush_events(qor_pop *p, float) {  // EVENT_ID_FLUSH: NPPCapacity::notify -> recalculate (actions/NPPCapacity.cpp:127-130)
    if (p->find("NPPCapacity")) p->nppRecalculate();
    { Action *nv = p->find("Navigate"); if (nv && nv->prio >= 0) p->navRecalculate(); }
    return 0;
}

int64_t qor_get_num_agents_effective(qor_pop *p) { return p->numUsed - (int64_t)p->prevDead.size(); }  // core/SPopulation.h:148

int qor_get_num_agents_array(qor_pop *p, uint64_t *out) {
    memcpy(out, p->counts.data(), sizeof(uint64_t) * p->nCells);
    return 0;
}

int64_t qor_get_agents(qor_pop *p, int64_t cap, int32_t *cell, int64_t *id, float *birth, uint8_t *gender,
                       float *age, float *last_birth, uint32_t *life, int64_t *mate_id, int32_t *slot) {
    int64_t k = 0;
    for (int i = 0; i < p->hi(); i++) {
        if (!p->active[i] || p->slots[i].life == LIFE_DEAD) continue;
        const Agent &a = p->slots[i];
        if (k < cap) {
            if (cell) cell[k] = a.cell;
            if (id) id[k] = a.id;
            if (birth) birth[k] = a.birth;
            if (gender) gender[k] = a.gender;
            if (age) age[k] = a.age;
            if (last_birth) last_birth[k] = a.lastBirth;
            if (life) life[k] = a.life;
            if (mate_id) mate_id[k] = (a.mate >= 0) ? p->slots[a.mate].id : (int64_t)a.mate;
            if (slot) slot[k] = i;
        }
        k++;
    }
    return k;
}

int qor_get_env_weights(qor_pop *p, double *out) {
    memcpy(out, p->W.data(), sizeof(double) * p->W.size());
    return 0;
}

int qor_get_birth_death_probs(qor_pop *p, double *b, double *d) {
    memcpy(b, p->B.data(), sizeof(double) * p->nCells);
    memcpy(d, p->D.data(), sizeof(double) * p->nCells);
    return 0;
}

int qor_set_navigation(qor_pop *p, int n_ports, const int32_t *port_cell, const int32_t *port_ptr, const int32_t *dest_cell,
                       const double *dist, int n_bridges, const int32_t *bridges) {
    p->navDest.clear();
    p->navBridges.clear();
    for (int pt = 0; pt < n_ports; pt++)
        for (int k = port_ptr[pt]; k < port_ptr[pt + 1]; k++) p->navDest[port_cell[pt]][dest_cell[k]] = dist[k];
    for (int b = 0; b < n_bridges; b++) p->navBridges.push_back({bridges[2 * b], bridges[2 * b + 1]});
    p->navNeedUpdate = true;
    return 0;
}

// genomes of the agents added so far, in add order (2*nBlocks words each); call after qor_add_agents
int qor_set_genomes(qor_pop *p, int64_t n, const uint64_t *g) {
    if (p->nBlocks <= 0 || n != (int64_t)p->slots.size()) return -1;
    const size_t row = 2 * (size_t)p->nBlocks;
    for (int64_t i = 0; i < n; i++) p->slots[i].genome.assign(g + i * row, g + (i + 1) * row);
    return 0;
}

// genomes and NumBabies of the live agents in the order of qor_get_agents
int64_t qor_get_genomes(qor_pop *p, int64_t cap, uint64_t *g, int32_t *num_babies) {
    const size_t row = 2 * (size_t)p->nBlocks;
    int64_t k = 0;
    for (int i = 0; i < p->hi(); i++) {
        if (!p->active[i] || p->slots[i].life == LIFE_DEAD) continue;
        if (k < cap) {
            if (g && p->slots[i].genome.size() == row) memcpy(g + k * row, p->slots[i].genome.data(), row * sizeof(uint64_t));
            if (num_babies) num_babies[k] = p->slots[i].numBabies;
        }
        k++;
    }
    return k;
}

// ---- genome primitives with the reference's own draw order from ONE WELL512 (pins them against genes/BitGeneUtils.cpp) ----
int qor_bit_crossover(const uint32_t *state16, const uint64_t *in, int genome_size, int n_cross, uint64_t *out) {
    Well512 w; w.seed(state16);
    bitCrossOver(out, in, genome_size, n_cross, [&w]() { return w.next(); });
    return 0;
}
int qor_bit_freereco(const uint32_t *state16, const uint64_t *in, int n_blocks, uint64_t *out) {
    Well512 w; w.seed(state16);
    bitFreeReco(out, in, n_blocks, [&w]() { return w.next(); });
    return 0;
}
int qor_bit_mutate(const uint32_t *state16, uint64_t *genome, int n_bits, int n_mut) {
    Well512 w; w.seed(state16);
    bitMutate(genome, n_bits, n_mut, [&w]() { return w.next(); });
    return 0;
}
// genes/GeneUtils.cpp (2-bit nucleotides): the same three primitives
int qor_gene2_crossover(const uint32_t *state16, const uint64_t *in, int genome_size, int n_cross, uint64_t *out) {
    Well512 w; w.seed(state16);
    bitCrossOver(out, in, genome_size, n_cross, [&w]() { return w.next(); }, 2);
    return 0;
}
int qor_gene2_freereco(const uint32_t *state16, const uint64_t *in, int n_blocks, uint64_t *out) {
    Well512 w; w.seed(state16);
    bitFreeReco(out, in, n_blocks, [&w]() { return w.next(); }, 2);
    return 0;
}
int qor_gene2_mutate(const uint32_t *state16, uint64_t *genome, int n_nucs, int n_mut) {
    Well512 w; w.seed(state16);
    nucMutate(genome, n_nucs, n_mut, [&w]() { return w.next(); });
    return 0;
}
// WELL mode: state and index of the Genetics action's own generator (in the reference it is built from aiSeeds[1] through MD5
// digests of seed phrases, utils/WELLUtils.cpp:127-148; the test reads it from the reference and hands it over)
int qor_set_genetics_well(qor_pop *p, const uint32_t *state16, uint32_t index) {
    p->genWell.seed(state16);
    p->genWell.idx = index & 15u;
    p->haveGenWell = true;
    return 0;
}

int qor_binomial_table(double prob, int n, double eps, int cap, double *out) {
    std::vector<double> t = binomialTable(prob, n, eps);
    for (size_t i = 0; i < t.size() && (int)i < cap; i++) out[i] = t[i];
    return (int)t.size();
}
int qor_binomial_get_n(double prob, int n, double eps, double r) { return binomialGetN(binomialTable(prob, n, eps), r); }

int qor_get_capacities(qor_pop *p, double *out) {
    if (p->cap.empty()) return -1;
    memcpy(out, p->cap.data(), sizeof(double) * p->nCells);
    return 0;
}

int qor_get_move_stats(qor_pop *p, int32_t *hops, double *dist, double *time) {  /* MoveStats: m_aiHops, m_adDist, m_adTime */
    if (p->msHops.empty()) return -1;
    memcpy(hops, p->msHops.data(), sizeof(int32_t) * p->nCells);
    memcpy(dist, p->msDist.data(), sizeof(double) * p->nCells);
    memcpy(time, p->msTime.data(), sizeof(double) * p->nCells);
    return 0;
}

int qor_atan_death_prob(qor_pop *p, int n, const float *age, double *out) {
    for (int i = 0; i < n; i++) {
        double x = p->atanSlope * (age[i] - p->atanMaxAge);
        out[i] = 0.5 + p->atanScale * (p->mode == QOR_MODE_WELL ? atan(x) : atan_portable(x)) / PI;
    }
    return 0;
}

// ---- sharded runs (the protocol of the CUDA path's multi-GPU mode, counter mode only) --------------------------
int64_t qor_get_pending_births(qor_pop *p) { return (int64_t)(p->birthList.size() / 3); }

int qor_set_birth_id_offset(qor_pop *p, int64_t offset, int64_t total) {
    p->birthIdOffset = offset;
    p->birthIdTotal = total;
    return 0;
}

// remove the live agents whose cell lies outside [c0, c1) and hand their records out
int64_t qor_extract_foreign(qor_pop *p, int c0, int c1, int64_t cap, int32_t *cell, int64_t *id, float *birth, uint8_t *gender,
                            float *age, float *last_birth, uint32_t *life) {
    int64_t k = 0;
    for (int i = 0; i < p->hi(); i++) {
        if (!p->active[i] || p->slots[i].life == LIFE_DEAD) continue;
        Agent &a = p->slots[i];
        if (a.cell >= c0 && a.cell < c1) continue;
        if (k < cap) {
            cell[k] = a.cell; id[k] = a.id; birth[k] = a.birth; gender[k] = a.gender; age[k] = a.age;
            last_birth[k] = a.lastBirth; life[k] = a.life;
            a.life = LIFE_DEAD;
            p->freeSlot(i);
        }
        k++;
    }
    p->updateNumAgentsPerCell();
    return k;
}

int qor_recount(qor_pop *p) { p->updateNumAgentsPerCell(); return 0; }

int64_t qor_get_max_id(qor_pop *p) { return p->maxID; }
int qor_set_max_id(qor_pop *p, int64_t v) { p->maxID = v; return 0; }

int qor_get_step_stats(qor_pop *p, uint64_t *births, uint64_t *deaths, uint64_t *moves) {
    if (births) *births = p->stepBirths;
    if (deaths) *deaths = p->stepDeaths;
    if (moves) *moves = p->stepMoves;
    return 0;
}

void qor_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox(ctr, key, out); }

int qor_well_sequence(const uint32_t *state16, int n, uint32_t *out) {
    Well512 w;
    w.seed(state16);
    for (int i = 0; i < n; i++) out[i] = w.next();
    return 0;
}

int qor_polyline_eval(const char *def, int n, const double *x, double *out, int float_cast) {
    PolyLine pl;
    if (!pl.parse(def)) return -1;
    for (int i = 0; i < n; i++) out[i] = float_cast ? pl.val((float)x[i]) : pl.val(x[i]);
    return 0;
}

}  // extern "C"
