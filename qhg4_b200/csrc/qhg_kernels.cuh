// qhg_kernels.cuh -- sm_100a kernels of the per-step agent update (device side only).
//
// Layout in HBM: agents are a structure of arrays kept BINNED BY CELL (cellStart[c]..cellStart[c+1]
// is the contiguous segment of cell c), double buffered; one step reads the current buffer and
// scatters survivors, movers and newborns into the other one (counting sort by destination cell).
// Per-cell arrays (neighbours, counts, b/d probabilities, cumulated weights) are small (a few MB)
// and stay L2 resident.  See DESIGN.md for the byte accounting of every kernel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "qhg_rng.cuh"

namespace qhg {

constexpr int MAXN = 6;            // core/SCell.h:4 MAX_NEIGH
constexpr int WSTRIDE = MAXN + 1;  // weight row: own cell + 6 neighbours (actions/SingleEvaluator.cpp:216-243)
constexpr int MAX_POLY = 16;
constexpr int MAX_OPS = 16;

// agent flag byte: gender and the FERTILE bit of the reference's life state (core/SPopulation.h:70-74)
constexpr uint8_t F_MALE = 1, F_FERTILE = 2, F_BORN = 4;  // F_BORN only in the per-step decision byte

enum Op : uint8_t { OP_GETOLD = 1, OP_ATANDEATH, OP_OLDAGEDEATH, OP_WEIGHTEDMOVE, OP_FERTILITY, OP_VERHULST, OP_DROWN, OP_NAVIGATE, OP_RANDOMMOVE, OP_WEIGHTEDMOVERAND, OP_SIGDEATH, OP_CONDWEIGHTEDMOVE };

struct AgentArrays {
    int64_t *id;
    float *birth;
    float *lastBirth;
    int32_t *cell;
    uint8_t *flags;
    float *age;  // only maintained when the action set does not refresh the age before it is read
    int32_t *gslot;    // genome row of the agent (populations with Genetics), else NULL
    int32_t *nbabies;  // m_iNumBabies (populations/OoANavGenPop.h:21-27), else NULL
};

struct BirthEntry {
    int babyPos;   // position of the newborn in the new buffer
    int mother;    // positions of the parents in the old buffer
    int father;
    int pad;
    long long cid; // id of the newborn
};

struct GenomeCtl {
    int nFree;     // genome rows on the free stack
    int hwm;       // rows ever used
    int nBirths;   // entries of the birth list this step
    int pad;
};

// called by the lanes of a warp that have a birth to record (a divergent branch): one atomic per warp, not per birth --
// all births of a step bump the same counter
__device__ __forceinline__ void record_birth(BirthEntry *births, GenomeCtl *ctl, int babyPos, int mother, int father, long long cid) {
    BirthEntry e;
    e.babyPos = babyPos; e.mother = mother; e.father = father; e.pad = 0; e.cid = cid;
    const unsigned m = __activemask();
    const int leader = __ffs(m) - 1, lane = (int)(threadIdx.x & 31);
    int base = 0;
    if (lane == leader) base = atomicAdd(&ctl->nBirths, __popc(m));
    base = __shfl_sync(m, base, leader);
    unsigned lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
    births[base + __popc(m & lt)] = e;
}

struct DevStats {
    int nAgents;   // live agents in the current buffer
    int nNew;      // live agents after the running step
    int nBirths, nDeaths, nMoves;
    int overflow;
    unsigned step;  // counter word of the random streams; +1 per finalizeStep
    int oversize;   // a cell did not fit the fast path: the host reruns the step on the generic path
    int workDecide, workScatter;  // dynamic work counters of the two fast-path passes
    long long nextID;
    // sharded runs over peer memory (qhgb_comm_p2p_connect)
    int commError;                // a peer did not reach a cross-GPU barrier in time, or the migrant buffer overflowed
    int nSent, nRecv;             // agents that left for / arrived from other ranks in the running step
    long long birthOffset;        // births of the lower ranks this step (newborn ids are global ranks)
    long long globalBirths;       // births of all ranks this step
    // several steps queued without a host round trip (qhgb_run): a step that cannot complete on the fast path (overflow,
    // oversize cell) raises `halt`; every kernel of the later steps then does nothing, and the host, when it finally looks,
    // finds the state as it was before that step (`step` tells which one) and redoes it the slow way
    int halt;
    int pad0;
    int doneBlocks;               // blocks of the step's last kernel that have finished (the last one does the step's book-keeping)
    int pad1;
    long long agentSteps;         // sum over the completed steps of the live agents at step start
    long long totSent, totRecv;   // agents sent to / received from other ranks, summed over the completed steps
};

struct ActParams {
    int nOps;
    unsigned long long prog;  // the action program: 4 bits per op, first op in the low nibble
    float t;
    int storeAge;
    // ATanDeath (actions/ATanDeath.cpp:49-59,66-90)
    double atanMaxAge, atanSlope, atanScale, atanXlo, atanXhi;
    float atanAgeLo, atanAgeHi;  // float ages strictly outside [lo, hi] are outside [Xlo, Xhi] (fast path pre-test)
    // OldAgeDeath (actions/OldAgeDeath.cpp:48-67)
    double oadMaxAge, oadLo, oadHi;
    // WeightedMove (actions/WeightedMove.cpp:45-106)
    double moveProb;
    unsigned long long tMove;  // ceil(moveProb * 2^32): "32-bit draw / 2^32 < moveProb" as an exact integer test
    // Fertility (actions/Fertility.cpp:49-74)
    float fertMinAge, fertMaxAge, fertInterbirth;
    RngKey key;
    RoundKeys rk;  // the ten Philox round keys of `key` (read straight from the constant bank by the fast path)
    // tut_ParthenoPop: no pairing action, every female counts as mated (actions/LinearBirth.cpp:139-142 only tests
    // m_iMateIndex >= 0; populations/tut_ParthenoPop.cpp:107-119 gives every newborn its own index) and newborns are female
    int selfMate;
    // ConfinedMove is active (actions/ConfinedMove.cpp:86-101): a move into a cell outside the region becomes a move to the
    // cell it starts from (still counted)
    int confine;
    // WeightedMoveRand (actions/WeightedMoveRand.cpp) and SigDeath (actions/SigDeath.cpp:49-90), generic path only
    double moveProbRand, sigMaxAge, sigSlope, sigScale;
    // CondWeightedMove (actions/CondWeightedMove.cpp) with a SimpleCondition over the altitudes (actions/SimpleCondition.cpp:
    // 0 never, 1 always, 2 greater, 3 less, 4 equal, 5 greater or equal, 6 less or equal, 7 different)
    int condMode;
};

// SimpleCondition::allow (actions/SimpleCondition.cpp:12-19,73-77): the candidate's value enters scaled by 0.2, and "less"
// also asks for a scaled value below 10
__device__ __forceinline__ bool cond_allow(int mode, double cur, double cand) {
    const double n = __dmul_rn(0.2, cand);
    switch (mode) {
    case 1: return true;
    case 2: return n > cur;
    case 3: return (n < 10.0) && (n < cur);
    case 4: return n == cur;
    case 5: return n >= cur;
    case 6: return n <= cur;
    case 7: return n != cur;
    default: return false;
    }
}

__host__ __device__ __forceinline__ int prog_op(const ActParams &P, int k) { return (int)((P.prog >> (4 * k)) & 15ull); }

struct PolyLineDev {  // utils/PolyLine.cpp:60-89
    int nseg;  // 0 => identity
    double x[MAX_POLY], v[MAX_POLY], a[MAX_POLY];
};

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// counter[key] += (number of lanes with this key); returns this lane's rank inside the cell.
// Agents are binned by cell, so a warp usually touches 1-3 distinct keys: one atomic per key, not per lane.
__device__ __forceinline__ int warp_agg_inc(int *counter, int key, bool active) {
    unsigned act = __ballot_sync(0xffffffffu, active);
    int r = 0;
    if (active) {
        unsigned peers = __match_any_sync(act, key);
        int leader = __ffs(peers) - 1;
        int base = 0;
        if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(&counter[key], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        r = base + __popc(peers & lanemask_lt());
    }
    return r;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// per-cell: Verhulst birth and death probabilities from last step's counts
// (actions/LinearBirth.cpp:97-112, actions/LinearDeath.cpp:101-119), and reset of the step's counters
__global__ void k_cell_init(DevStats *__restrict__ st, int cLo, int cHi, const int *__restrict__ count, double *__restrict__ B, double *__restrict__ D,
                            double b0, double d0, double theta, double K, const double *__restrict__ Kcell, int doVerhulst,
                            int *__restrict__ stay, int *__restrict__ arrive, int *__restrict__ cursor,
                            int *__restrict__ birthCount, int *__restrict__ nFert, unsigned long long *__restrict__ TB,
                            unsigned long long *__restrict__ TD) {
    if (st->halt) return;  // an earlier queued step failed: leave everything as it is
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // the step's tallies and work counters start from zero
        st->nBirths = 0;
        st->nDeaths = 0;
        st->nMoves = 0;
        st->nNew = 0;
        st->oversize = 0;
        st->workDecide = 0;
        st->workScatter = 0;
        st->nSent = 0;
        st->nRecv = 0;
        st->birthOffset = 0;
    }
    for (int c = cLo + blockIdx.x * blockDim.x + threadIdx.x; c < cHi; c += gridDim.x * blockDim.x) {  // the cells this GPU owns
        if (doVerhulst) {
            const double Kc = Kcell ? Kcell[c] : K;  // VerhulstVarK: the carrying capacity of the cell (actions/VerhulstVarK.cpp)
            if (Kcell && Kc <= 0) {               // LinearBirth.cpp:104-105, LinearDeath.cpp:113-114
                B[c] = 0;
                D[c] = 1;
            } else {
                double q = __ddiv_rn((double)count[c], Kc);
                B[c] = __dadd_rn(b0, __dmul_rn(__dadd_rn(theta, -b0), q));
                D[c] = __dadd_rn(d0, __dmul_rn(__dadd_rn(theta, -d0), q));
            }
            // the fast path's probability tests as exact integer thresholds on the 32-bit draws, once per cell:
            // TB = threshold of |b| (33 bits) with bit 62 = (b > 0), bit 63 = (b < 0); TD = threshold of d
            const double b = B[c];
            unsigned long long tb = prob_threshold(b > 0 ? b : -b);
            if (tb > (1ull << 32)) tb = 1ull << 32;  // |b| >= 1: every draw is below it
            TB[c] = tb | (b > 0 ? (1ull << 62) : 0ull) | (b < 0 ? (1ull << 63) : 0ull);
            TD[c] = prob_threshold(D[c]);
        }
        stay[c] = 0;
        arrive[c] = 0;
        cursor[c] = 0;
        birthCount[c] = 0;
        nFert[2 * c] = 0;
        nFert[2 * c + 1] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// cell weights: SingleEvaluator::calcValues (actions/SingleEvaluator.cpp:174-207) ...
__device__ __forceinline__ double polyline_val(const PolyLineDev &pl, double fx) {
    if (pl.nseg == 0) return fx;
    if (fx >= pl.x[pl.nseg]) return pl.v[pl.nseg];
    int i = 0;
    while (i <= pl.nseg && fx > pl.x[i]) i++;
    if (i == 0) return pl.v[0];
    if (i <= pl.nseg) return __dadd_rn(pl.v[i - 1], __dmul_rn(pl.a[i - 1], __dadd_rn(fx, -pl.x[i - 1])));
    return pl.v[pl.nseg];
}

__global__ void k_weights_own(int nCells, const double *__restrict__ in, const uint8_t *__restrict__ ice,
                              PolyLineDev pl, int usePoly, double *__restrict__ W) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nCells; c += gridDim.x * blockDim.x) {
        double v = 0;
        if (!ice || !ice[c]) {
            double dv = usePoly ? polyline_val(pl, (double)(float)in[c]) : in[c];  // the (float) cast is the reference's, :185
            v = (dv > 0) ? dv : 0;
        }
        W[(size_t)c * WSTRIDE] = v;
    }
}

// ... and exchangeAndCumulate (:216-243): row c = running sum over [own, n1..n6], missing neighbours count 0
__global__ void k_weights_cumulate(int nCells, const int *__restrict__ nbr, double *__restrict__ W, int cumulate) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nCells; c += gridDim.x * blockDim.x) {
        double w = W[(size_t)c * WSTRIDE];
#pragma unroll
        for (int k = 0; k < MAXN; k++) {
            int n = nbr[(size_t)c * MAXN + k];
            double cw = (n >= 0) ? W[(size_t)n * WSTRIDE] : 0.0;
            cw = (cw > 0) ? cw : 0;
            w = cumulate ? __dadd_rn(w, cw) : cw;
            W[(size_t)c * WSTRIDE + k + 1] = w;
        }
    }
}

// MultiEvaluator::addSingleWeights (actions/MultiEvaluator.cpp:221-253): out += single * weight for every evaluator ...
__global__ void k_multi_accumulate(size_t n, const double *__restrict__ single, double weight, double *__restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = __dadd_rn(out[i], __dmul_rn(single[i], weight));
}
// the other combine modes of MultiEvaluator (actions/MultiEvaluator.h:19-26; actions/MultiEvaluator.cpp:263-298 ADD_BLOCK,
// 308-342 MUL_SIMPLE, 350-381 MAX_SIMPLE, 391-432 MAX_BLOCK, 440-478 MIN_SIMPLE): one elementwise pass per evaluator over the
// nCells x 7 weight rows; `allowed` (BLOCK modes, else NULL) masks the entries findBlockings ruled out
enum MultiMode : int { MM_ADD_SIMPLE = 0, MM_ADD_BLOCK = 1, MM_MUL_SIMPLE = 2, MM_MAX_SIMPLE = 3, MM_MAX_BLOCK = 4, MM_MIN_SIMPLE = 5 };
__global__ void k_multi_combine(size_t n, int mode, const double *__restrict__ single, double weight,
                                const uint8_t *__restrict__ allowed, double *__restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (allowed && !allowed[i]) continue;
        const double v = __dmul_rn(single[i], weight);
        const double o = out[i];
        if (mode == MM_ADD_SIMPLE || mode == MM_ADD_BLOCK) out[i] = __dadd_rn(o, v);
        else if (mode == MM_MUL_SIMPLE) out[i] = __dmul_rn(o, v);
        else if (mode == MM_MAX_SIMPLE || mode == MM_MAX_BLOCK) { if (o < v) out[i] = v; }
        else { if (o > v) out[i] = v; }
    }
}
__global__ void k_fill_f64(size_t n, double v, double *__restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}
// MultiEvaluator::findBlockings (actions/MultiEvaluator.cpp:579-598): an entry where ANY evaluator is <= 0 is blocked
__global__ void k_find_blockings(size_t n, const double *__restrict__ single, uint8_t *__restrict__ allowed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (single[i] <= 0) allowed[i] = 0;
}
// ... then the rows are cumulated (again: the evaluators inside were built cumulating, docs/DoubleCumulateArtifactsBug.odt)
__global__ void k_rows_cumulate(int nCells, double *__restrict__ W) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nCells; c += gridDim.x * blockDim.x) {
        double w = W[(size_t)c * WSTRIDE];
#pragma unroll
        for (int k = 1; k < WSTRIDE; k++) {
            w = __dadd_rn(W[(size_t)c * WSTRIDE + k], w);
            W[(size_t)c * WSTRIDE + k] = w;
        }
    }
}

// AutoInterpolator::interpolate (core/AutoInterpolator.cpp:461-483): target += steps * diff, one multiply and one add per
// cell, rounded separately like the reference's compiled loop (no fused multiply-add)
__global__ void k_env_interpolate(int n, double steps, const double *__restrict__ diff, double *__restrict__ target) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        target[i] = __dadd_rn(target[i], __dmul_rn(steps, diff[i]));
}

// NPPCapacity::recalculate (actions/NPPCapacity.cpp:138-217) with NPPCalcMiami::calcNPP (core/NPPCalcMiami.cpp:28-42)
struct NppParams {
    double waterFactor, coastalFactor, coastMinLat, coastMaxLat, nppMin, nppMax, kMax, kMin, efficiency;
};
__global__ void k_npp_capacity(int nCells, NppParams Q, const double *__restrict__ T, const double *__restrict__ Pr,
                               const double *__restrict__ water, const double *__restrict__ npp, const double *__restrict__ alt,
                               const double *__restrict__ lon, const double *__restrict__ lat, const double *__restrict__ coastal,
                               double *__restrict__ cap) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nCells; i += gridDim.x * blockDim.x) {
        const double nppT = __ddiv_rn(3000.0, __dadd_rn(1.0, exp_rn(__dadd_rn(1.315, -__dmul_rn(0.119, T[i])))));
        const double nppP = __dmul_rn(3000.0, __dadd_rn(1.0, -exp_rn(__dmul_rn(-0.000664, Pr[i]))));
        const double miami = __dmul_rn(0.000475, (nppT < nppP) ? nppT : nppP);  // GDM_TO_KGC
        double tc;
        const double a = alt[i];
        if (a > 0) {
            double af = (a < 1500.0) ? 1.0 : __ddiv_rn(__dadd_rn(2500.0, -a), 1000.0);
            if (af < 0) af = 0;
            double tn = npp[i];
            if (lon[i] > 115.0 && lat[i] > -12.0 && lon[i] < 150.0 && lat[i] < 1.0) {  // Oceania box (:22-25)
                if (npp[i] < Q.nppMin) tn = miami;
            }
            tn = __dmul_rn(tn, af);
            if (tn < Q.nppMin) tc = Q.kMin;
            else if (tn > Q.nppMax) tc = Q.kMax;
            else tc = __dadd_rn(Q.kMin, __ddiv_rn(__dmul_rn(tn, Q.kMax), __dadd_rn(Q.nppMax, -Q.nppMin)));
            if (coastal[i] != 0 && lat[i] > Q.coastMinLat && lat[i] < Q.coastMaxLat) tc = __dadd_rn(tc, __dmul_rn(Q.coastalFactor, Q.kMax));
            tc = __dadd_rn(tc, __dmul_rn(__dmul_rn(water[i], Q.waterFactor), Q.kMax));
            if (tc > Q.kMax) tc = Q.kMax;
        } else {
            tc = 0;
        }
        cap[i] = __dmul_rn(tc, Q.efficiency);
    }
}

// ---------------------------------------------------------------------------------------------
// pairing (RandomPair::initialize/findMates, actions/RandomPair.cpp:107-128,146-279), counter-mode law:
// inside a cell the fertile females and the fertile males are ranked by (random key, id); equal ranks mate.
__global__ void k_pair_keys(const DevStats *__restrict__ st, AgentArrays a, RngKey key, uint32_t *__restrict__ pkey,
                            int *__restrict__ mate, int *__restrict__ nFert) {
    const int n = st->nAgents;
    const unsigned step = st->step;
    for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
        int i = i0 + threadIdx.x;
        bool fert = false;
        int slot = 0;
        if (i < n) {
            uint8_t f = a.flags[i];
            mate[i] = -3;
            fert = (f & F_FERTILE) != 0;
            if (fert) {
                pkey[i] = agent_draws(a.id[i], step, STREAM_PAIR, key).x;
                slot = 2 * a.cell[i] + (f & F_MALE);
            }
        }
        warp_agg_inc(nFert, slot, fert);
    }
}

__global__ void k_pair_rank(const DevStats *__restrict__ st, AgentArrays a, const int *__restrict__ cellStart,
                            const uint32_t *__restrict__ pkey, int *__restrict__ prank, int *__restrict__ ranked) {
    const int n = st->nAgents;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint8_t f = a.flags[i];
        if (!(f & F_FERTILE)) continue;
        int c = a.cell[i];
        int s = cellStart[c], e = cellStart[c + 1];
        uint32_t k = pkey[i];
        int64_t id = a.id[i];
        int r = 0;
        for (int j = s; j < e; j++) {
            uint8_t fj = a.flags[j];
            if ((fj & (F_FERTILE | F_MALE)) != (f & (F_FERTILE | F_MALE))) continue;
            uint32_t kj = pkey[j];
            if (kj < k || (kj == k && a.id[j] < id)) r++;
        }
        prank[i] = r;
        if (f & F_MALE) ranked[e - 1 - r] = i; else ranked[s + r] = i;
    }
}

__global__ void k_pair_match(const DevStats *__restrict__ st, AgentArrays a, const int *__restrict__ cellStart,
                             const int *__restrict__ nFert, const int *__restrict__ prank, const int *__restrict__ ranked,
                             int *__restrict__ mate) {
    const int n = st->nAgents;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint8_t f = a.flags[i];
        if (!(f & F_FERTILE)) continue;
        int c = a.cell[i];
        int np = min(nFert[2 * c], nFert[2 * c + 1]);
        int r = prank[i];
        if (r < np) mate[i] = (f & F_MALE) ? ranked[cellStart[c] + r] : ranked[cellStart[c + 1] - 1 - r];
    }
}

// ---------------------------------------------------------------------------------------------
// per-cell read-only data the actions touch (L2 resident)
// MoveStats (actions/MoveStats.cpp): per cell, hops / distance / time of the arrival that counts under MoveStats_Mode (0 the
// first, 1 the minimum, 2 the last).  The reference walks the step's move list; the device knows no order of the agents, so
// "first" is the move of the agent with the smallest id and "last" that of the largest (the oracle's counter mode).
// Every registered move leaves one atomic in the per-step arrays; k_move_stats_temp / k_move_stats_merge apply the step.
struct MoveStatsDev {
    int mode;
    int *hops; double *dist, *time;          // m_aiHops, m_adDist, m_adTime
    int *hopsT; double *distT, *timeT;       // the reference's per-thread arrays (one thread): never reset
    unsigned long long *key;                 // modes 0 / 2: (agent id << 20 | source cell) of the move that counts, +1
    int *stepHops; unsigned long long *stepDist;  // mode 1: minima over this step's moves (the distance as ordered bits)
    uint8_t *changed;
    const double *lon, *lat;
};
constexpr int MS_CELL_BITS = 20;

// utils/geomutils.cpp:309-333 (spherdistDeg -> spherdist) with the radius of the Geography (6371.3 km)
__device__ __forceinline__ double ms_distance(const MoveStatsDev &M, int from, int to) {
    const double PI_ = 3.14159265358979323846;
    const double lo1 = __ddiv_rn(__dmul_rn(M.lon[from], PI_), 180.0), la1 = __ddiv_rn(__dmul_rn(M.lat[from], PI_), 180.0);
    const double lo2 = __ddiv_rn(__dmul_rn(M.lon[to], PI_), 180.0), la2 = __ddiv_rn(__dmul_rn(M.lat[to], PI_), 180.0);
    const double x1 = __dmul_rn(cos(lo1), cos(la1)), y1 = __dmul_rn(sin(lo1), cos(la1)), z1 = sin(la1);
    const double x2 = __dmul_rn(cos(lo2), cos(la2)), y2 = __dmul_rn(sin(lo2), cos(la2)), z2 = sin(la2);
    double pr = __dadd_rn(__dadd_rn(__dmul_rn(x1, x2), __dmul_rn(y1, y2)), __dmul_rn(z1, z2));
    if (pr > 1) pr = 1; else if (pr < -1) pr = -1;
    return __dmul_rn(6371.3, acos(pr));
}

// one registered move (SPopulation::registerMove) as MoveStats::finalize will see it in the move list
__device__ __forceinline__ void ms_register(const MoveStatsDev *M, int from, int to, long long id) {
    if (!M) return;
    const unsigned long long k = (((unsigned long long)id << MS_CELL_BITS) | (unsigned long long)from) + 1ull;
    if (M->mode == 0) atomicMin(&M->key[to], k);
    else if (M->mode == 2) atomicMax(&M->key[to], k);
    else {
        atomicMin(&M->stepHops[to], M->hops[from] + 1);
        atomicMin(&M->stepDist[to], (unsigned long long)__double_as_longlong(__dadd_rn(M->dist[from], ms_distance(*M, from, to))));
    }
}

// MoveStats::finalize, first half (actions/MoveStats.cpp:204-241): the step's moves into every cell update the Temp arrays --
// all values come from the arrays as they stood BEFORE the step
__global__ void k_move_stats_temp(MoveStatsDev M, int nCells, double t) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nCells; c += gridDim.x * blockDim.x) {
        if (M.mode == 1) {
            const int h = M.stepHops[c];
            if (h == 0x7fffffff) continue;
            const double d = __longlong_as_double((long long)M.stepDist[c]);
            if (M.hopsT[c] < 0) { M.hopsT[c] = h; M.distT[c] = d; }
            else { if (h < M.hopsT[c]) M.hopsT[c] = h; if (d < M.distT[c]) M.distT[c] = d; }
            M.timeT[c] = t;
            M.changed[c] = 1;
            M.stepHops[c] = 0x7fffffff; M.stepDist[c] = ~0ull;
        } else {
            const unsigned long long none = (M.mode == 0) ? ~0ull : 0ull;
            const unsigned long long k = M.key[c];
            if (k == none) continue;
            M.key[c] = none;
            if (M.hopsT[c] < 0 || M.mode == 2) {
                const int from = (int)((k - 1ull) & ((1ull << MS_CELL_BITS) - 1ull));
                M.hopsT[c] = M.hops[from] + 1;
                M.distT[c] = __dadd_rn(M.dist[from], ms_distance(M, from, c));
                M.timeT[c] = t;
                M.changed[c] = 1;
            }
        }
    }
}

// ... second half (:249-283): the cells touched in this step take over what their Temp entries say
__global__ void k_move_stats_merge(MoveStatsDev M, int nCells) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nCells; c += gridDim.x * blockDim.x) {
        if (!M.changed[c]) continue;
        M.changed[c] = 0;
        if (M.hops[c] < 0 || M.mode == 2) {
            M.hops[c] = M.hopsT[c]; M.dist[c] = M.distT[c]; M.time[c] = M.timeT[c];
        } else if (M.mode == 1) {
            if (M.hopsT[c] < M.hops[c]) M.hops[c] = M.hopsT[c];
            if (M.distT[c] < M.dist[c]) M.dist[c] = M.distT[c];
            if (M.timeT[c] < M.time[c]) M.time[c] = M.timeT[c];
        }
    }
}

// MoveStats::preLoop + initializeOccupied (actions/MoveStats.cpp:107-186): -1 everywhere, 0 in the cells that are occupied
__global__ void k_move_stats_init(MoveStatsDev M, int nCells, const int *__restrict__ count) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nCells; c += gridDim.x * blockDim.x) {
        const bool occ = count[c] > 0;
        M.hops[c] = occ ? 0 : -1; M.dist[c] = occ ? 0.0 : -1.0; M.time[c] = occ ? 0.0 : -1.0;
        M.hopsT[c] = -1; M.distT[c] = -1.0; M.timeT[c] = -1.0;
        M.key[c] = (M.mode == 0) ? ~0ull : 0ull;
        M.stepHops[c] = 0x7fffffff; M.stepDist[c] = ~0ull;
        M.changed[c] = 0;
    }
}

struct CellEnv {
    const int *nbr;
    const uint8_t *nNbr;
    const uint8_t *ice;   // may be NULL
    const double *alt;
    const double *W;
    const double *B;
    const double *D;
    const unsigned long long *TB, *TD;  // the same as integer thresholds (k_cell_init), read by the fast path
    // Navigate (actions/Navigate.cpp): ports as CSR (every port has n+1 entries: 0 = stay at home), current bridges
    const int *navRow;       // per cell: port index or -1; NULL if the population does not navigate
    const int *navPtr;       // first entry of port p
    const int *navDest;      // destination cell of every entry (-1 for entry 0)
    const double *navCum;    // cumulated jump probability of every entry
    const int2 *bridges;
    int nBridges;
    double bridgeProb;
    const uint8_t *allowed;  // ConfinedMove::m_bAllowed (actions/ConfinedMove.cpp:44-78), NULL without the action
    const MoveStatsDev *ms;  // MoveStats is active in this step (generic path only), else NULL
};

struct Decision {
    bool alive, born, moving;
    bool movingBit;  // LIFE_STATE_MOVING as later actions see it (Fertility overwrites the whole life state)
    int nMoves;      // registered moves (every registerMove counts, the last one wins: core/SPopulation.cpp:1058-1092)
    int pick;    // 0: stays, 1..6: moves to neighbour slot pick-1
    int to;      // destination cell
    uint8_t f;   // new flag byte (gender | fertile)
    float age;
};

// every enabled action in priority order for ONE agent (core/SPopulation.cpp:554-577 runs them as separate passes
// over all agents; an action only touches its own agent and queues births/deaths/moves, so running them back to
// back per agent in the same order is equivalent).  A dead agent skips the remaining actions (:568).
__device__ __forceinline__ Decision run_actions(const ActParams &P, const CellEnv &E, unsigned step, int64_t id, float birth,
                                                float ageIn, int c, uint8_t f, bool hasMate, const float *lastBirthPtr) {
    Decision d;
    d.alive = true; d.born = false; d.moving = false; d.movingBit = false; d.nMoves = 0; d.pick = 0; d.to = c; d.f = f; d.age = ageIn;
    uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
    bool have0 = false, have1 = false;
#pragma unroll 1
    for (int k = 0; k < P.nOps; k++) {
        if (!d.alive) break;
        switch (prog_op(P, k)) {
        case OP_GETOLD:  // actions/GetOld.cpp:37-48
            d.age = __fsub_rn(P.t, birth);
            break;
        case OP_ATANDEATH: {  // actions/ATanDeath.cpp:66-90
            d.age = __fsub_rn(P.t, birth);
            double x = __dmul_rn(P.atanSlope, __dadd_rn((double)d.age, -P.atanMaxAge));
            if (x > P.atanXlo) {  // below Xlo the probability is negative: nobody dies, no draw needed
                if (!have0) { r0 = agent_draws(id, step, STREAM_ACT0, P.key); have0 = true; }
                bool dies = true;  // above Xhi the probability exceeds 1
                if (x < P.atanXhi) {
                    double p = __dadd_rn(0.5, __ddiv_rn(__dmul_rn(P.atanScale, atan_rn(x)), 3.141592653589793));
                    dies = u2d(r0.x) < p;
                }
                if (dies) d.alive = false;
            }
            break;
        }
        case OP_OLDAGEDEATH: {  // actions/OldAgeDeath.cpp:48-67
            d.age = __fsub_rn(P.t, birth);
            if (!have1) { r1 = agent_draws(id, step, STREAM_ACT1, P.key); have1 = true; }
            double r = u2range(r1.w, P.oadLo, P.oadHi);
            if ((double)d.age > __dadd_rn(P.oadMaxAge, r)) d.alive = false;
            break;
        }
        case OP_WEIGHTEDMOVE: {  // actions/WeightedMove.cpp:45-106
            if (!have0) { r0 = agent_draws(id, step, STREAM_ACT0, P.key); have0 = true; }
            if (u2d(r0.y) < P.moveProb) {
                if (!have1) { r1 = agent_draws(id, step, STREAM_ACT1, P.key); have1 = true; }
                const int nreal = E.nNbr[c];
                const double *row = E.W + (size_t)c * WSTRIDE;
                int pick = -1;
                const double wmax = row[nreal];
                if (row[0] == wmax) {
                    pick = (int)u2int(r1.x, 0, nreal + 1);
                } else {
                    double r2 = __dmul_rn(u2d(r1.x), wmax);
                    for (int q = 0; q < nreal + 1; q++) {
                        if (r2 < row[q]) { pick = q; break; }
                    }
                }
                if (pick > 0) {
                    int dst = E.nbr[(size_t)c * MAXN + pick - 1];
                    if (dst >= 0 && !(E.ice && E.ice[dst])) { d.to = dst; d.pick = pick; d.moving = true; d.movingBit = true; d.nMoves++; ms_register(E.ms, c, dst, id); }
                }
            }
            break;
        }
        case OP_CONDWEIGHTEDMOVE: {  // actions/CondWeightedMove.cpp:41-86: the whole row up to the grid's connectivity, no special
            // case for equal weights, the ice test looks at the cell the agent is IN, the MoveCondition decides last
            if (!have0) { r0 = agent_draws(id, step, STREAM_ACT0, P.key); have0 = true; }
            if (u2d(r0.y) < P.moveProb) {
                if (!have1) { r1 = agent_draws(id, step, STREAM_ACT1, P.key); have1 = true; }
                const double *row = E.W + (size_t)c * WSTRIDE;
                int pick = -1;
                const double r2 = __dmul_rn(u2d(r1.x), row[MAXN]);
                for (int q = 0; q < MAXN + 1; q++) {
                    if (r2 < row[q]) { pick = q; break; }
                }
                if (pick > 0) {
                    const int dst = E.nbr[(size_t)c * MAXN + pick - 1];
                    if (dst >= 0 && !(E.ice && E.ice[c]) && cond_allow(P.condMode, E.alt[c], E.alt[dst])) {
                        d.to = dst; d.pick = pick; d.moving = true; d.movingBit = true; d.nMoves++; ms_register(E.ms, c, dst, id);
                    }
                }
            }
            break;
        }
        case OP_WEIGHTEDMOVERAND: {  // actions/WeightedMoveRand.cpp:43-100: the weights decide unless they are all zero
            if (!have0) { r0 = agent_draws(id, step, STREAM_ACT0, P.key); have0 = true; }
            if (u2d(r0.y) < P.moveProbRand) {
                if (!have1) { r1 = agent_draws(id, step, STREAM_ACT1, P.key); have1 = true; }
                const int nreal = E.nNbr[c];
                const double *row = E.W + (size_t)c * WSTRIDE;
                int pick = -1;
                if (row[MAXN] > 0) {
                    const double r2 = __dmul_rn(u2d(r1.x), row[MAXN]);
                    for (int q = 0; q < MAXN + 1; q++) {
                        if (r2 < row[q]) { pick = q; break; }
                    }
                } else {
                    pick = (int)u2range(r1.x, 0.0, (double)(nreal + 1));  // (int) wrandr(0, iNumActualNeigh+1)
                }
                if (pick > 0) {
                    int dst = E.nbr[(size_t)c * MAXN + pick - 1];
                    if (dst >= 0 && !(E.ice && E.ice[dst])) { d.to = dst; d.pick = pick; d.moving = true; d.movingBit = true; d.nMoves++; ms_register(E.ms, c, dst, id); }
                }
            }
            break;
        }
        case OP_SIGDEATH: {  // actions/SigDeath.cpp:66-90
            d.age = __fsub_rn(P.t, birth);
            if (!have1) { r1 = agent_draws(id, step, STREAM_ACT1, P.key); have1 = true; }
            const double x = __dmul_rn(-P.sigSlope, __dadd_rn((double)d.age, -P.sigMaxAge));
            const double p = __ddiv_rn(P.sigScale, __dadd_rn(1.0, exp_rn(x)));
            if (u2d(r1.z) < p) d.alive = false;
            break;
        }
        case OP_RANDOMMOVE: {  // actions/RandomMove.cpp:65-100: direction = (int)(r2 * (neighbours + 1)), 0 = stay; no ice test
            if (!have0) { r0 = agent_draws(id, step, STREAM_ACT0, P.key); have0 = true; }
            if (u2d(r0.y) < P.moveProb) {
                if (!have1) { r1 = agent_draws(id, step, STREAM_ACT1, P.key); have1 = true; }
                const int pick = (int)__dmul_rn(u2d(r1.x), (double)(E.nNbr[c] + 1));
                if (pick > 0) {
                    int dst = E.nbr[(size_t)c * MAXN + pick - 1];
                    if (dst >= 0) { d.to = dst; d.pick = pick; d.moving = true; d.movingBit = true; d.nMoves++; ms_register(E.ms, c, dst, id); }
                }
            }
            break;
        }
        case OP_FERTILITY: {  // actions/Fertility.cpp:49-74
            bool fert;
            if (!(d.f & F_MALE)) {
                fert = (d.age > P.fertMinAge) && (d.age < P.fertMaxAge) && (__fsub_rn(P.t, *lastBirthPtr) > P.fertInterbirth);
            } else {
                fert = d.age > P.fertMinAge;
            }
            d.f = (uint8_t)((d.f & F_MALE) | (fert ? F_FERTILE : 0));
            d.movingBit = false;  // the life state is overwritten, MOVING bit included (actions/Fertility.cpp:56-68)
            break;
        }
        case OP_NAVIGATE: {  // actions/Navigate.cpp:181-250
            if (d.movingBit || !E.navRow) break;
            const int port = E.navRow[c];
            if (port >= 0) {
                if (!have1) { r1 = agent_draws(id, step, STREAM_ACT1, P.key); have1 = true; }
                const int p0 = E.navPtr[port], nd = E.navPtr[port + 1] - p0 - 1;
                const int lim = (c < nd) ? c : nd;  // the reference bounds the search by the port's cell index (:194); entries end at nd
                const double r = u2d(r1.y);
                int i = 0;
                while (i < lim && r > E.navCum[p0 + i]) i++;
                if (i > 0) {
                    const int dst = E.navDest[p0 + i];
                    if (!(E.ice && E.ice[dst])) { d.to = dst; d.pick = 0; d.moving = true; d.movingBit = true; d.nMoves++; ms_register(E.ms, c, dst, id); }
                }
            }
            for (int b = 0; b < E.nBridges; b++) {  // manual bridges: one draw per incident bridge (:228-247)
                const int2 br = E.bridges[b];
                const int dst = (br.x == c) ? br.y : ((br.y == c) ? br.x : -1);
                if (dst >= 0) {
                    const uint4 db = agent_draws(id, step, 0x04000000u | (unsigned)(b / 4), P.key);
                    const unsigned wv = (b & 3) == 0 ? db.x : (b & 3) == 1 ? db.y : (b & 3) == 2 ? db.z : db.w;
                    if (u2d(wv) < E.bridgeProb) { d.to = dst; d.pick = 0; d.moving = true; d.movingBit = true; d.nMoves++; ms_register(E.ms, c, dst, id); }
                }
            }
            break;
        }
        case OP_VERHULST: {  // actions/Verhulst.cpp:101-115 -> LinearBirth.cpp:122-168, LinearDeath.cpp:131-153
            if (!have0) { r0 = agent_draws(id, step, STREAM_ACT0, P.key); have0 = true; }
            const double b = E.B[c];
            if (b > 0) {
                if (!(d.f & F_MALE) && hasMate) {
                    if (u2d(r0.z) < b) d.born = true;
                }
            } else if (b < 0) {
                if (u2d(r0.z) < -b) d.alive = false;
            }
            if (d.alive && u2d(r0.w) < E.D[c]) d.alive = false;
            break;
        }
        case OP_DROWN:  // populations/tut_EnvironAltPop.cpp:100-116 (EVENT_ID_GEO)
            if (E.alt[c] < 0 || (E.ice && E.ice[c])) d.alive = false;
            break;
        }
    }
    // ConfinedMove::finalize runs in finalizeStep over the whole move list, whatever the action's priority
    // (core/SPopulation.cpp:445-455, actions/ConfinedMove.cpp:86-101): the last registered move decides where the agent ends up
    if (P.confine && d.moving && !E.allowed[d.to]) { d.to = c; d.pick = 0; }
    return d;
}

// ---------------------------------------------------------------------------------------------
// generic path, pass 1: one thread per agent, pairing read from `mate`.  Output per agent: destination cell
// (-1 = dead), rank inside the destination cell, new flag byte.  Used for cells too large for the tiled path.
__global__ void __launch_bounds__(256)
k_actions(DevStats *__restrict__ st, AgentArrays a, const int *__restrict__ mate, ActParams P, CellEnv E,
          int *__restrict__ arrive, int *__restrict__ birthCount,
          int *__restrict__ dest, int *__restrict__ rank, uint8_t *__restrict__ oflags) {
    const int n = st->nAgents;
    const unsigned step = st->step;
    int nDead = 0, nMove = 0, nBorn = 0;
    for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + threadIdx.x;
        const bool valid = i < n;
        Decision d;
        d.alive = false; d.born = false; d.moving = false; d.to = 0; d.f = 0;
        int c = 0;
        if (valid) {
            c = a.cell[i];
            bool needMate = false;
            for (int k = 0; k < P.nOps; k++) needMate |= (prog_op(P, k) == OP_VERHULST);
            const uint8_t f = a.flags[i];
            const bool hasMate = needMate && !(f & F_MALE) && (P.selfMate || mate[i] >= 0);
            d = run_actions(P, E, step, a.id[i], a.birth[i], P.storeAge ? a.age[i] : 0.0f, c, f, hasMate, a.lastBirth + i);
            if (P.storeAge && d.alive) a.age[i] = d.age;  // moved with the agent by k_scatter
            nMove += d.nMoves;  // registered moves count even if the agent dies later in the step (core/SPopulation.cpp:1067)
            if (!d.alive) nDead++;
            if (d.born) nBorn++;
        }
        int r = warp_agg_inc(arrive, d.to, d.alive);
        warp_agg_inc(birthCount, c, d.born);
        if (valid) {
            dest[i] = d.alive ? d.to : -1;
            rank[i] = r;
            oflags[i] = (uint8_t)(d.f | (d.born ? F_BORN : 0));
        }
    }
    nDead = warp_sum(nDead); nMove = warp_sum(nMove); nBorn = warp_sum(nBorn);
    if ((threadIdx.x & 31) == 0) {
        if (nDead) atomicAdd(&st->nDeaths, nDead);
        if (nMove) atomicAdd(&st->nMoves, nMove);
        if (nBorn) atomicAdd(&st->nBirths, nBorn);
    }
}

// ---------------------------------------------------------------------------------------------
// exclusive scan over cells of (survivors + births) -> newStart, of births -> birthBase.
// Two kernels: per-tile sums, then every tile adds the sums of the tiles before it.
constexpr int SCAN_TILE = 2048;  // cells per block (256 threads x 8)

// The scan runs over the cells [cA, cHi): the rank's own range, its start rounded down to a multiple of 8 (the counters
// of cells owned by other ranks are zero), so that every thread's 8 cells are two aligned 128-bit words.
__device__ __forceinline__ void load8(const int *__restrict__ p, int c0, int cHi, int v[8]) {
    if (c0 + 8 <= cHi) {
        const int4 x = *reinterpret_cast<const int4 *>(p + c0), y = *reinterpret_cast<const int4 *>(p + c0 + 4);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (c0 + k < cHi) ? p[c0 + k] : 0;
    }
}
__device__ __forceinline__ void store8(int *__restrict__ p, int c0, int cHi, const int v[8]) {
    if (c0 + 8 <= cHi) {
        *reinterpret_cast<int4 *>(p + c0) = make_int4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<int4 *>(p + c0 + 4) = make_int4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) if (c0 + k < cHi) p[c0 + k] = v[k];
    }
}

__global__ void __launch_bounds__(256)
k_scan_tiles(int cA, int cHi, const int *__restrict__ stay, const int *__restrict__ arrive, const int *__restrict__ birthCount,
             int2 *__restrict__ tileSums) {
    __shared__ int sa[8], sb[8];
    const int c0 = cA + blockIdx.x * SCAN_TILE + threadIdx.x * 8;
    int s8[8], a8[8], b8[8];
    load8(stay, c0, cHi, s8); load8(arrive, c0, cHi, a8); load8(birthCount, c0, cHi, b8);
    int sumA = 0, sumB = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { sumA += s8[k] + a8[k] + b8[k]; sumB += b8[k]; }
    sumA = warp_sum(sumA); sumB = warp_sum(sumB);
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = sumA; sb[threadIdx.x >> 5] = sumB; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int A = 0, Bq = 0;
        for (int w = 0; w < 8; w++) { A += sa[w]; Bq += sb[w]; }
        tileSums[blockIdx.x] = make_int2(A, Bq);
    }
}

__global__ void __launch_bounds__(256)
k_scan_apply(int cA, int cHi, int nTiles, const int *__restrict__ stay, const int *__restrict__ arrive, const int *__restrict__ birthCount,
             const int2 *__restrict__ tileSums, int *__restrict__ newStart, int *__restrict__ birthBase,
             int *__restrict__ count, DevStats *__restrict__ st, int capacity) {
    __shared__ int sa[8], sb[8];
    __shared__ int baseA, baseB;
    if (st->halt) return;  // an earlier queued step failed: the other buffer's cell starts are the valid ones, keep them
    // sum of the tiles before this one
    int pa = 0, pb = 0;
    for (int k = threadIdx.x; k < (int)blockIdx.x; k += 256) { int2 v = tileSums[k]; pa += v.x; pb += v.y; }
    pa = warp_sum(pa); pb = warp_sum(pb);
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = pa; sb[threadIdx.x >> 5] = pb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int A = 0, Bq = 0;
        for (int w = 0; w < 8; w++) { A += sa[w]; Bq += sb[w]; }
        baseA = A; baseB = Bq;
    }
    __syncthreads();
    // each thread owns 8 consecutive cells
    const int c0 = cA + blockIdx.x * SCAN_TILE + threadIdx.x * 8;
    int s8[8], a8[8], b8[8], va[8], vb[8], vc[8];
    load8(stay, c0, cHi, s8); load8(arrive, c0, cHi, a8); load8(birthCount, c0, cHi, b8);
    int ta = 0, tb = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        vc[k] = s8[k] + a8[k] + b8[k];
        va[k] = ta; vb[k] = tb;
        ta += vc[k]; tb += b8[k];
    }
    // block exclusive scan of (ta, tb)
    int ia = ta, ib = tb;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int xa = __shfl_up_sync(0xffffffffu, ia, o), xb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += xa; ib += xb; }
    }
    __syncthreads();
    if (lane == 31) { sa[wid] = ia; sb[wid] = ib; }
    __syncthreads();
    int wa = 0, wb = 0;
    for (int w = 0; w < wid; w++) { wa += sa[w]; wb += sb[w]; }
    const int exA = baseA + wa + ia - ta, exB = baseB + wb + ib - tb;
#pragma unroll
    for (int k = 0; k < 8; k++) { va[k] += exA; vb[k] += exB; }
    store8(newStart, c0, cHi, va);
    store8(birthBase, c0, cHi, vb);
    store8(count, c0, cHi, vc);
    if (blockIdx.x == nTiles - 1 && threadIdx.x == 255) {
        const int total = exA + ta;
        newStart[cHi] = total;
        st->nNew = total;
        if (total > capacity) st->overflow = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// counting-sort scatter of the survivors + creation of the newborns
// (performMoves core/SPopulation.cpp:1058-1092; makeOffspring / createAgentAtIndex :823-847,880-918;
//  makePopSpecificOffspring populations/tut_EnvironAltPop.cpp:141-149)
__global__ void __launch_bounds__(256)
k_scatter(const DevStats *__restrict__ st, AgentArrays a, AgentArrays o, const int *__restrict__ cellStart,
          const int *__restrict__ dest, const int *__restrict__ rank, const uint8_t *__restrict__ oflags,
          const int *__restrict__ newStart, const int *__restrict__ stay, const int *__restrict__ arrive,
          const int *__restrict__ birthBase, float t, int storeAge, int femaleOnly, RngKey key,
          const int *__restrict__ mate, BirthEntry *__restrict__ births, GenomeCtl *__restrict__ gctl) {
    if (st->overflow) return;
    const int n = st->nAgents;
    const unsigned step = st->step;
    const long long nextID = st->nextID;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int d = dest[i];
        const uint8_t f = oflags[i];
        int64_t id = 0;
        if (d >= 0 || (f & F_BORN)) id = a.id[i];
        if (d >= 0) {
            int pos = newStart[d] + stay[d] + rank[i];
            o.id[pos] = id;
            o.birth[pos] = a.birth[i];
            o.lastBirth[pos] = a.lastBirth[i];
            o.cell[pos] = d;
            o.flags[pos] = (uint8_t)(f & (F_MALE | F_FERTILE));
            if (storeAge) o.age[pos] = a.age[i];
            if (a.gslot) o.gslot[pos] = a.gslot[i];
            if (a.nbabies) o.nbabies[pos] = a.nbabies[i] + ((f & F_BORN) ? 1 : 0);  // populations/OoANavGenPop.cpp:243
        }
        if (f & F_BORN) {
            // newborn id = nextID + rank of (cell, mother id) among this step's births
            const int c = a.cell[i];
            const int s = cellStart[c], e = cellStart[c + 1];
            int r = 0;
            for (int j = s; j < e; j++) {
                if ((oflags[j] & F_BORN) && a.id[j] < id) r++;
            }
            const int64_t cid = nextID + birthBase[c] + r;
            const uint32_t g = agent_draws(cid, step, STREAM_BABY, key).x >> 31;  // (uchar)(2*wrandd())
            const int pos = newStart[c] + stay[c] + arrive[c] + r;
            o.id[pos] = cid;
            o.birth[pos] = t;
            o.lastBirth[pos] = 0.0f;
            o.cell[pos] = c;
            // females are born FERTILE, core/SPopulation.cpp:895-898; tut_ParthenoPop turns the drawn males into (non-fertile) females
            o.flags[pos] = (uint8_t)(g ? (femaleOnly ? 0 : F_MALE) : F_FERTILE);
            if (storeAge) o.age[pos] = 0.0f;
            if (o.nbabies) o.nbabies[pos] = 0;
            if (births) record_birth(births, gctl, pos, i, mate[i], cid);  // the genome is made by k_make_offspring
        }
    }
}

// the step's book-keeping on the device (one thread): the new buffer becomes the current one, ids and the step counter advance
__device__ __forceinline__ void step_end_body(DevStats *st, int advanceStep, long long globalBirths) {
    if (st->overflow || st->oversize || st->halt) { st->halt = 1; return; }
    if (advanceStep) { st->agentSteps += st->nAgents; st->totSent += st->nSent; st->totRecv += st->nRecv; }
    st->nAgents = st->nNew;
    // globalBirths: -1 single GPU, -2 the sum the ranks exchanged on the device (st->globalBirths), else the sum from the host
    st->nextID += (globalBirths >= 0) ? globalBirths : (globalBirths == -2 ? st->globalBirths : (long long)st->nBirths);
    if (advanceStep) st->step++;
}
__global__ void k_step_end(DevStats *st, int advanceStep, long long globalBirths) { step_end_body(st, advanceStep, globalBirths); }

// PopBase::getNumAgentsArray hands out ulong counts (core/SPopulation.h); a shard reports 0 for the cells of other ranks
__global__ void k_counts_u64(int cBegin, int cEnd, int cLo, int cHi, const int *__restrict__ count, unsigned long long *__restrict__ out) {
    for (int c = cBegin + blockIdx.x * blockDim.x + threadIdx.x; c < cEnd; c += gridDim.x * blockDim.x)
        out[c] = (c >= cLo && c < cHi) ? (unsigned long long)count[c] : 0ull;
}

// OccTracker::calcBitMap (core/OccTracker.cpp:95-106) asks every population "is anybody in this cell?" for a short list of
// tracked cells after every step: one byte per tracked cell instead of the whole per-cell count array
__global__ void k_occupied(int n, const int *__restrict__ cells, int cLo, int cHi, const int *__restrict__ count, uint8_t *__restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = cells[i];
        out[i] = (c >= cLo && c < cHi && count[c] > 0) ? 1 : 0;
    }
}

// the host has seen the failed step and is about to redo it (or to retry on the generic path)
__global__ void k_clear_halt(DevStats *st) {
    st->halt = 0;
    st->overflow = 0;
    st->oversize = 0;
    st->commError = 0;
}

__global__ void k_fill_age(const DevStats *__restrict__ st, const float *__restrict__ birth, float *__restrict__ age, float t) {
    const int n = st->nAgents;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) age[i] = __fsub_rn(t, birth[i]);
}

__global__ void k_atan_prob(int n, const float *__restrict__ age, double *__restrict__ out, double maxAge, double slope, double scale) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double x = __dmul_rn(slope, __dadd_rn((double)age[i], -maxAge));
        out[i] = __dadd_rn(0.5, __ddiv_rn(__dmul_rn(scale, atan_rn(x)), 3.141592653589793));
    }
}

__global__ void k_gather_mate_id(const DevStats *__restrict__ st, const int64_t *__restrict__ id, const int *__restrict__ mate,
                                 int64_t *__restrict__ out) {
    const int n = st->nAgents;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int m = mate[i];
        out[i] = (m >= 0) ? id[m] : (int64_t)m;
    }
}

}  // namespace qhg
