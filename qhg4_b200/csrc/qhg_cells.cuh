// qhg_cells.cuh -- the fast path: two warp-per-cell passes over the agent state per step.
//
// Agents are binned by cell, so one warp can own one cell: it walks the cell's contiguous segment in chunks of
// 32 agents, keeps the per-cell quantities (fertile counts, pairing keys, provisional decisions) in its own slice
// of shared memory and never needs a block-wide barrier.
//
//   pass 1  k_cell_decide   reads id 8 + birth 4 + lastBirth 4 + flags 1 B per agent; pairing inside the cell,
//                           all actions, ONE decision byte per agent written back (1 B), per-cell stay / birth
//                           counts stored, movers counted into arrive[dest] with one atomic per direction
//   scan    k_scan_*        new cell starts from the counts
//   pass 2  k_cell_scatter  reads the decision byte + id, birth, lastBirth (17 B), writes the survivors, movers
//                           and newborns at their new position (id, birth, lastBirth, cell, flags: 21 B)
//
// Rare, expensive work is not done where it is found (a warp would wait for the few lanes that need it) but queued
// in shared memory and worked off 32 at a time with all lanes busy: the double-precision atan of ATanDeath (only
// ages inside the window where the death probability is strictly between 0 and 1) and the neighbour choice of
// WeightedMove (only the agents whose first draw said "move").  Actions after a queued ATanDeath are evaluated
// speculatively and voided at commit time if the agent turns out dead, which is equivalent because an action has
// no side effect before the commit and draws are keyed by (agent, step, stream), not consumed from a sequence.
//
// Nothing depends on the order of agents inside a cell: pairing ranks by (random key, id), newborn ids by
// (cell, mother id).  So positions may be handed out by atomics in any order and the result is still identical
// to the oracle as a set of agents.
#pragma once
#include <type_traits>
#include "qhg_kernels.cuh"

namespace qhg {

constexpr int CW = 4;            // warps per CTA (scatter pass)
#ifndef QHG_DCW
#define QHG_DCW 4
#endif
constexpr int DCW = QHG_DCW;     // warps per CTA in the decide pass (1: the per-warp shared-memory slice has a compile-time address)
#ifndef QHG_BATCH_FETCH
#define QHG_BATCH_FETCH 0
#endif
#ifndef QHG_PRETHRESH
#define QHG_PRETHRESH 1   // LinearBirth / LinearDeath thresholds come precomputed per cell from k_cell_init
#endif
#ifndef QHG_OPAQUE_SMEM
#define QHG_OPAQUE_SMEM 1        // the address of a warp's shared-memory slice is kept in a register instead of being rebuilt at every use (pass 1: -3 %)
#endif
constexpr int WCAP = 1024;       // largest cell (agents) the fast path handles; larger ones -> generic path
#ifndef QHG_MAXF
#define QHG_MAXF 512
#endif
#ifndef QHG_QCAP
#define QHG_QCAP 64
#endif
constexpr int MAXF = QHG_MAXF;   // most fertile females of one cell that can be ranked in shared memory
constexpr int QCAP = QHG_QCAP;   // work-queue entries per warp (flushed when the next round could overflow them)
constexpr int MVCAP = 96;        // movers queued per warp in the scatter pass (flushed once per window, or when the next round could overflow); 64 for the smaller windows
constexpr int MOVE_STRIDE = 8;    // ints per cell in moveBase[] (one 32-byte sector)
constexpr int AGENT_SLACK = 64;  // elements allocated past the capacity of every per-agent array (aligned bulk reads)
constexpr int MAXMOTHERS = 128;  // most births of one cell per step on the fast path
constexpr int BIRTH_FLUSH = 24;  // pass 2 collects the mothers of several cells and places their babies together once it has this many
#ifndef QHG_CELL_BATCH
#define QHG_CELL_BATCH 4
#endif
constexpr int CELL_BATCH = QHG_CELL_BATCH;    // consecutive cells a warp takes per grab of the work counter
#ifndef QHG_DU
#define QHG_DU 1
#endif
constexpr int DU = QHG_DU;       // agents per lane and chunk in the decide pass

// decision byte handed from pass 1 to pass 2: bit0 male, bit1 fertile (the agent's new flags), bit2 gave birth,
// bits 3-5 move code: 0 stays, 1..6 neighbour slot + 1, 7 dead
constexpr int DEC_MOVE_SHIFT = 3;
constexpr uint8_t DEC_DEAD = 7;
// transient bits while a cell is being worked on (bit2 = birth candidate until the pairing is settled)
constexpr uint8_t T_ATANDIES = 0x40, T_DEADNOW = 0x80;

struct WarpSmem {
    double row[8];             // the cell's cumulated weight row (7 used)
    long long qmId[QCAP];      // WeightedMove queue: agent id
    float qaAge[QCAP];         // ATanDeath queue: the agent's age
    uint32_t qaU[QCAP];        //                  the agent's death draw
    alignas(16) uint32_t keys[MAXF];  // pairing keys of the cell's fertile females
    uint16_t ffJ[MAXF];        //   and their position in the cell
    uint16_t qaJ[QCAP];
    uint16_t qmJ[QCAP];        // WeightedMove queue: position in the cell
    uint16_t candQ[MAXF];      // birth candidates among the fertile females (index into keys / ffJ)
    int outC[8];               // movers of the cell per direction
    int nbrC[8];               // the cell's neighbours (6 used)
    alignas(4) uint8_t dec[WCAP + 4];  // provisional decision of every agent of the cell, shifted by (cell start & 3)
};

// the same for populations with Genetics (k_cell_decide<false, true>): births need the identity of the father, so the fertile
// males are listed and keyed like the females; fewer ranked agents per cell fit (larger cells take the generic path)
constexpr int MAXF_G = 192;
struct WarpSmemG {
    double row[8];
    long long qmId[QCAP];
    float qaAge[QCAP];
    uint32_t qaU[QCAP];
    alignas(16) uint32_t keys[MAXF_G];   // pairing keys of the fertile females ...
    alignas(16) uint32_t mkeys[MAXF_G];  // ... and of the fertile males
    uint16_t ffJ[MAXF_G];
    uint16_t mmJ[MAXF_G];                // positions of the fertile males in the cell
    uint16_t qaJ[QCAP];
    uint16_t qmJ[QCAP];
    uint16_t candQ[MAXF_G];
    uint16_t candR[MAXF_G];              // rank of every birth candidate among the fertile females
    uint16_t maleOfRank[MAXF_G];         // position of the fertile male of rank r
    int outC[8];
    int nbrC[8];
    alignas(4) uint8_t dec[WCAP + 4];
};

// NAV = true (k_cell_decide<false, GEN, true>): Navigate is the last action of the program.  Far jumps have no direction code:
// the agent is marked as leaving (move code 7, like a dead one, so the scatter pass skips it) and goes into a jump list with a
// slot among the arrivals of its destination; k_place_jumpers writes it there after the scatter pass.
struct JumpEntry {
    int src;    // position in the current buffer
    int to;     // destination cell
    int slot;   // slot among the arrivals of that cell
    int bits;   // bit0 male, bit1 fertile, bit2 gave birth this step
};
// does Navigate see the MOVING bit of an earlier move?  (Fertility overwrites the whole life state, actions/Fertility.cpp:56-68)
__device__ __forceinline__ bool nav_sees_moving(unsigned long long prog, int nOps) {
    bool moving = false;
    for (int k = 0; k < nOps; k++) {
        const int op = (int)((prog >> (4 * k)) & 15ull);
        if (op == OP_WEIGHTEDMOVE || op == OP_RANDOMMOVE || op == OP_CONDWEIGHTEDMOVE) moving = true;
        if (op == OP_FERTILITY) moving = false;
        if (op == OP_NAVIGATE) return moving;
    }
    return false;
}

// the action program the fast path is specialised for at compile time: the tutorial populations' order
// GetOld, ATanDeath, WeightedMove, Fertility, Verhulst (tutorial_data/xmldat/tut_EnvironAlt.xml priorities)
constexpr unsigned long long PROG_TUT5 = (unsigned long long)OP_GETOLD | ((unsigned long long)OP_ATANDEATH << 4) |
                                         ((unsigned long long)OP_WEIGHTEDMOVE << 8) | ((unsigned long long)OP_FERTILITY << 12) |
                                         ((unsigned long long)OP_VERHULST << 16);

struct ProgramInfo {  // warp-uniform facts about the action program
    bool needAct0, hasFert, hasVerhulst;
    bool moveAfterAtan, bornAfterAtan;
    bool randomMove;  // the move action is RandomMove (uniform direction, no ice test) instead of WeightedMove
    bool condMove;    // ... or CondWeightedMove (k_seg_decide only)
};

__device__ __forceinline__ ProgramInfo program_info(unsigned long long prog, int nOps) {
    ProgramInfo I{false, false, false, false, false, false, false};
    int ka = -1;
    for (int k = 0; k < nOps; k++) {
        int op = (int)((prog >> (4 * k)) & 15ull);
        if (op == OP_ATANDEATH) { ka = k; I.needAct0 = true; }
        if (op == OP_WEIGHTEDMOVE || op == OP_RANDOMMOVE || op == OP_CONDWEIGHTEDMOVE) {
            I.needAct0 = true; I.randomMove = (op == OP_RANDOMMOVE); I.condMove = (op == OP_CONDWEIGHTEDMOVE);
            if (ka >= 0) I.moveAfterAtan = true;
        }
        if (op == OP_VERHULST) { I.needAct0 = true; I.hasVerhulst = true; if (ka >= 0) I.bornAfterAtan = true; }
        if (op == OP_FERTILITY) I.hasFert = true;
    }
    return I;
}

// ---------------------------------------------------------------------------------------------
// pass 1.  SPEC = true: the program is PROG_TUT5, known at compile time (straight-line code);
//          SPEC = false: any program, interpreted from P.prog.
#ifndef QHG_PF2
#define QHG_PF2 0
#endif
#ifndef QHG_DECIDE_MINB
#define QHG_DECIDE_MINB (32 / QHG_DCW)
#endif
constexpr int DECIDE_CTAS_PER_SM = QHG_DECIDE_MINB;
// GEN = true (never together with SPEC): the population has Genetics; `father[i]` receives, for every mother-to-be at position
// i of the current buffer, the position of her mate (the scatter pass turns the two into a birth record)
template <bool SPEC, bool GEN = false, bool NAV = false>
__global__ void __launch_bounds__(DCW * 32, QHG_DECIDE_MINB)
k_cell_decide(DevStats *__restrict__ st, AgentArrays a, ActParams P, CellEnv E, int cLo, int cHi, const int *__restrict__ cellStart,
              int doPair, int *__restrict__ stay, int *__restrict__ arrive, int *__restrict__ birthCount, uint8_t *__restrict__ dec,
              int *__restrict__ moveBase, int *__restrict__ father = nullptr, JumpEntry *__restrict__ jumps = nullptr,
              int *__restrict__ jumpCount = nullptr, int jumpCap = 0) {
    static_assert(!(SPEC && (GEN || NAV)), "the compile-time program has neither Genetics nor Navigate");
    using WS = typename std::conditional<GEN, WarpSmemG, WarpSmem>::type;
    constexpr int FCAP = GEN ? MAXF_G : MAXF;
    __shared__ WS smem[DCW];
    const int lane = threadIdx.x & 31, wid = (DCW == 1) ? 0 : (int)(threadIdx.x >> 5);
    WS &S = smem[wid];
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = lanemask_lt();
    if (st->halt) return;  // an earlier queued step failed (qhgb_run): nothing happens until the host has dealt with it
    const unsigned step = st->step;
    const unsigned long long prog = SPEC ? PROG_TUT5 : P.prog;
    const int nOps = SPEC ? 5 : P.nOps;
    const ProgramInfo I = program_info(prog, nOps);
    int nDead = 0, nMove = 0, nBorn = 0;  // warp-uniform tallies
    int pendCell = -1, pendBase = 0;      // slot reservation of the previous cell, stored one cell later (atomic latency)
    // loop invariants in registers
    const float tNow = P.t, fertMin = P.fertMinAge, fertMax = P.fertMaxAge, fertInter = P.fertInterbirth;
    const float atanAgeLo = P.atanAgeLo, atanAgeHi = P.atanAgeHi;
    const unsigned long long tMove = prob_threshold(P.moveProb);
    const RngKey key = P.key;
    const RoundKeys &RK = P.rk;
    const bool storeAge = P.storeAge != 0;
    const bool selfMate = !SPEC && P.selfMate != 0;                    // tut_ParthenoPop: every female counts as mated
    const bool confine = !SPEC && P.confine != 0 && E.allowed != nullptr;  // ConfinedMove filters the chosen destinations

    // cells are handed out dynamically in batches of CELL_BATCH consecutive cells (sea cells are empty, land cells are
    // not: a static split leaves a long tail)
    for (;;) {
    int cBase = 0;
    if (lane == 0) cBase = cLo + atomicAdd(&st->workDecide, CELL_BATCH);
    cBase = __shfl_sync(FULL, cBase, 0);
    if (cBase >= cHi) break;
    // everything per cell of the batch is fetched at once, one lane per item (lane = 8 * cell + k for the rows): one
    // memory round trip per batch instead of two dependent ones per cell
    static_assert(CELL_BATCH * 8 <= 32 && WSTRIDE <= 8 && MAXN <= 8, "one lane per (cell of the batch, row entry)");
    const int nB = min(CELL_BATCH, cHi - cBase);
    const int bc = lane >> 3, bk = lane & 7;
    int csL = 0, nnL = 0, nbL = -1;
    double wL = 0.0, bL = 0.0, dL = 0.0;
#if QHG_BATCH_FETCH
    if (lane <= nB) csL = cellStart[cBase + lane];
    if (lane < nB) {
        nnL = E.nNbr[cBase + lane];
        if (I.hasVerhulst) { bL = E.B[cBase + lane]; dL = E.D[cBase + lane]; }
    }
    if (bc < nB) {
        if (bk < WSTRIDE) wL = E.W[(size_t)(cBase + bc) * WSTRIDE + bk];
        if (bk < MAXN) nbL = E.nbr[(size_t)(cBase + bc) * MAXN + bk];
    }
#endif
    for (int ci = 0; ci < nB; ci++) {
        const int c = cBase + ci;
#if QHG_BATCH_FETCH
        const int s = __shfl_sync(FULL, csL, ci), n = __shfl_sync(FULL, csL, ci + 1) - s;
#else
        const int s = cellStart[c], n = cellStart[c + 1] - s;
#endif
        if (n == 0) continue;
        if (n > WCAP) {
            if (lane == 0) atomicExch(&st->oversize, 1);
            continue;
        }
        // the cell's bytes sit at dec[gOff + j]: shared-memory word k then is the aligned global word of dec[] it is stored to
        const int gOff = s & 3;
        uint8_t *const sdec = S.dec + gOff;
#if QHG_BATCH_FETCH
        {
            const double rv = __shfl_sync(FULL, wL, ci * 8 + bk);
            const int nv = __shfl_sync(FULL, nbL, ci * 8 + bk);
            if (lane < 8) {
                S.outC[lane] = 0;
                S.row[lane] = rv;
                S.nbrC[lane] = nv;
            }
        }
#else
        if (lane < 8) {
            S.outC[lane] = 0;
            S.row[lane] = (lane < WSTRIDE) ? E.W[(size_t)c * WSTRIDE + lane] : 0.0;
            S.nbrC[lane] = (lane < MAXN) ? E.nbr[(size_t)c * MAXN + lane] : -1;
        }
#endif
        int nF = 0, nM = 0, nqa = 0, nqm = 0;
        int confL = 0;  // moves of this lane that ConfinedMove turned back (they stay in the move list: counted, core/SPopulation.cpp:1067)
        bool tooMany = false;
#if QHG_BATCH_FETCH
        const int nreal = __shfl_sync(FULL, nnL, ci);
        const double bC = __shfl_sync(FULL, bL, ci), dC = __shfl_sync(FULL, dL, ci);
#else
        const int nreal = E.nNbr[c];
#if !QHG_PRETHRESH
        const double bC = I.hasVerhulst ? E.B[c] : 0.0, dC = I.hasVerhulst ? E.D[c] : 0.0;
#endif
#endif
        const double *row = S.row;
        // the probability tests of LinearBirth / LinearDeath as exact integer thresholds on the 32-bit draws
#if QHG_PRETHRESH && !QHG_BATCH_FETCH
        const unsigned long long tbw = I.hasVerhulst ? E.TB[c] : 0ull;
        const unsigned long long tDeath = I.hasVerhulst ? E.TD[c] : 0ull;
        const bool bPos = (tbw >> 62) & 1ull, bNeg = (tbw >> 63) != 0;
        const unsigned long long tBirth = tbw & 0x1ffffffffull, tBirthNeg = tBirth;
#else
        const unsigned long long tBirth = prob_threshold(bC), tBirthNeg = prob_threshold(-bC), tDeath = prob_threshold(dC);
        const bool bPos = bC > 0, bNeg = bC < 0;
#endif
        auto flush_atan = [&]() {  // ATanDeath::execute, actions/ATanDeath.cpp:75-83, for the queued agents
            for (int e = lane; e < nqa; e += 32) {
                const double x = __dmul_rn(P.atanSlope, __dadd_rn((double)S.qaAge[e], -P.atanMaxAge));
                const double p = __dadd_rn(0.5, __ddiv_rn(__dmul_rn(P.atanScale, atan_rn(x)), 3.141592653589793));
                if (u2d(S.qaU[e]) < p) sdec[S.qaJ[e]] |= T_ATANDIES;
            }
            nqa = 0;
            __syncwarp();
        };
        auto flush_move = [&]() {  // WeightedMove::execute, actions/WeightedMove.cpp:56-98, for the queued agents
            for (int e = lane; e < nqm; e += 32) {
                const int j = S.qmJ[e];
                const uint32_t u = agent_draws_rk(S.qmId[e], step, STREAM_ACT1, RK).x;
                int pick = -1;
                if (!SPEC && I.randomMove) {  // RandomMove: uniform over "stay" and the neighbours, no ice test
                    pick = (int)__dmul_rn(u2d(u), (double)(nreal + 1));
                    if (pick > 0 && S.nbrC[pick - 1] >= 0) {
                        if (confine && !E.allowed[S.nbrC[pick - 1]]) { if (!(I.moveAfterAtan && (sdec[j] & T_ATANDIES))) confL++; }
                        else sdec[j] |= (uint8_t)(pick << DEC_MOVE_SHIFT);
                    }
                    continue;
                }
                const double wmax = row[nreal];
                if (row[0] == wmax) {
                    pick = (int)u2int(u, 0, nreal + 1);
                } else {
                    const double r2 = __dmul_rn(u2d(u), wmax);
                    for (int q = 0; q < nreal + 1; q++) {
                        if (r2 < row[q]) { pick = q; break; }
                    }
                }
                if (pick > 0) {
                    const int dst = S.nbrC[pick - 1];
                    if (dst >= 0 && !(E.ice && E.ice[dst])) {
                        // ConfinedMove (actions/ConfinedMove.cpp:86-101): the move is registered and counted, but leads back to
                        // the cell it starts from -- unless ATanDeath (flushed before, see below) removed the agent before it moved
                        if (confine && !E.allowed[dst]) { if (!(I.moveAfterAtan && (sdec[j] & T_ATANDIES))) confL++; }
                        else sdec[j] |= (uint8_t)(pick << DEC_MOVE_SHIFT);
                    }
                }
            }
            nqm = 0;
            __syncwarp();
        };

        // ---- one pass over the cell: fertile census + all actions, provisional decisions --------------------------
        // software pipeline: the next chunk's loads are in flight while this chunk is evaluated; every lane carries
        // DU agents per chunk so that their (independent) Philox chains overlap
        int64_t idN[DU]; float birthN[DU], lastN[DU], ageN[DU]; uint8_t fN[DU];
#pragma unroll
        for (int u = 0; u < DU; u++) {
            const int j = u * 32 + lane;
            idN[u] = 0; birthN[u] = 0; lastN[u] = 0; ageN[u] = 0; fN[u] = 0;
            if (j < n) {
                idN[u] = a.id[s + j]; birthN[u] = a.birth[s + j]; fN[u] = a.flags[s + j];
                if (I.hasFert) lastN[u] = a.lastBirth[s + j];
                if (storeAge) ageN[u] = a.age[s + j];
            }
        }
#if QHG_PF2  // second prefetch stage: two chunks ahead
        int64_t idM[DU]; float birthM[DU], lastM[DU], ageM[DU]; uint8_t fM[DU];
#pragma unroll
        for (int u = 0; u < DU; u++) {
            const int j = 32 * DU + u * 32 + lane;
            idM[u] = 0; birthM[u] = 0; lastM[u] = 0; ageM[u] = 0; fM[u] = 0;
            if (j < n) {
                idM[u] = a.id[s + j]; birthM[u] = a.birth[s + j]; fM[u] = a.flags[s + j];
                if (I.hasFert) lastM[u] = a.lastBirth[s + j];
                if (storeAge) ageM[u] = a.age[s + j];
            }
        }
#endif
        for (int j0 = 0; j0 < n; j0 += 32 * DU) {
            int64_t id[DU]; float birth[DU], lastBirth[DU], age[DU]; uint8_t f0[DU];
            uint4 r0[DU];
#pragma unroll
            for (int u = 0; u < DU; u++) {
                id[u] = idN[u]; birth[u] = birthN[u]; lastBirth[u] = lastN[u]; age[u] = ageN[u]; f0[u] = fN[u];
#if QHG_PF2
                idN[u] = idM[u]; birthN[u] = birthM[u]; lastN[u] = lastM[u]; ageN[u] = ageM[u]; fN[u] = fM[u];
                const int j2 = j0 + 2 * 32 * DU + u * 32 + lane;
                if (j2 < n) {
                    idM[u] = a.id[s + j2]; birthM[u] = a.birth[s + j2]; fM[u] = a.flags[s + j2];
                    if (I.hasFert) lastM[u] = a.lastBirth[s + j2];
                    if (storeAge) ageM[u] = a.age[s + j2];
                }
#else
                const int j2 = j0 + 32 * DU + u * 32 + lane;
                if (j2 < n) {
                    idN[u] = a.id[s + j2]; birthN[u] = a.birth[s + j2]; fN[u] = a.flags[s + j2];
                    if (I.hasFert) lastN[u] = a.lastBirth[s + j2];
                    if (storeAge) ageN[u] = a.age[s + j2];
                }
#endif
            }
#pragma unroll
            for (int u = 0; u < DU; u++) r0[u] = I.needAct0 ? agent_draws_rk(id[u], step, STREAM_ACT0, RK) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < DU; u++) {
                if (u > 0 && j0 + u * 32 >= n) break;  // warp-uniform: the cell's tail leaves this sub-chunk empty
                const int j = j0 + u * 32 + lane;
                const bool valid = j < n;
                // fertile census (for the pairing): positions of the fertile females, number of fertile males
                const bool fertF = valid && ((f0[u] & (F_FERTILE | F_MALE)) == F_FERTILE);
                const unsigned mF = __ballot_sync(FULL, fertF);
                if constexpr (GEN) {  // the fertile males are listed too
                    const bool fertM = valid && ((f0[u] & (F_FERTILE | F_MALE)) == (F_FERTILE | F_MALE));
                    const unsigned mMm = __ballot_sync(FULL, fertM);
                    if (nM + __popc(mMm) > FCAP) tooMany = true;
                    else if (fertM) S.mmJ[nM + __popc(mMm & lt)] = (uint16_t)j;
                    nM += __popc(mMm);
                } else {
                    nM += __popc(__ballot_sync(FULL, valid && ((f0[u] & (F_FERTILE | F_MALE)) == (F_FERTILE | F_MALE))));
                }
                if (nF + __popc(mF) > FCAP) tooMany = true;
                else if (fertF) S.ffJ[nF + __popc(mF & lt)] = (uint16_t)j;
                nF += __popc(mF);

                bool needAtan = false, needMove = false;
                float ag = age[u];
                if (valid) {
                    uint8_t f = f0[u] & (F_MALE | F_FERTILE);
                    bool alive = true, cand = false;
#pragma unroll
                    for (int k = 0; k < (SPEC ? 5 : MAX_OPS); k++) {
                        if (!SPEC && k >= nOps) break;
                        if (!alive) break;
                        const int op = (int)((prog >> (4 * k)) & 15ull);
                        if (op == OP_GETOLD) {  // actions/GetOld.cpp:37-48
                            ag = __fsub_rn(tNow, birth[u]);
                        } else if (op == OP_ATANDEATH) {  // actions/ATanDeath.cpp:66-90
                            ag = __fsub_rn(tNow, birth[u]);
                            if (ag >= atanAgeLo) {        // below the window the probability is < 0: nobody dies
                                if (ag <= atanAgeHi) needAtan = true;  // inside: decided exactly when the queue is flushed
                                else alive = false;       // above the window the probability is > 1
                            }
                        } else if (op == OP_OLDAGEDEATH) {  // actions/OldAgeDeath.cpp:48-67
                            ag = __fsub_rn(tNow, birth[u]);
                            const uint32_t uo = agent_draws(id[u], step, STREAM_ACT1, key).w;
                            if ((double)ag > __dadd_rn(P.oadMaxAge, u2range(uo, P.oadLo, P.oadHi))) alive = false;
                        } else if (op == OP_WEIGHTEDMOVE || op == OP_RANDOMMOVE) {  // actions/WeightedMove.cpp:45-106, RandomMove.cpp:65-100; the neighbour is chosen at the flush
                            if ((unsigned long long)r0[u].y < tMove) needMove = true;
                        } else if (op == OP_FERTILITY) {  // actions/Fertility.cpp:49-74
                            bool fert;
                            if (!(f & F_MALE)) fert = (ag > fertMin) && (ag < fertMax) && (__fsub_rn(tNow, lastBirth[u]) > fertInter);
                            else fert = ag > fertMin;
                            f = (uint8_t)((f & F_MALE) | (fert ? F_FERTILE : 0));
                        } else if (op == OP_VERHULST) {  // actions/Verhulst.cpp:101-115 -> LinearBirth.cpp:122-168, LinearDeath.cpp:131-153
                            if (bPos) {
                                // a birth needs a mate (LinearBirth.cpp:142); whether this fertile female got one is settled
                                // once the whole cell has been seen: she is a candidate until then
                                const bool mayBear = selfMate ? !(f0[u] & F_MALE) : ((f0[u] & (F_FERTILE | F_MALE)) == F_FERTILE);
                                if (mayBear && (unsigned long long)r0[u].z < tBirth) cand = true;
                            } else if (bNeg) {
                                if ((unsigned long long)r0[u].z < tBirthNeg) alive = false;
                            }
                            if (alive && (unsigned long long)r0[u].w < tDeath) alive = false;
                        } else if (op == OP_DROWN) {  // populations/tut_EnvironAltPop.cpp:100-116 (EVENT_ID_GEO)
                            if (E.alt[c] < 0 || (E.ice && E.ice[c])) alive = false;
                        }
                    }
                    if (storeAge) a.age[s + j] = ag;
                    sdec[j] = (uint8_t)(f | (cand ? F_BORN : 0) | (alive ? 0 : T_DEADNOW));
                }
                // queue the rare expensive work
                const unsigned ma = __ballot_sync(FULL, needAtan), mm = __ballot_sync(FULL, needMove);
                if (needAtan) { const int e = nqa + __popc(ma & lt); S.qaAge[e] = ag; S.qaU[e] = r0[u].x; S.qaJ[e] = (uint16_t)j; }
                if (needMove) { const int e = nqm + __popc(mm & lt); S.qmJ[e] = (uint16_t)j; S.qmId[e] = id[u]; }
                nqa += __popc(ma);
                nqm += __popc(mm);
            }
            __syncwarp();
            // one call site each (code size): flush when a queue could overflow in the next round, and at the end of the cell
            const bool last = j0 + 32 * DU >= n;
            // with ConfinedMove the move flush reads the ATanDeath verdicts of its agents: the death queue goes first
            if (nqa > QCAP - 32 * DU || (last && nqa > 0) || (confine && nqa > 0 && (nqm > QCAP - 32 * DU || (last && nqm > 0)))) flush_atan();
            if (nqm > QCAP - 32 * DU || (last && nqm > 0)) flush_move();
        }

        // ---- Navigate (actions/Navigate.cpp:181-250), the last action of the program: the agents of port and bridge cells ------
        if constexpr (NAV) {
            const int port = E.navRow ? E.navRow[c] : -1;
            bool hasBridge = false;
            for (int b = 0; b < E.nBridges; b++) { const int2 br = E.bridges[b]; hasBridge |= (br.x == c) || (br.y == c); }
            if (port >= 0 || hasBridge) {  // warp-uniform
                const bool seesMoving = nav_sees_moving(prog, nOps);
                int p0 = 0, nd = 0;
                if (port >= 0) { p0 = E.navPtr[port]; nd = E.navPtr[port + 1] - p0 - 1; }
                const int lim = (c < nd) ? c : nd;  // the reference bounds the search by the port's cell index (:194)
                for (int j = lane; j < n; j += 32) {
                    const uint8_t b0 = sdec[j];
                    if (b0 & (T_ATANDIES | T_DEADNOW)) continue;          // dead before the last action
                    const int code0 = (b0 >> DEC_MOVE_SHIFT) & 7;
                    if (seesMoving && code0 != 0) continue;               // LIFE_STATE_MOVING is still set
                    const int64_t idj = a.id[s + j];
                    int to = -1, navMoves = 0;
                    if (port >= 0) {
                        const double r = u2d(agent_draws(idj, step, STREAM_ACT1, key).y);
                        int i = 0;
                        while (i < lim && r > E.navCum[p0 + i]) i++;
                        if (i > 0) {
                            const int dst = E.navDest[p0 + i];
                            if (!(E.ice && E.ice[dst])) { to = dst; navMoves++; }
                        }
                    }
                    for (int b = 0; b < E.nBridges; b++) {  // manual bridges: one draw per incident bridge (:228-247)
                        const int2 br = E.bridges[b];
                        const int dst = (br.x == c) ? br.y : ((br.y == c) ? br.x : -1);
                        if (dst >= 0) {
                            const uint4 db = agent_draws(idj, step, 0x04000000u | (unsigned)(b / 4), key);
                            const unsigned wv = (b & 3) == 0 ? db.x : (b & 3) == 1 ? db.y : (b & 3) == 2 ? db.z : db.w;
                            if (u2d(wv) < E.bridgeProb) { to = dst; navMoves++; }
                        }
                    }
                    if (navMoves > 0) {  // the last registered move decides where the agent ends up; every one of them counts
                        const int slot = atomicAdd(&arrive[to], 1);
                        const int e = atomicAdd(jumpCount, 1);
                        if (e < jumpCap) jumps[e] = JumpEntry{s + j, to, slot, (int)(b0 & 7)};
                        else atomicExch(&st->oversize, 1);  // the list is full: the step is redone on the generic path
                        sdec[j] = (uint8_t)((b0 & 7) | (DEC_DEAD << DEC_MOVE_SHIFT));
                        confL += (code0 != 0 ? 1 : 0) + navMoves - 1;  // the commit counts one move for a leaving agent
                    }
                }
                __syncwarp();
            }
        }

        // ---- pairing: RandomPair::findMates (actions/RandomPair.cpp:146-279) under the counter-mode law ------------
        // fertile females and fertile males are ranked by (random key, id); equal ranks mate.  Only "does this female
        // have a mate" matters to the actions: with nF <= nM every fertile female has one, otherwise the nM females
        // with the smallest keys -- and only the birth candidates need to know.
        bool mates = selfMate || (doPair && nF > 0 && nM > 0);
        if constexpr (GEN) {
            // every birth needs her mate's identity: both sexes are ranked by (key, id), equal ranks mate
            if (mates && !selfMate) {
                if (tooMany) {
                    if (lane == 0) atomicExch(&st->oversize, 1);
                    continue;
                }
                int nCand = 0;
                for (int q0 = 0; q0 < nF; q0 += 32) {
                    const int q = q0 + lane;
                    const bool isCand = (q < nF) && (sdec[S.ffJ[q]] & F_BORN);
                    const unsigned mc = __ballot_sync(FULL, isCand);
                    if (isCand) S.candQ[nCand + __popc(mc & lt)] = (uint16_t)q;
                    nCand += __popc(mc);
                }
                if (nCand > 0) {  // warp-uniform
                    for (int q = lane; q < nF; q += 32) S.keys[q] = agent_draws_rk(a.id[s + S.ffJ[q]], step, STREAM_PAIR, RK).x;
                    for (int m = lane; m < nM; m += 32) S.mkeys[m] = agent_draws_rk(a.id[s + S.mmJ[m]], step, STREAM_PAIR, RK).x;
                    __syncwarp();
                    const int np = min(nF, nM);
                    for (int i = lane; i < nCand; i += 32) {  // rank of every candidate among the fertile females
                        const int q = S.candQ[i];
                        const uint32_t k = S.keys[q];
                        int r = 0;
                        bool tie = false;
                        for (int e = 0; e < nF; e++) {
                            const uint32_t ke = S.keys[e];
                            r += (ke < k) ? 1 : 0;
                            tie |= (ke == k) && (e != q);
                        }
                        if (tie) {  // equal keys: the id decides
                            const int64_t myId = a.id[s + S.ffJ[q]];
                            for (int e = 0; e < nF; e++) if (e != q && S.keys[e] == k && a.id[s + S.ffJ[e]] < myId) r++;
                        }
                        S.candR[i] = (uint16_t)r;
                        if (r >= np) sdec[S.ffJ[q]] &= (uint8_t)~F_BORN;  // no mate: no birth
                    }
                    for (int m = lane; m < nM; m += 32) {  // rank of every fertile male; the first np of them are mates
                        const uint32_t k = S.mkeys[m];
                        int r = 0;
                        bool tie = false;
                        for (int e = 0; e < nM; e++) {
                            const uint32_t ke = S.mkeys[e];
                            r += (ke < k) ? 1 : 0;
                            tie |= (ke == k) && (e != m);
                        }
                        if (tie) {
                            const int64_t myId = a.id[s + S.mmJ[m]];
                            for (int e = 0; e < nM; e++) if (e != m && S.mkeys[e] == k && a.id[s + S.mmJ[e]] < myId) r++;
                        }
                        if (r < np) S.maleOfRank[r] = S.mmJ[m];
                    }
                    __syncwarp();
                    for (int i = lane; i < nCand; i += 32) {
                        const int r = S.candR[i];
                        if (r < np) father[s + S.ffJ[S.candQ[i]]] = s + S.maleOfRank[r];
                    }
                    __syncwarp();
                }
            }
        } else
        if (mates && !selfMate && nF > nM) {
            if (tooMany) {
                if (lane == 0) atomicExch(&st->oversize, 1);
                continue;
            }
            for (int q = lane; q < nF; q += 32) S.keys[q] = agent_draws_rk(a.id[s + S.ffJ[q]], step, STREAM_PAIR, RK).x;
            __syncwarp();
            // the candidates, compacted, so that every lane of the ranking loop has work
            int nCand = 0;
            for (int q0 = 0; q0 < nF; q0 += 32) {
                const int q = q0 + lane;
                const bool isCand = (q < nF) && (sdec[S.ffJ[q]] & F_BORN);
                const unsigned mc = __ballot_sync(FULL, isCand);
                if (isCand) S.candQ[nCand + __popc(mc & lt)] = (uint16_t)q;
                nCand += __popc(mc);
            }
            __syncwarp();
            for (int i = lane; i < nCand; i += 32) {
                const int q = S.candQ[i];
                const uint32_t k = S.keys[q];
                // rank = number of smaller keys; four keys per shared-memory load; "<=" counts reveal ties (the key itself is one)
                int r = 0, le = 0;
                const int nF4 = nF & ~3;
                for (int e = 0; e < nF4; e += 4) {
                    const uint4 kk = *reinterpret_cast<const uint4 *>(&S.keys[e]);
                    r += (kk.x < k) + (kk.y < k) + (kk.z < k) + (kk.w < k);
                    le += (kk.x <= k) + (kk.y <= k) + (kk.z <= k) + (kk.w <= k);
                }
                for (int e = nF4; e < nF; e++) {
                    const uint32_t ke = S.keys[e];
                    r += (ke < k) ? 1 : 0;
                    le += (ke <= k) ? 1 : 0;
                }
                const bool tie = (le - r) > 1;
                if (tie) {  // equal keys (about one pair in 10^8): the id decides
                    const int64_t myId = a.id[s + S.ffJ[q]];
                    for (int e = 0; e < nF; e++) {
                        if (e != q && S.keys[e] == k && a.id[s + S.ffJ[e]] < myId) r++;
                    }
                }
                if (r >= nM) sdec[S.ffJ[q]] &= (uint8_t)~F_BORN;  // no mate: no birth
            }
            __syncwarp();
        }

        // ---- commit: final decision bytes, per-cell counts ---------------------------------------------------------
        // four agents (one 32-bit word of decision bytes) per lane and round, all byte lanes in parallel; tallies in
        // registers, one warp reduction per cell
        int stayL = 0, bornL = 0, moveL = 0, outL = 0;
        {
            const uint32_t ONES = 0x01010101u;
            const uint32_t matesM = mates ? ONES : 0u, bornVoid = I.bornAfterAtan ? ONES : 0u, moveVoid = I.moveAfterAtan ? ONES : 0u;
            const int nWords = (gOff + n + 3) >> 2;
            const uint32_t *sw = reinterpret_cast<const uint32_t *>(S.dec);
            uint32_t *gw32 = reinterpret_cast<uint32_t *>(dec + (s - gOff));
            for (int k = lane; k < nWords; k += 32) {
                const uint32_t w = sw[k];
                uint32_t vm = ONES;  // valid bytes of this word
                if (4 * k < gOff || 4 * k + 4 > gOff + n) {
                    vm = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) if (4 * k + b >= gOff && 4 * k + b < gOff + n) vm |= 1u << (8 * b);
                }
                const uint32_t at = (w >> 6) & ONES, dn = (w >> 7) & ONES, dead = at | dn;
                const uint32_t code = (w >> DEC_MOVE_SHIFT) & 0x07070707u;
                const uint32_t nz = ((code + 0x07070707u) >> 3) & ONES;       // move code != 0
                const uint32_t born = (w >> 2) & matesM & ~(at & bornVoid) & ONES;
                const uint32_t alive = ~dead & vm;
                const uint32_t out = alive & nz;
                const uint32_t d7 = (dead << 3) - dead;                        // 7 in every dead byte
                const uint32_t fin = (w & 0x03030303u) | (born << 2) | ((code | d7) << DEC_MOVE_SHIFT);
                stayL += __popc(alive & ~nz);
                bornL += __popc(born & vm);
                moveL += __popc(nz & ~(at & moveVoid) & vm);                   // registered moves (core/SPopulation.cpp:1067)
                outL += __popc(out);
                if (vm == ONES) {
                    gw32[k] = fin;
                } else {
#pragma unroll
                    for (int b = 0; b < 4; b++) if (vm & (1u << (8 * b))) dec[s - gOff + 4 * k + b] = (uint8_t)(fin >> (8 * b));
                }
                if (out) {
#pragma unroll
                    for (int b = 0; b < 4; b++) if (out & (1u << (8 * b))) atomicAdd(&S.outC[((code >> (8 * b)) & 7) - 1], 1);
                }
            }
        }
        const int stayC = __reduce_add_sync(FULL, stayL), bornC = __reduce_add_sync(FULL, bornL);
        const int outC = __reduce_add_sync(FULL, outL);
        nMove += __reduce_add_sync(FULL, moveL + confL);
        nDead += n - stayC - outC;
        nBorn += bornC;
        __syncwarp();
        if (bornC > MAXMOTHERS && lane == 0) atomicExch(&st->oversize, 1);
        if (lane == 0) { stay[c] = stayC; birthCount[c] = bornC; }  // a cell belongs to exactly one warp: plain stores
        // the movers towards neighbour k take the slots [base, base+cnt) of that cell's arrivals: pass 2 places them
        // without atomics
        if (lane < MAXN) {
            if (pendCell >= 0) moveBase[(size_t)pendCell * MOVE_STRIDE + lane] = pendBase;
            const int cnt = S.outC[lane];
            pendBase = cnt ? atomicAdd(&arrive[S.nbrC[lane]], cnt) : 0;
        }
        pendCell = c;
        __syncwarp();
    }
    }
    if (lane < MAXN && pendCell >= 0) moveBase[(size_t)pendCell * MOVE_STRIDE + lane] = pendBase;
    if (lane == 0) {
        if (nDead) atomicAdd(&st->nDeaths, nDead);
        if (nMove) atomicAdd(&st->nMoves, nMove);
        if (nBorn) atomicAdd(&st->nBirths, nBorn);
    }
}

// ---------------------------------------------------------------------------------------------
// pass 2: counting-sort scatter (performMoves core/SPopulation.cpp:1058-1092) + newborns
// (makeOffspring / createAgentAtIndex :823-847,880-918; makePopSpecificOffspring populations/tut_EnvironAltPop.cpp:141-149)

// ---- multi-GPU: the grid is sharded by contiguous cell ranges, one range per rank ------------------
// An agent that moves into a cell of another rank is not written to the local buffer but packed into the send
// buffer of the owning rank; the receiver places it with k_place_migrants.  Environment arrays are replicated.
struct Migrant {  // 32 bytes
    long long id;
    float birth, lastBirth, age;
    int cell;
    unsigned flags;  // bits 0-7: the agent's flag byte; bits 8-31: m_iNumBabies (populations with Genetics)
    unsigned pad;    // peer-memory exchange: slot among the arrivals of `cell`
};

// Exchange over peer memory (NVLink): every rank owns one XchgBlock and one double-buffered array of remote arrival
// counts, both mapped into every other rank's address space (cudaIpc).  Per step: each rank adds its arrivals into the
// owners' counters (atomicAdd_system returns the first slot), announces its births, and after a cross-GPU barrier
// writes the records of the agents that leave straight into the owner's receive buffer.  No host round trip.
constexpr int MAXR = 16;
struct XchgBlock {
    unsigned flagA[MAXR];   // stamp of rank r: its arrival counts and births of this step are in
    unsigned flagB[MAXR];   // stamp of rank r: its migrant records of this step are in
    long long births[MAXR];
    int recvCount;          // migrant records reserved in recv[] this step
    int recvCap;            // records the owner's buffer holds (read by the peers when they connect)
    int rowWords;           // 64-bit words per genome row (0: no Genetics); the rows follow the records, one per record slot
    int pad[29];
};
static_assert(sizeof(XchgBlock) % 32 == 0, "the migrant records follow the header");
struct PeerTable {
    int *arriveRemote[MAXR];  // [2][nCells] per rank
    XchgBlock *x[MAXR];
    int recvCap[MAXR];        // capacity of every rank's receive buffer (the OWNER's, not the sender's)
};

struct ShardArgs {
    int on;                 // 0: single GPU
    int rank, nranks;
    int c0, c1;             // owned cells [c0, c1)
    const int *cellBegin;   // nranks+1 range boundaries (device)
    Migrant *sendBuf;       // NCCL mode: packed by destination rank, sendOff[q] .. sendOff[q+1]
    const int *sendOff;
    int *sendCursor;        // nranks counters
    long long birthOffset;  // births of the lower ranks this step (newborn ids are global ranks); p2p: in DevStats
    int p2p;                // 1: records go straight into the owner's receive buffer
    int recvCap;            // records of this rank's own receive buffer
    const int *remoteBase;  // per foreign halo cell: first arrival slot of this rank's movers in the owner's cell
    const PeerTable *peers; // device copy
    // populations with Genetics: the genome row travels with the agent
    const unsigned long long *pool;  // this rank's genome pool
    int rowWords;                    // 64-bit words per row
    unsigned long long *sendGenomes; // NCCL mode: rows packed like sendBuf
};

__device__ __forceinline__ Migrant *xchg_recv(XchgBlock *x) { return reinterpret_cast<Migrant *>(x + 1); }
// the genome rows of the received agents follow the `cap` record slots of the owner's buffer
__device__ __forceinline__ unsigned long long *xchg_genomes(XchgBlock *x, int cap) {
    return reinterpret_cast<unsigned long long *>(xchg_recv(x) + cap);
}

__device__ __forceinline__ int shard_owner(const ShardArgs &H, int c) {
    int q = 0;
    while (q + 1 < H.nranks && c >= H.cellBegin[q + 1]) q++;
    return q;
}

// One leaver per calling lane (the lanes of `who`, a ballot of the callers): its record goes to the owner of cell m.cell,
// for populations with Genetics followed by its genome row, copied by the whole warp.  Slots of the owner's receive buffer are
// reserved with ONE remote atomic per (warp, owner), not one per agent.  ALL lanes of the warp must call (who may be 0).
template <bool GEN>
__device__ __forceinline__ void send_leavers(const ShardArgs &H, unsigned who, bool leaves, Migrant m, int srcRow) {
    const unsigned FULL = 0xffffffffu;
    const int lane = (int)(threadIdx.x & 31);
    if (who == 0) return;
    int qo = -1, slot = -1;
    if (leaves) {
        qo = shard_owner(H, m.cell);
        if (H.p2p) {
            const unsigned peers = __match_any_sync(who, qo);
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd_system(&H.peers->x[qo]->recvCount, __popc(peers));
            base = __shfl_sync(peers, base, leader);
            slot = base + __popc(peers & lanemask_lt());
            if (slot < H.peers->recvCap[qo]) xchg_recv(H.peers->x[qo])[slot] = m;  // the count still grows: the owner sees the overflow
            else slot = -1;
        } else {
            slot = H.sendOff[qo] + atomicAdd(&H.sendCursor[qo], 1);
            H.sendBuf[slot] = m;
        }
    }
    if constexpr (GEN) {
        unsigned todo = __ballot_sync(FULL, leaves && slot >= 0);
        while (todo) {
            const int L = __ffs(todo) - 1;
            todo &= todo - 1;
            const int q = __shfl_sync(FULL, qo, L), sl = __shfl_sync(FULL, slot, L), sr = __shfl_sync(FULL, srcRow, L);
            unsigned long long *dst = H.p2p ? xchg_genomes(H.peers->x[q], H.peers->recvCap[q]) + (size_t)sl * H.rowWords
                                            : H.sendGenomes + (size_t)sl * H.rowWords;
            const unsigned long long *src = H.pool + (size_t)sr * H.rowWords;
            for (int w = lane; w < H.rowWords; w += 32) dst[w] = src[w];
        }
    }
}


// Arrivals cross a shard boundary only in the halo: the cells with a neighbour owned by another rank (the list is the
// same on every rank, built in qhgb_comm_init).  Their arrival counts travel as one compact array that is summed over
// the ranks; this kernel fills it, counts what this rank sends to every other rank (info[q]) and clears the foreign
// cells' local counters.  info[] is zeroed before the launch.
constexpr int MAX_RANKS_SMEM = 64;
__global__ void __launch_bounds__(256)
k_halo_gather(int nHalo, const int *__restrict__ halo, const int *__restrict__ cellBegin, int rank, int nranks,
              int *__restrict__ arrive, int *__restrict__ cursor, int *__restrict__ buf, const DevStats *__restrict__ st,
              int *__restrict__ info) {
    __shared__ int sInfo[MAX_RANKS_SMEM];
    const bool useSmem = nranks <= MAX_RANKS_SMEM;
    if (useSmem) {
        for (int q = threadIdx.x; q < nranks; q += blockDim.x) sInfo[q] = 0;
        __syncthreads();
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nHalo; i += gridDim.x * blockDim.x) {
        const int c = halo[i];
        const int v = arrive[c];
        buf[i] = v;
        int q = 0;
        while (q + 1 < nranks && c >= cellBegin[q + 1]) q++;
        if (q != rank) {
            if (v) atomicAdd(useSmem ? &sInfo[q] : &info[q], v);
            arrive[c] = 0;
        } else {
            cursor[c] = v;  // the local movers hold the slots [0, v) of the cell's arrivals; migrants follow
        }
    }
    if (useSmem) {
        __syncthreads();
        for (int q = threadIdx.x; q < nranks; q += blockDim.x) if (sInfo[q]) atomicAdd(&info[q], sInfo[q]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) info[nranks] = st->oversize ? -1 : st->nBirths;  // -1: a cell beyond the limits of pass 1, see k_halo_push
}

// after the all-reduce of the halo array: the owned halo cells take the sum over all ranks
__global__ void k_halo_apply(int nHalo, const int *__restrict__ halo, int c0, int c1, const int *__restrict__ buf,
                             int *__restrict__ arrive) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nHalo; i += gridDim.x * blockDim.x) {
        const int c = halo[i];
        if (c >= c0 && c < c1) arrive[c] = buf[i];
    }
}

// ---- exchange over peer memory ----------------------------------------------------------------------------
// (1) arrivals into the cells of other ranks: one remote atomicAdd per halo cell, the returned slot is kept for pass 2;
//     own halo cells: the counters the peers will use in the NEXT step are cleared.  Births are announced to every rank.
__global__ void __launch_bounds__(256)
k_halo_push(int nHalo, const int *__restrict__ halo, const int *__restrict__ cellBegin, int rank, int nranks, int nCells,
            int parity, int *__restrict__ arrive, int *__restrict__ remoteBase, const PeerTable *__restrict__ T,
            DevStats *__restrict__ st) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nHalo; i += gridDim.x * blockDim.x) {
        const int c = halo[i];
        int q = 0;
        while (q + 1 < nranks && c >= cellBegin[q + 1]) q++;
        if (q != rank) {
            const int v = arrive[c];
            if (v) {
                remoteBase[c] = atomicAdd_system(T->arriveRemote[q] + (size_t)parity * nCells + c, v);
                arrive[c] = 0;
            }
        } else {
            T->arriveRemote[rank][(size_t)(parity ^ 1) * nCells + c] = 0;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < nranks) {
        // (a rank whose pass 1 met a cell beyond its limits says so instead: every rank then leaves the step undone and all of
        // them redo it with the recovery kernels)
        T->x[threadIdx.x]->births[rank] = st->oversize ? -1ll : (long long)st->nBirths;
        if (threadIdx.x == 0) T->x[rank]->recvCount = 0;  // nobody reserves records before the barrier that follows
    }
}

// (2) cross-GPU barrier: every rank tells every other "I am at `stamp`" and waits for the others' words.  A rank that does not
//     show up within the time limit (QHG_XBARRIER_TIMEOUT_S) raises commError and halt: the host fails the step instead of hanging.
// the waiting half of a cross-GPU barrier, executed by every block of a kernel that comes after it: block 0 tells the peers
// "this rank is at `stamp`", every block waits until all peers have said so (the words live in this rank's own memory)
__device__ __forceinline__ void xbarrier_inline(int rank, int nranks, int which, unsigned stamp, const PeerTable *__restrict__ T,
                                                DevStats *__restrict__ st, long long timeoutClocks) {
    const int r = threadIdx.x;
    if (r < nranks) {
        if (blockIdx.x == 0) {
            __threadfence_system();
            volatile unsigned *theirs = which ? &T->x[r]->flagB[rank] : &T->x[r]->flagA[rank];
            *theirs = stamp;
        }
        volatile unsigned *mine = which ? &T->x[rank]->flagB[r] : &T->x[rank]->flagA[r];
        const long long t0 = clock64();
        while ((int)(*mine - stamp) < 0) {
            if (clock64() - t0 > timeoutClocks) { st->commError = 1; st->halt = 1; break; }
            __nanosleep(100);
        }
        __threadfence_system();
    }
    __syncthreads();
}

// (2)+(3) in one launch: the barrier, then the owned halo cells take the arrivals of the other ranks (the local movers hold the
// first slots of a cell's arrivals, the migrants follow); block 0 also derives this rank's id offset from everybody's births
__global__ void __launch_bounds__(256)
k_xbarrier_merge(int nHalo, const int *__restrict__ halo, int c0, int c1, int nCells, int parity, const PeerTable *__restrict__ T,
                 int rank, int nranks, unsigned stamp, int *__restrict__ arrive, int *__restrict__ cursor, DevStats *__restrict__ st,
                 long long timeoutClocks) {
    xbarrier_inline(rank, nranks, 0, stamp, T, st, timeoutClocks);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long below = 0, total = 0;
        bool peerOversize = false;
        for (int q = 0; q < nranks; q++) {
            const long long b = ((volatile long long *)T->x[rank]->births)[q];
            peerOversize |= b < 0;
            if (q < rank) below += b;
            total += b;
        }
        if (peerOversize) st->oversize = 1;  // the kernels that follow do nothing, the step's end raises `halt`
        st->birthOffset = below;
        st->globalBirths = total;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nHalo; i += gridDim.x * blockDim.x) {
        const int c = halo[i];
        if (c >= c0 && c < c1) {
            const int local = arrive[c];
            cursor[c] = local;
            arrive[c] = local + ((volatile int *)T->arriveRemote[rank])[(size_t)parity * nCells + c];
        }
    }
}

// (4) after barrier B: the records the other ranks wrote into this rank's receive buffer go to their slots
// GEN: one warp per record -- the agent takes a genome row of this rank's pool (the rows after those of the step's births, in
// the order k_make_offspring hands them out: top of the free stack, then the unused tail) and its genome is copied in
template <bool GEN>
__global__ void k_place_migrants_p2p(DevStats *__restrict__ st, const PeerTable *__restrict__ T, int rank, int recvCap, AgentArrays o,
                                     const int *__restrict__ newStart, const int *__restrict__ stay, const int *__restrict__ cursor,
                                     int storeAge, const GenomeCtl *__restrict__ ctl = nullptr, const int *__restrict__ freeStack = nullptr,
                                     unsigned long long *__restrict__ pool = nullptr, int rowWords = 0, int poolRows = 0,
                                     int nranks = 0, unsigned stamp = 0, long long timeoutClocks = 0, int stepEnd = 0, int advanceStep = 0) {
    // nranks > 0: barrier B happens here (every block waits for the peers' "my records are in"), not in a kernel of its own
    if (nranks > 0) xbarrier_inline(rank, nranks, 1, stamp, T, st, timeoutClocks);
    const bool skip = st->overflow || st->oversize || st->halt;
    const int n = skip ? 0 : ((volatile int *)&T->x[rank]->recvCount)[0];
    if (!skip && blockIdx.x == 0 && threadIdx.x == 0) {
        st->nRecv = n;
        if (n > recvCap) { st->commError = 2; st->halt = 1; }  // the step is void (the book-keeping below keeps the old state)
    }
    const Migrant *in = xchg_recv(T->x[rank]);
    const int per = GEN ? 32 : 1;  // threads per record
    const int lane = GEN ? (int)(threadIdx.x & 31) : 0;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) / per; i < min(n, recvCap); i += (gridDim.x * blockDim.x) / per) {
        const Migrant m = in[i];
        const int pos = newStart[m.cell] + stay[m.cell] + cursor[m.cell] + (int)m.pad;
        if (lane == 0) {
            o.id[pos] = m.id;
            o.birth[pos] = m.birth;
            o.lastBirth[pos] = m.lastBirth;
            o.flags[pos] = (uint8_t)(m.flags & 0xffu);
            if (storeAge) o.age[pos] = m.age;
        }
        if constexpr (GEN) {
            const int e = ctl->nBirths + i, nFree0 = ctl->nFree;
            const int slot = (e < nFree0) ? freeStack[nFree0 - 1 - e] : ctl->hwm + (e - nFree0);
            if (slot >= poolRows) { if (lane == 0) st->overflow = 1; continue; }  // k_step_end turns it into `halt`
            if (lane == 0) { o.gslot[pos] = slot; o.nbabies[pos] = (int)(m.flags >> 8); }
            const unsigned long long *src = xchg_genomes(T->x[rank], recvCap) + (size_t)i * rowWords;
            unsigned long long *dst = pool + (size_t)slot * rowWords;
            for (int w = lane; w < rowWords; w += 32) dst[w] = src[w];
        }
    }
    if (stepEnd) {  // this is the step's last kernel: the block that finishes last does the book-keeping (k_step_end's)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&st->doneBlocks, 1) == (int)gridDim.x - 1) {
                st->doneBlocks = 0;
                __threadfence();
                step_end_body(st, advanceStep, -2);
            }
        }
    }
}

// the agents Navigate sent far away (k_cell_decide<.., true>): written at their slot in the destination cell; their decision
// byte goes back to "stays" so that k_free_genomes_dec does not take them for dead
// Sharded runs: a destination on another rank (far jumps go anywhere; their destination cells are part of the halo list, so the
// arrival slots were exchanged like those of the boundary cells) -- the record, and the genome row, go to the owner like those of
// the neighbour movers; the decision byte stays "gone" so that the row is freed here.
template <bool GEN>
__global__ void k_place_jumpers(DevStats *__restrict__ st, const int *__restrict__ jumpCount, const JumpEntry *__restrict__ jumps,
                                int jumpCap, AgentArrays a, AgentArrays o, const int *__restrict__ newStart,
                                const int *__restrict__ stay, uint8_t *__restrict__ dec, int storeAge, ShardArgs H) {
    if (st->overflow || st->oversize || st->halt) return;
    const unsigned FULL = 0xffffffffu;
    const int n = min(*jumpCount, jumpCap);
    int nSentL = 0;
    for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {  // whole warps: send_leavers is cooperative
        const int i = i0 + (int)threadIdx.x;
        const bool act = i < n;
        JumpEntry e{};
        if (act) e = jumps[i];
        const bool leaves = act && H.on && (e.to < H.c0 || e.to >= H.c1);
        const unsigned who = H.on ? __ballot_sync(FULL, leaves) : 0u;
        Migrant m{};
        int srcRow = 0;
        if (act) {
            // the jump entry was written before the pairing was settled: whether the agent gave birth is in the committed byte
            const uint8_t fin = dec[e.src];
            if (leaves) {
                m.id = a.id[e.src]; m.birth = a.birth[e.src]; m.lastBirth = a.lastBirth[e.src];
                m.age = storeAge ? a.age[e.src] : 0.0f;
                m.cell = e.to;
                m.flags = (unsigned)(fin & (F_MALE | F_FERTILE));
                m.pad = (unsigned)((H.p2p ? H.remoteBase[e.to] : 0) + e.slot);
                if constexpr (GEN) {
                    srcRow = a.gslot[e.src];
                    m.flags |= (unsigned)(a.nbabies[e.src] + ((fin & F_BORN) ? 1 : 0)) << 8;
                }
                nSentL++;
            } else {
                const int pos = newStart[e.to] + stay[e.to] + e.slot;
                o.id[pos] = a.id[e.src];
                o.birth[pos] = a.birth[e.src];
                o.lastBirth[pos] = a.lastBirth[e.src];
                o.flags[pos] = (uint8_t)(fin & (F_MALE | F_FERTILE));
                if (storeAge) o.age[pos] = a.age[e.src];
                if constexpr (GEN) {
                    o.gslot[pos] = a.gslot[e.src];
                    o.nbabies[pos] = a.nbabies[e.src] + ((fin & F_BORN) ? 1 : 0);
                }
                dec[e.src] = (uint8_t)(fin & 7);  // not dead: k_free_genomes_dec keeps its row
            }
        }
        if (who) send_leavers<GEN>(H, who, leaves, m, srcRow);
    }
    if (H.on) {
        nSentL = __reduce_add_sync(FULL, nSentL);
        if ((threadIdx.x & 31) == 0 && nSentL) atomicAdd(&st->nSent, nSentL);
    }
}

// the per-agent cell index is implied by cellStart on the fast path; this writes it out when somebody needs it
// (generic path, records handed back to the host)
__global__ void __launch_bounds__(256)
k_fill_cells(int cLo, int cHi, const int *__restrict__ cellStart, int *__restrict__ cell) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    for (int c = cLo + gw; c < cHi; c += nW) {  // a shard holds agents in its own cells only
        const int s = cellStart[c], e = cellStart[c + 1];
        for (int i = s + lane; i < e; i += 32) cell[i] = c;
    }
}

template <bool GEN>
__global__ void k_place_migrants(DevStats *__restrict__ st, const Migrant *__restrict__ in, int n, AgentArrays o,
                                 const int *__restrict__ newStart, const int *__restrict__ stay, int *__restrict__ cursor,
                                 int storeAge, const GenomeCtl *__restrict__ ctl = nullptr, const int *__restrict__ freeStack = nullptr,
                                 unsigned long long *__restrict__ pool = nullptr, const unsigned long long *__restrict__ inGenomes = nullptr,
                                 int rowWords = 0, int poolRows = 0) {
    if (st->overflow || st->oversize) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->nRecv = n;
    const int per = GEN ? 32 : 1;
    const int lane = GEN ? (int)(threadIdx.x & 31) : 0;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) / per; i < n; i += (gridDim.x * blockDim.x) / per) {
        const Migrant m = in[i];
        int pos = 0;
        if (lane == 0) pos = newStart[m.cell] + stay[m.cell] + atomicAdd(&cursor[m.cell], 1);
        if (lane == 0) {
            o.id[pos] = m.id;
            o.birth[pos] = m.birth;
            o.lastBirth[pos] = m.lastBirth;
            o.flags[pos] = (uint8_t)(m.flags & 0xffu);
            if (storeAge) o.age[pos] = m.age;
        }
        if constexpr (GEN) {
            const int e = ctl->nBirths + i, nFree0 = ctl->nFree;
            const int slot = (e < nFree0) ? freeStack[nFree0 - 1 - e] : ctl->hwm + (e - nFree0);
            if (slot >= poolRows) { if (lane == 0) st->overflow = 1; continue; }
            if (lane == 0) { o.gslot[pos] = slot; o.nbabies[pos] = (int)(m.flags >> 8); }
            const unsigned long long *src = inGenomes + (size_t)i * rowWords;
            unsigned long long *dst = pool + (size_t)slot * rowWords;
            for (int w = lane; w < rowWords; w += 32) dst[w] = src[w];
        }
    }
}

// ---- pass 2 with bulk-copy staging -------------------------------------------------------------------------------
// The agents of a batch of consecutive cells are one contiguous range of every array.  The warp walks it in windows
// of SCH agents; lane 0 brings each window into shared memory with four 1-D bulk copies (cp.async.bulk, completion on
// an mbarrier), SNST windows ahead, so the bytes in flight do not depend on registers or occupancy.  Windows start
// on a multiple of 16 agents: every source address and size is a multiple of 16 bytes.
#ifndef QHG_SCH
#define QHG_SCH 384
#endif
#ifndef QHG_SNST
#define QHG_SNST 1
#endif
constexpr int SCH_DENSE = QHG_SCH;   // agents per window for dense populations (6 CTAs per SM)
constexpr int SCH_SPARSE = 256;      // ... and for sparse ones: smaller windows, 8 CTAs per SM (measured: -12 % at 20 agents per cell, +8 % at 150)
constexpr int SNST = QHG_SNST;

template <int SCH>
struct alignas(128) StagedAgents {
    int64_t id[SCH];
    float birth[SCH];
    float lastBirth[SCH];
    uint8_t dec[SCH];
};
template <int SCH, int NST, int MM = MAXMOTHERS, int SG = CELL_BATCH>
struct alignas(128) WarpSmemS {
    StagedAgents<SCH> win[NST];
    int64_t motherId[MM + BIRTH_FLUSH + 8];  // the mothers of the cells since the last placement of babies (a cell adds at most MM)
    uint8_t motherC[MM + BIRTH_FLUSH + 8];   // ... and the cell of the grab each belongs to
    static constexpr int MVC = (SCH >= 384) ? MVCAP : 64;
    uint16_t mvJ[MVC];                       // queued movers: position in the window
    uint8_t mvC[MVC];                        // ... and cell of the grab
    uint16_t dirOff[SG][8];                  // movers of (cell, direction) placed so far
    unsigned long long bar[NST];
};
template <int SCH, int NST, int MM = MAXMOTHERS, int SG = CELL_BATCH>
struct alignas(128) WarpSmemSG {  // populations with Genetics: the mothers' positions in the old buffer as well
    StagedAgents<SCH> win[NST];
    int64_t motherId[MM + BIRTH_FLUSH + 8];
    int motherIdx[MM + BIRTH_FLUSH + 8];
    uint8_t motherC[MM + BIRTH_FLUSH + 8];
    static constexpr int MVC = (SCH >= 384) ? MVCAP : 64;
    uint16_t mvJ[MVC];
    uint8_t mvC[MVC];
    uint16_t dirOff[SG][8];
    unsigned long long bar[NST];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#ifndef QHG_SCATTER_S_MINB
#define QHG_SCATTER_S_MINB 6
#endif
constexpr int SCATTER_CTAS_DENSE = QHG_SCATTER_S_MINB, SCATTER_CTAS_SPARSE = 8;  // persistent grids: this many CTAs per SM
// GEN = true: the population has Genetics -- the genome handle and m_iNumBabies follow the agent (read straight from global
// memory, like the optional age), every newborn leaves a birth record (baby position, mother, father) for k_make_offspring
// SG = cells per grab of the work counter: their agents are ONE contiguous range of every array, streamed through NST windows
// of SCH agents -- the copies of the next window run while this one is consumed, and only the first window of a grab is waited
// for with nothing else in flight (with 4-cell grabs that was every window at 20 agents per cell and every second one at 150)
// BIG = the recovery variant (see k_seg_decide<..., BIG>): up to 2048 births per cell, per-warp slices in dynamic shared memory
constexpr int MAXMOTHERS_BIG = 2048;
extern __shared__ __align__(128) unsigned char qhg_dyn_smem_s[];
// (Taking the NEXT grab while this one is worked on and handing the freed stages to its first windows -- so that a warp never
// waits for a copy with nothing else in flight -- was built and measured: no gain at 150 agents per cell, -13 % at 22,
// profiles/ab_scatter_r02c.txt.  The other warps of the SM already cover those waits.)
template <bool GEN = false, int SCH = SCH_DENSE, int MINB = SCATTER_CTAS_DENSE, int SG = CELL_BATCH, int NST = SNST, bool BIG = false>
__global__ void __launch_bounds__(CW * 32, MINB)
k_cell_scatter(DevStats *__restrict__ st, AgentArrays a, AgentArrays o, int cLo, int cHi, const int *__restrict__ cellStart,
               const uint8_t *__restrict__ dec, const int *__restrict__ nbr, const int *__restrict__ newStart,
               const int *__restrict__ stay, const int *__restrict__ arrive, const int *__restrict__ moveBase,
               const int *__restrict__ birthBase, float t, int storeAge, int femaleOnly, RngKey key, ShardArgs H,
               const int *__restrict__ father = nullptr, BirthEntry *__restrict__ births = nullptr, GenomeCtl *__restrict__ gctl = nullptr,
               uint8_t *decMark = nullptr, int shrink = 0, int stepEnd = 0, int advanceStep = 0) {
    static_assert(SG + 1 <= 32, "one lane per cell start of the grab");
    // stepEnd: this is the step's last kernel (one GPU, no Genetics, no Navigate): the block that finishes last does the step's
    // book-keeping (k_step_end's), also when the step has already failed
    auto step_end = [&]() {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&st->doneBlocks, 1) == (int)gridDim.x - 1) {
                st->doneBlocks = 0;
                __threadfence();
                step_end_body(st, advanceStep, -1);
            }
        }
    };
    constexpr int MM = BIG ? MAXMOTHERS_BIG : MAXMOTHERS;
    using WSS = typename std::conditional<GEN, WarpSmemSG<SCH, NST, MM, SG>, WarpSmemS<SCH, NST, MM, SG>>::type;
    WSS *smem;
    if constexpr (BIG) {
        smem = reinterpret_cast<WSS *>(qhg_dyn_smem_s);
    } else {
        __shared__ WSS smemStatic[CW];
        smem = smemStatic;
    }
    if (st->overflow || st->oversize || st->halt) { if (stepEnd) step_end(); return; }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#if QHG_OPAQUE_SMEM
    WSS *Sp = &smem[wid];  // kept in a register (see k_seg_decide)
    asm volatile("" : "+l"(Sp));
    __builtin_assume(__isShared(Sp));
    WSS &S = *Sp;
#else
    WSS &S = smem[wid];
#endif
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = lanemask_lt();
    const unsigned step = st->step;
    const long long nextID = st->nextID;
    const long long birthOffset = H.p2p ? st->birthOffset : H.birthOffset;
    int nSentL = 0;
    if (lane == 0) {
        for (int k = 0; k < NST; k++) mbar_init(&S.bar[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0;  // parity of the next completion of every stage's barrier
    const int nWarpsS = gridDim.x * CW;
    int lastEnd = cLo;
    int wIss = 0, wCons = 0;  // windows issued / consumed so far by this warp: stage = count % NST, at most NST in flight
    // a grab: lane l keeps the numbers of cell cBase+l (lane SG-or-less: the end of the grab)
    int cBase = 0, cEnd = 0, csL = 0, nsL = 0, arL = 0, bbL = 0;
    int curIssued = 0;        // windows of the current grab that are in flight
    auto fetch = [&](int &B, int &E, int &cs, int &ns_, int &ar, int &bb) -> bool {
        // (short ranges -- the shards of a many-GPU run -- get grabs that shrink towards the end: no tail of a whole batch)
        const int g = shrink ? max(1, min(SG, (cHi - lastEnd) / (2 * nWarpsS))) : SG;
        int b = 0;
        if (lane == 0) b = cLo + atomicAdd(&st->workScatter, g);
        b = __shfl_sync(FULL, b, 0);
        if (b >= cHi) return false;
        lastEnd = b + g;
        B = b; E = min(b + g, cHi);
        cs = 0; ns_ = 0; ar = 0; bb = 0;
        if (b + lane <= E) cs = cellStart[b + lane];
        if (b + lane < E) { ns_ = newStart[b + lane]; ar = arrive[b + lane]; bb = birthBase[b + lane]; }
        return true;
    };
    auto issueWin = [&](int g0w, int gew, int k) {  // start the copies of window k of the grab whose agents are [.., gew), first window at g0w
        if (lane == 0) {
            const int w0 = g0w + k * SCH;
            const int cnt = min(SCH, (gew - w0 + 15) & ~15);
            StagedAgents<SCH> &W = S.win[wIss % NST];
            unsigned long long *bar = &S.bar[wIss % NST];
            mbar_expect_tx(bar, (uint32_t)cnt * 17u);
            bulk_g2s(W.id, a.id + w0, (uint32_t)cnt * 8u, bar);
            bulk_g2s(W.birth, a.birth + w0, (uint32_t)cnt * 4u, bar);
            bulk_g2s(W.lastBirth, a.lastBirth + w0, (uint32_t)cnt * 4u, bar);
            bulk_g2s(W.dec, dec + w0, (uint32_t)cnt, bar);
        }
        wIss++;
    };
    bool have = fetch(cBase, cEnd, csL, nsL, arL, bbL);
    while (have) {
        const int gs = __shfl_sync(FULL, csL, 0), ge = __shfl_sync(FULL, csL, cEnd - cBase);
        if (ge == gs) {  // sea
            have = fetch(cBase, cEnd, csL, nsL, arL, bbL);
            curIssued = 0;
            continue;
        }
        const int g0 = gs & ~15;
        const int nWin = (ge - g0 + SCH - 1) / SCH;
        while (curIssued < min(nWin, NST)) { issueWin(g0, ge, curIssued); curIssued++; }

        int ci = 0;  // cell of the grab the walk is in
        int s = gs, e = __shfl_sync(FULL, csL, 1);
        while (ci < cEnd - cBase && e == s) { ci++; s = e; e = __shfl_sync(FULL, csL, min(ci + 1, 31)); }
        int ns = 0, stayBase = 0, nMothers = 0, nmv = 0, cellM0 = 0;
        int mFirstL = 0, mCntL = 0, babyBaseL = 0;  // lane c: the mothers of cell c in the list, where its babies go
        for (int i = lane; i < SG * 8; i += 32) (&S.dirOff[0][0])[i] = 0;
        __syncwarp();
        auto begin_cell = [&]() {
            ns = __shfl_sync(FULL, nsL, ci);
            stayBase = 0;
            cellM0 = nMothers;
        };
        // the babies of the cells collected so far: newborn id = nextID + rank of (cell, mother id) among this step's births; the
        // same rank places the baby.  One pass for the mothers of several cells (a pass per cell leaves most lanes idle where
        // cells hold a few dozen agents).  Only called between two cells: the list holds whole cells.
        auto place_babies = [&]() {
            __syncwarp();
            for (int m0 = 0; m0 < nMothers; m0 += 32) {
                const int m = m0 + lane;
                const bool act = m < nMothers;
                const int c = act ? S.motherC[m] : 0;
                const int first = __shfl_sync(FULL, mFirstL, c), cnt = __shfl_sync(FULL, mCntL, c);
                const int babyBase = __shfl_sync(FULL, babyBaseL, c), bb = __shfl_sync(FULL, bbL, c);
                if (act) {
                    const int64_t mid = S.motherId[m];
                    int r = 0;
                    for (int q = first; q < first + cnt; q++) r += (S.motherId[q] < mid) ? 1 : 0;
                    const int64_t cid = nextID + birthOffset + bb + r;
                    const uint32_t gnd = agent_draws(cid, step, STREAM_BABY, key).x >> 31;  // (uchar)(2*wrandd())
                    const int pos = babyBase + r;
                    o.id[pos] = cid;
                    o.birth[pos] = t;
                    o.lastBirth[pos] = 0.0f;
                    // females are born FERTILE, core/SPopulation.cpp:895-898; tut_ParthenoPop turns the drawn males into females
                    o.flags[pos] = (uint8_t)(gnd ? (femaleOnly ? 0 : F_MALE) : F_FERTILE);
                    if (storeAge) o.age[pos] = 0.0f;
                    if constexpr (GEN) {  // the genome is made by k_make_offspring from the parents' rows in the old buffer
                        o.nbabies[pos] = 0;
                        const int mi = S.motherIdx[m];
                        record_birth(births, gctl, pos, mi, father[mi], cid);
                    }
                }
            }
            nMothers = 0;
            __syncwarp();
        };
        begin_cell();
        for (int k = 0; k < nWin; k++) {
            const int w0 = g0 + k * SCH, w1 = min(w0 + SCH, ge);
            const int stg = wCons % NST;
            const StagedAgents<SCH> &W = S.win[stg];
            mbar_wait(&S.bar[stg], (phase >> stg) & 1u);
            phase ^= 1u << stg;
            // the queued movers (of any cell of the grab; their records are in this window): every lane looks up where ITS mover
            // goes -- the neighbour, that cell's new start and stayers, the slot pass 1 reserved for (cell, direction) -- and takes
            // the next free place of its (cell, direction)
            auto flush_movers = [&]() {
                __syncwarp();
                for (int q0 = 0; q0 < nmv; q0 += 32) {
                    const int q = q0 + lane;
                    const bool act = q < nmv;
                    const int x = act ? S.mvJ[q] : 0, cx = act ? S.mvC[q] : 0;
                    const uint8_t v = act ? W.dec[x] : (uint8_t)0;
                    const int dir = act ? (v >> DEC_MOVE_SHIFT) - 1 : 0;
                    const int c = cBase + cx;
                    const int d = act ? nbr[(size_t)c * MAXN + dir] : 0;
                    const bool leaves = act && H.on && (d < H.c0 || d >= H.c1);  // into a cell of another rank
                    int base = 0;
                    if (act) {
                        const int mb = moveBase[(size_t)c * MOVE_STRIDE + dir];
                        if (!leaves) base = newStart[d] + stay[d] + mb;
                        else if (H.p2p) base = H.remoteBase[d] + mb;  // slot among the arrivals of the owner's cell (k_halo_push)
                    }
                    // slots were reserved per (cell, direction) in pass 1: the movers of one take consecutive places
                    const unsigned peers = __match_any_sync(FULL, act ? ((cx << 3) | dir) : (0x1000 | lane));
                    const int off = act ? (int)S.dirOff[cx][dir] : 0;
                    __syncwarp();
                    if (act && lane == __ffs(peers) - 1) S.dirOff[cx][dir] = (uint16_t)(off + __popc(peers));
                    const int pos = base + off + __popc(peers & lt);
                    const unsigned who = H.on ? __ballot_sync(FULL, leaves) : 0u;
                    Migrant m{};
                    int srcRow = 0;
                    if (act) {
                        const int64_t id = W.id[x];
                        const float birth = W.birth[x], lastBirth = W.lastBirth[x];
                        const float age = storeAge ? a.age[w0 + x] : 0.0f;
                        if (leaves) {  // the record goes to the owner of cell d (straight into its receive buffer over NVLink, or packed for NCCL)
                            m.id = id; m.birth = birth; m.lastBirth = lastBirth; m.age = age; m.cell = d;
                            m.flags = (unsigned)(v & (F_MALE | F_FERTILE)); m.pad = (unsigned)pos;
                            if constexpr (GEN) {
                                srcRow = a.gslot[w0 + x];
                                m.flags |= (unsigned)(a.nbabies[w0 + x] + ((v & F_BORN) ? 1 : 0)) << 8;
                                // its genome row is free once the step's births have read their parents: for k_free_genomes_dec
                                // the agent is as good as dead
                                decMark[w0 + x] = (uint8_t)((v & 7) | (DEC_DEAD << DEC_MOVE_SHIFT));
                            }
                            nSentL++;
                        } else {
                            o.id[pos] = id;
                            o.birth[pos] = birth;
                            o.lastBirth[pos] = lastBirth;
                            o.flags[pos] = (uint8_t)(v & (F_MALE | F_FERTILE));
                            if (storeAge) o.age[pos] = age;
                            if constexpr (GEN) {
                                o.gslot[pos] = a.gslot[w0 + x];
                                o.nbabies[pos] = a.nbabies[w0 + x] + ((v & F_BORN) ? 1 : 0);  // populations/OoANavGenPop.cpp:243
                            }
                        }
                    }
                    if (who) send_leavers<GEN>(H, who, leaves, m, srcRow);
                    __syncwarp();  // the next round reads the places taken in this one
                }
                nmv = 0;
                __syncwarp();
            };
            while (ci < cEnd - cBase) {
                const int lo = max(s, w0), hi = min(e, w1);
                for (int j0 = lo; j0 < hi; j0 += 32) {
                    const int j = j0 + lane, x = j - w0;
                    const bool valid = j < hi;
                    const uint8_t v = valid ? W.dec[x] : (uint8_t)(DEC_DEAD << DEC_MOVE_SHIFT);
                    const int code = v >> DEC_MOVE_SHIFT;
                    const bool alive = code != DEC_DEAD, born = (v & F_BORN) != 0;
                    const bool stays = alive && code == 0, mover = alive && code != 0;
                    const unsigned ms = __ballot_sync(FULL, stays), mb = __ballot_sync(FULL, born), mm = __ballot_sync(FULL, mover);
                    if (mover) { const int qe = nmv + __popc(mm & lt); S.mvJ[qe] = (uint16_t)x; S.mvC[qe] = (uint8_t)ci; }
                    nmv += __popc(mm);
                    if (stays) {
                        const int pos = ns + stayBase + __popc(ms & lt);
                        o.id[pos] = W.id[x];
                        o.birth[pos] = W.birth[x];
                        o.lastBirth[pos] = W.lastBirth[x];
                        o.flags[pos] = (uint8_t)(v & (F_MALE | F_FERTILE));
                        if (storeAge) o.age[pos] = a.age[j];
                        if constexpr (GEN) {
                            o.gslot[pos] = a.gslot[j];
                            o.nbabies[pos] = a.nbabies[j] + (born ? 1 : 0);
                        }
                    }
                    if (born) { const int me = nMothers + __popc(mb & lt); S.motherId[me] = W.id[x]; S.motherC[me] = (uint8_t)ci; }
                    if constexpr (GEN) { if (born) S.motherIdx[nMothers + __popc(mb & lt)] = j; }
                    stayBase += __popc(ms);
                    nMothers += __popc(mb);
                    if (nmv > WSS::MVC - 32) flush_movers();  // the queue could overflow in the next round
                }
                if (e > w1) break;  // the cell goes on in the next window
                // ---- the cell is complete ----
                if (lane == ci) { mFirstL = cellM0; mCntL = nMothers - cellM0; babyBaseL = ns + stayBase + arL; }
                if (nMothers >= BIRTH_FLUSH) place_babies();
                do { ci++; s = e; e = __shfl_sync(FULL, csL, min(ci + 1, 31)); } while (ci < cEnd - cBase && e == s);
                if (ci < cEnd - cBase) begin_cell();
            }
            if (nmv > 0) flush_movers();  // the window is about to be recycled
            __syncwarp();  // every lane is done with the window: it can be overwritten
            wCons++;
            if (curIssued < nWin) { issueWin(g0, ge, curIssued); curIssued++; }
        }
        if (nMothers > 0) place_babies();
        have = fetch(cBase, cEnd, csL, nsL, arL, bbL);
        curIssued = 0;
    }
    if (H.on) {
        nSentL = __reduce_add_sync(FULL, nSentL);
        if (lane == 0 && nSentL) atomicAdd(&st->nSent, nSentL);
    }
    if (stepEnd) step_end();
}

}  // namespace qhg
