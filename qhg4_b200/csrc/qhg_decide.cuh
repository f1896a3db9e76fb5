// qhg_decide.cuh -- pass 1 of the fast path, second generation: one warp per BATCH of consecutive cells.
//
// k_cell_decide (qhg_cells.cuh) gives every cell its own warp-level prologue, queue flushes, pairing and commit: about 450
// warp instructions of fixed cost per cell and idle lanes in the last 32-agent chunk of every cell -- 13 % of the pass at
// 150 agents per cell and more than half of it at 20 (profiles/README.md, round 1).  Here a warp takes up to SB consecutive
// cells whose agents are ONE contiguous segment of every array (agents are binned by cell) and walks the segment in chunks
// of 32 agents that may straddle cell boundaries:
//   * every lane finds the cell of its agent by comparing its position with the (warp-uniform) cell starts of the batch;
//     per-cell quantities (the LinearBirth / LinearDeath thresholds, the weight row, the neighbours) sit in shared memory,
//     indexed by the cell's number inside the batch;
//   * the fertile census is a pair of counters per cell (segmented popcounts, one writer per cell and chunk); the list of
//     the fertile females and their pairing keys is only built for the cells that need it -- more fertile females than males,
//     the only case in which "does this birth candidate have a mate" is not simply yes;
//   * the queues of the rare expensive work (the double-precision atan of ATanDeath, the neighbour choice of the movers)
//     are flushed when they run full and once at the end of the BATCH, not once per cell;
//   * the commit splits the lanes among the cells of the batch; tallies of the whole warp are kept in registers and reduced
//     once per kernel.
// The per-agent law is exactly that of k_cell_decide (same draws, same thresholds, same decision byte), so the two are
// interchangeable bit for bit (QHG_DECIDE=cell selects the old kernel; tests/test_parity_gpu.py runs both).
// GEN: populations with Genetics -- every birth needs the identity of the father, so in every cell with a birth candidate both
// sexes are listed, keyed and ranked (father[] receives the mate's position).  NAV: programs that end with Navigate -- the
// agents of port and bridge cells get a second look after the move queue is flushed (far jumps go through the jump list).
#pragma once
#include "qhg_cells.cuh"

namespace qhg {

#ifndef QHG_SMALL_CELL_PAIRING
#define QHG_SMALL_CELL_PAIRING 1  // cells of at most 32 agents are paired with one lane per agent (no lists, no scans)
#endif
constexpr int SB_MAX = 8;      // most cells a warp takes per grab of the work counter (template parameter SB: 4 or 8)
constexpr int SEGCAP = WCAP;   // most agents of one sub-batch (bytes of the provisional decisions in shared memory)
constexpr int MAXF_S = 384;    // most fertile females of one cell that can be ranked here (larger cells: generic path)
constexpr int MAXF_SG = 160;   // the same per sex for populations with Genetics

// BIG = the recovery variant of the pass: cells of up to 8192 agents, 4096 ranked fertile females (2048 per sex with Genetics) and
// 2048 births per cell, one CTA per SM with its per-warp slices in dynamic shared memory.  A sharded run redoes a step with it
// when any rank met a cell the default limits do not hold (a single GPU redoes such a step on the generic path).
template <bool BIG>
struct SegLim {
    static constexpr int CAP = BIG ? 8192 : SEGCAP;
    static constexpr int MF = BIG ? 4096 : MAXF_S;
    static constexpr int MFG = BIG ? 2048 : MAXF_SG;
    static constexpr int MM = BIG ? 2048 : MAXMOTHERS;
};

template <int SB, bool BIG = false>
struct SegSmem {
    static constexpr int CAP = SegLim<BIG>::CAP, MF = SegLim<BIG>::MF, MFG = SegLim<BIG>::MFG;
    double row[SB][8];                 // cumulated weight rows (7 used)
    unsigned long long tb[SB + 1], td[SB + 1]; // LinearBirth / LinearDeath thresholds of the cells (k_cell_init)
    int nbr[SB][8];                    // neighbours (6 used)
    int out[SB][8];                    // movers per (cell, direction)
    int cs[SB + 8];                    // starts of the cells inside the segment (cs[nc] = its length)
    int nreal[SB];
    uint32_t mask[3][CAP / 32];        // per chunk of 32 agents: fertile females / fertile males at step start, birth candidates
    union {
        struct {                       // work queues (empty whenever the pairing needs the space below)
            long long qmId[QCAP];
            float qaAge[QCAP];
            uint32_t qaU[QCAP];
            uint16_t qaJ[QCAP], qmJ[QCAP];
        } q;
        struct {                       // pairing of ONE cell with more fertile females than males
            alignas(16) uint32_t keys[MF];
            uint16_t ffJ[MF], candQ[MF];
        } p;
        struct {                       // populations with Genetics: the males are listed, keyed and ranked too
            alignas(16) uint32_t keys[MFG];
            alignas(16) uint32_t mkeys[MFG];
            uint16_t ffJ[MFG], mmJ[MFG], candQ[MFG], candR[MFG], maleOfRank[MFG];
        } g;
    } u;
    alignas(4) uint8_t dec[CAP + 4];     // provisional decisions, shifted by (segment start & 3)
};

// SB = cells per grab: the slot reservations of a sub-batch stay in registers until the next one ends, SB / 4 pairs of them --
// 8 cells per grab pay at 20 agents per cell, 4 at 150 (register pressure: the kernel is capped at 64 registers)
extern __shared__ __align__(16) unsigned char qhg_dyn_smem[];

template <bool SPEC, int SB, bool GEN = false, bool NAV = false, bool BIG = false>
__global__ void __launch_bounds__(DCW * 32, BIG ? 1 : QHG_DECIDE_MINB)
k_seg_decide(DevStats *__restrict__ st, AgentArrays a, ActParams P, CellEnv E, int cLo, int cHi, const int *__restrict__ cellStart,
             int doPair, int *__restrict__ stay, int *__restrict__ arrive, int *__restrict__ birthCount, uint8_t *__restrict__ dec,
             int *__restrict__ moveBase, int shrink, int *__restrict__ father = nullptr, JumpEntry *__restrict__ jumps = nullptr,
             int *__restrict__ jumpCount = nullptr, int jumpCap = 0) {
    static_assert(!(SPEC && (GEN || NAV)), "the compile-time program has neither Genetics nor Navigate");
    static_assert(SB + 1 <= 32 && SB * 8 <= 4 * 32, "one lane per cell start; at most four rounds of (cell, direction) lanes");
    using L = SegLim<BIG>;
    SegSmem<SB, BIG> *smem;
    if constexpr (BIG) {
        smem = reinterpret_cast<SegSmem<SB, true> *>(qhg_dyn_smem);
    } else {
        __shared__ SegSmem<SB, false> smemStatic[DCW];
        smem = smemStatic;
    }
    const int lane = threadIdx.x & 31, wid = (DCW == 1) ? 0 : (int)(threadIdx.x >> 5);
#if QHG_OPAQUE_SMEM
    // the address of the warp's slice as a value the compiler cannot recompute (it would otherwise rebuild it from the CTA's
    // shared-memory window and the thread index at every use rather than keep it in a register)
    SegSmem<SB, BIG> *Sp = &smem[wid];
    asm volatile("" : "+l"(Sp));
    __builtin_assume(__isShared(Sp));
    SegSmem<SB, BIG> &S = *Sp;
#else
    SegSmem<SB, BIG> &S = smem[wid];
#endif
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = lanemask_lt();
    if (st->halt) return;  // an earlier queued step failed (qhgb_run): nothing happens until the host has dealt with it
    const unsigned step = st->step;
    const unsigned long long prog = SPEC ? PROG_TUT5 : P.prog;
    const int nOps = SPEC ? 5 : P.nOps;
    const ProgramInfo I = program_info(prog, nOps);
    const float tNow = P.t, fertMin = P.fertMinAge, fertMax = P.fertMaxAge, fertInter = P.fertInterbirth;
    const float atanAgeLo = P.atanAgeLo, atanAgeHi = P.atanAgeHi;
    const unsigned long long tMove = P.tMove;  // "draw < WeightedMove_prob" as an integer threshold on the 32-bit draw (host)
    const RngKey key = P.key;
    const RoundKeys &RK = P.rk;
    const bool storeAge = !SPEC && P.storeAge != 0;                        // the tutorial order refreshes the age first: never stored
    const bool selfMate = !SPEC && P.selfMate != 0;                        // tut_ParthenoPop: every female counts as mated
    const bool confine = !SPEC && P.confine != 0 && E.allowed != nullptr;  // ConfinedMove filters the chosen destinations
    const bool pairing = selfMate || doPair != 0;                          // without either nobody has a mate: no births
    int nDeadL = 0, nMoveL = 0, nBornL = 0;   // lane-local tallies, reduced once at the end of the kernel
    constexpr int PEND = (SB * 8 + 31) / 32;
    int pendIdx[PEND], pendVal[PEND];  // slot reservations whose atomics are still in flight
#pragma unroll
    for (int r = 0; r < PEND; r++) { pendIdx[r] = -1; pendVal[r] = 0; }

    // cells are handed out dynamically, SB at a time (sea cells are empty, land cells are not: a static split leaves a tail).
    // A range with few grabs per warp (a shard of an 8-GPU run: about four) would end with a tail of up to one grab -- 20 % of
    // the pass there -- so the host asks for grabs that shrink towards the end of the range (`shrink`; it costs 2 % where the
    // range is long and is off there)
    const int nWarps = gridDim.x * DCW;
    int lastEnd = cLo;  // where this warp's last grab ended: the work counter is at least there
    for (;;) {
    const int g = shrink ? max(1, min(SB, (cHi - lastEnd) / (2 * nWarps))) : SB;
    int cBase = 0;
    if (lane == 0) cBase = cLo + atomicAdd(&st->workDecide, g);
    cBase = __shfl_sync(FULL, cBase, 0);
    if (cBase >= cHi) break;
    lastEnd = cBase + g;
    const int nB = min(g, cHi - cBase);
    const int csL = (lane <= nB) ? cellStart[cBase + lane] : 0;
    int g0 = 0;
    while (g0 < nB) {
        // the sub-batch: as many of the remaining cells as fit the shared-memory segment
        const int s = __shfl_sync(FULL, csL, g0);
        const unsigned fit = __ballot_sync(FULL, lane > g0 && lane <= nB && csL - s <= L::CAP);
        if (!fit) {  // one cell larger than the segment: the host reruns the step on the generic path
            if (lane == 0) atomicExch(&st->oversize, 1);
            g0++;
            continue;
        }
        const int g1 = 31 - __clz(fit);
        const int nc = g1 - g0;
        const int n = __shfl_sync(FULL, csL, g1) - s;
        const int c0 = cBase + g0;
        if (n == 0) { g0 = g1; continue; }  // sea

        // ---- per-cell data of the sub-batch -----------------------------------------------------------------------------
        // the cell (inside the sub-batch) of position j of the segment: number of cell starts at or below j
        auto cell_of = [&](int j) {
            int ci = 0;
            for (int k = 1; k < nc; k++) ci += (j >= S.cs[k]) ? 1 : 0;
            return ci;
        };
        const int gOff = s & 3;  // shared-memory word k of dec[] then is an aligned global word
        uint8_t *const sdec = S.dec + gOff;
        {
            const int v = __shfl_sync(FULL, csL, min(g0 + lane, 31));
            if (lane <= nc) S.cs[lane] = v - s;
        }
        if (lane < nc) {
            S.tb[lane] = I.hasVerhulst ? E.TB[c0 + lane] : 0ull;
            S.td[lane] = I.hasVerhulst ? E.TD[c0 + lane] : 0ull;
            S.nreal[lane] = E.nNbr[c0 + lane];
        }
        if (lane > nc && lane < SB + 8) S.cs[lane] = 0x7fffffff;  // sentinels: the walk below never runs past the last cell
        for (int q = lane; q < nc * 8; q += 32) {
            const int ci = q >> 3, k = q & 7;
            S.row[ci][k] = (k < WSTRIDE) ? E.W[(size_t)(c0 + ci) * WSTRIDE + k] : 0.0;
            S.nbr[ci][k] = (k < MAXN) ? E.nbr[(size_t)(c0 + ci) * MAXN + k] : -1;
            S.out[ci][k] = 0;
        }
        __syncwarp();

        int nqa = 0, nqm = 0;
        int confL = 0;  // moves of this lane that ConfinedMove turned back (still counted: core/SPopulation.cpp:1067)
        auto flush_atan = [&]() {  // ATanDeath::execute, actions/ATanDeath.cpp:75-83, for the queued agents
            for (int e = lane; e < nqa; e += 32) {
                const double x = __dmul_rn(P.atanSlope, __dadd_rn((double)S.u.q.qaAge[e], -P.atanMaxAge));
                const double p = __dadd_rn(0.5, __ddiv_rn(__dmul_rn(P.atanScale, atan_rn(x)), 3.141592653589793));
                if (u2d(S.u.q.qaU[e]) < p) sdec[S.u.q.qaJ[e]] |= T_ATANDIES;
            }
            nqa = 0;
            __syncwarp();
        };
        auto flush_move = [&]() {  // WeightedMove::execute, actions/WeightedMove.cpp:56-98, for the queued agents
            for (int e = lane; e < nqm; e += 32) {
                const int j = S.u.q.qmJ[e];
                const int ci = cell_of(j);
                const uint32_t u = agent_draws_rk(S.u.q.qmId[e], step, STREAM_ACT1, RK).x;
                const int nreal = S.nreal[ci];
                const double *row = S.row[ci];
                int pick = -1;
                if (!SPEC && I.randomMove) {  // RandomMove: uniform over "stay" and the neighbours, no ice test
                    pick = (int)__dmul_rn(u2d(u), (double)(nreal + 1));
                    if (pick > 0 && S.nbr[ci][pick - 1] >= 0) {
                        if (confine && !E.allowed[S.nbr[ci][pick - 1]]) { if (!(I.moveAfterAtan && (sdec[j] & T_ATANDIES))) confL++; }
                        else sdec[j] |= (uint8_t)(pick << DEC_MOVE_SHIFT);
                    }
                    continue;
                }
                if (!SPEC && I.condMove) {  // CondWeightedMove (actions/CondWeightedMove.cpp:41-86): the whole row, the ice of the cell
                    // the agent is in, the MoveCondition (a SimpleCondition over the altitudes)
                    const double r2 = __dmul_rn(u2d(u), row[MAXN]);
                    for (int q = 0; q < MAXN + 1; q++) {
                        if (r2 < row[q]) { pick = q; break; }
                    }
                    if (pick > 0) {
                        const int dst = S.nbr[ci][pick - 1];
                        if (dst >= 0 && !(E.ice && E.ice[c0 + ci]) && cond_allow(P.condMode, E.alt[c0 + ci], E.alt[dst])) {
                            if (confine && !E.allowed[dst]) { if (!(I.moveAfterAtan && (sdec[j] & T_ATANDIES))) confL++; }
                            else sdec[j] |= (uint8_t)(pick << DEC_MOVE_SHIFT);
                        }
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
                    const int dst = S.nbr[ci][pick - 1];
                    if (dst >= 0 && !(E.ice && E.ice[dst])) {
                        // ConfinedMove (actions/ConfinedMove.cpp:86-101): registered and counted, but it leads back to the cell it
                        // starts from -- unless ATanDeath (flushed before, see below) removed the agent before it moved
                        if (confine && !E.allowed[dst]) { if (!(I.moveAfterAtan && (sdec[j] & T_ATANDIES))) confL++; }
                        else sdec[j] |= (uint8_t)(pick << DEC_MOVE_SHIFT);
                    }
                }
            }
            nqm = 0;
            __syncwarp();
        };

        // the lanes are split among the cells of the sub-batch (32, 16, 8 or 4 lanes per cell) for the census and the commit
        const int lpcShift = (nc <= 1) ? 5 : (nc <= 2) ? 4 : (nc <= 4) ? 3 : (nc <= 8) ? 2 : 1;
        const int LPC = 1 << lpcShift, myc = lane >> lpcShift, li = lane & (LPC - 1);
        const uint32_t ONES = 0x01010101u;
        // ---- one pass over the segment: all actions, provisional decisions ----------------------------------------------
        // (arrays are indexed by the global position s + j: base pointers are kernel parameters, one multiply-add per address)
        int64_t idN = 0; float birthN = 0, lastN = 0, ageN = 0; uint8_t fN = 0;
        if (lane < n) {
            const int g = s + lane;
            idN = a.id[g]; birthN = a.birth[g]; fN = a.flags[g];
            if (I.hasFert) lastN = a.lastBirth[g];
            if (storeAge) ageN = a.age[g];
        }
        // every lane walks through the cells as its position advances: the cell number and the cell's thresholds stay in registers
        int ci = -1, nextStart = 0;
        unsigned long long tbw = 0, tDeath = 0;
        const bool wantMasks = !selfMate && doPair && I.hasVerhulst;
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            const bool valid = j < n;
            while (j >= nextStart) {  // at most once per chunk unless empty cells lie in between
                ci++;
                nextStart = S.cs[ci + 1];
                tbw = S.tb[ci];
                tDeath = S.td[ci];
            }
            const int64_t id = idN; const float birth = birthN, lastBirth = lastN; const uint8_t f0 = fN;
            float ag = ageN;
            if (j + 32 < n) {
                const int g2 = s + j + 32;
                idN = a.id[g2]; birthN = a.birth[g2]; fN = a.flags[g2];
                if (I.hasFert) lastN = a.lastBirth[g2];
                if (storeAge) ageN = a.age[g2];
            }
            const uint4 r0 = (SPEC || I.needAct0) ? agent_draws_rk(id, step, STREAM_ACT0, RK) : make_uint4(0, 0, 0, 0);
            const bool fertF = valid && ((f0 & (F_FERTILE | F_MALE)) == F_FERTILE);
            const bool fertM = valid && ((f0 & (F_FERTILE | F_MALE)) == (F_FERTILE | F_MALE));
            bool needAtan = false, needMove = false;
            uint8_t f = f0 & (F_MALE | F_FERTILE);
            bool dead = false, cand = false;
            if constexpr (SPEC) {
                // GetOld, ATanDeath, WeightedMove, Fertility, Verhulst as straight-line code (tutorial_data/xmldat/tut_EnvironAlt.xml)
                ag = __fsub_rn(tNow, birth);                                   // actions/GetOld.cpp:37-48
                const bool above = ag > atanAgeHi;                             // actions/ATanDeath.cpp:66-90: p > 1 above the window,
                needAtan = valid && (ag >= atanAgeLo) && !above;               //   decided exactly at the flush inside it,
                dead = above;                                                  //   nobody dies below it
                needMove = valid && !above && ((unsigned long long)r0.y < tMove);  // actions/WeightedMove.cpp:45-106, neighbour chosen at the flush
                // actions/Fertility.cpp:49-74
                const bool fert = (ag > fertMin) && ((f & F_MALE) || ((ag < fertMax) && (__fsub_rn(tNow, lastBirth) > fertInter)));
                f = (uint8_t)((f & F_MALE) | (fert ? F_FERTILE : 0));
                // actions/Verhulst.cpp:101-115 -> LinearBirth.cpp:122-168, LinearDeath.cpp:131-153
                const bool bPos = (tbw >> 62) & 1ull, bNeg = (tbw >> 63) != 0;
                const bool below = (unsigned long long)r0.z < (tbw & 0x1ffffffffull);
                cand = !above && bPos && pairing && fertF && below;            // a birth needs a mate: settled per cell after the pass
                dead = dead || (bNeg && below) || ((unsigned long long)r0.w < tDeath);
            } else {
                bool alive = true;
#pragma unroll 1
                for (int k = 0; k < nOps; k++) {
                    if (!alive) break;
                    const int op = (int)((prog >> (4 * k)) & 15ull);
                    if (op == OP_GETOLD) {  // actions/GetOld.cpp:37-48
                        ag = __fsub_rn(tNow, birth);
                    } else if (op == OP_ATANDEATH) {  // actions/ATanDeath.cpp:66-90
                        ag = __fsub_rn(tNow, birth);
                        if (ag >= atanAgeLo) {
                            if (ag <= atanAgeHi) needAtan = true;
                            else alive = false;
                        }
                    } else if (op == OP_OLDAGEDEATH) {  // actions/OldAgeDeath.cpp:48-67
                        ag = __fsub_rn(tNow, birth);
                        const uint32_t uo = agent_draws(id, step, STREAM_ACT1, key).w;
                        if ((double)ag > __dadd_rn(P.oadMaxAge, u2range(uo, P.oadLo, P.oadHi))) alive = false;
                    } else if (op == OP_WEIGHTEDMOVE || op == OP_RANDOMMOVE || op == OP_CONDWEIGHTEDMOVE) {  // actions/WeightedMove.cpp:45-106, RandomMove.cpp:65-100, CondWeightedMove.cpp:41-86
                        if ((unsigned long long)r0.y < tMove) needMove = true;
                    } else if (op == OP_FERTILITY) {  // actions/Fertility.cpp:49-74
                        bool fert;
                        if (!(f & F_MALE)) fert = (ag > fertMin) && (ag < fertMax) && (__fsub_rn(tNow, lastBirth) > fertInter);
                        else fert = ag > fertMin;
                        f = (uint8_t)((f & F_MALE) | (fert ? F_FERTILE : 0));
                    } else if (op == OP_VERHULST) {  // actions/Verhulst.cpp:101-115 -> LinearBirth.cpp:122-168, LinearDeath.cpp:131-153
                        const bool bPos = (tbw >> 62) & 1ull, bNeg = (tbw >> 63) != 0;
                        const bool below = (unsigned long long)r0.z < (tbw & 0x1ffffffffull);
                        if (bPos) {
                            const bool mayBear = selfMate ? !(f0 & F_MALE) : fertF;
                            if (mayBear && pairing && below) cand = true;
                        } else if (bNeg) {
                            if (below) alive = false;
                        }
                        if (alive && (unsigned long long)r0.w < tDeath) alive = false;
                    } else if (op == OP_DROWN) {  // populations/tut_EnvironAltPop.cpp:100-116 (EVENT_ID_GEO)
                        // (lanes past the end of the segment have walked to cell nc: they must not read a cell past the grid)
                        if (valid && (E.alt[c0 + ci] < 0 || (E.ice && E.ice[c0 + ci]))) alive = false;
                    }
                }
                needAtan = needAtan && valid;
                needMove = needMove && valid;
                dead = !alive;
                if (storeAge && valid) a.age[s + j] = ag;
            }
            if (valid) sdec[j] = (uint8_t)(f | (cand ? F_BORN : 0) | (dead ? T_DEADNOW : 0));
            if (wantMasks) {  // the pairing's census: one bit per agent and kind, one word per chunk
                const unsigned bF = __ballot_sync(FULL, fertF), bM = __ballot_sync(FULL, fertM), bC = __ballot_sync(FULL, cand && valid);
                if (lane < 3) S.mask[lane][j0 >> 5] = (lane == 0) ? bF : (lane == 1) ? bM : bC;
            }
            // queue the rare expensive work
            const unsigned ma = __ballot_sync(FULL, needAtan), mm = __ballot_sync(FULL, needMove);
            if (needAtan) { const int e = nqa + __popc(ma & lt); S.u.q.qaAge[e] = ag; S.u.q.qaU[e] = r0.x; S.u.q.qaJ[e] = (uint16_t)j; }
            if (needMove) { const int e = nqm + __popc(mm & lt); S.u.q.qmJ[e] = (uint16_t)j; S.u.q.qmId[e] = id; }
            nqa += __popc(ma);
            nqm += __popc(mm);
            __syncwarp();
            // flushed when a queue could overflow in the next round, and at the end of the SEGMENT (not of every cell)
            const bool last = j0 + 32 >= n;
            // with ConfinedMove the move flush reads the ATanDeath verdicts of its agents: the death queue goes first
            if (nqa > QCAP - 32 || (last && nqa > 0) || (confine && nqa > 0 && (nqm > QCAP - 32 || (last && nqm > 0)))) flush_atan();
            if (nqm > QCAP - 32 || (last && nqm > 0)) flush_move();
        }

        // ---- Navigate (actions/Navigate.cpp:181-250), the last action of the program: the agents of port and bridge cells --------
        if constexpr (NAV) {
            const bool seesMoving = nav_sees_moving(prog, nOps);
            for (int ci = 0; ci < nc; ci++) {  // warp-uniform
                const int c = c0 + ci, b0 = S.cs[ci], b1 = S.cs[ci + 1];
                if (b1 == b0) continue;
                const int port = E.navRow ? E.navRow[c] : -1;
                bool hasBridge = false;
                for (int b = 0; b < E.nBridges; b++) { const int2 br = E.bridges[b]; hasBridge |= (br.x == c) || (br.y == c); }
                if (port < 0 && !hasBridge) continue;
                int p0 = 0, nd = 0;
                if (port >= 0) { p0 = E.navPtr[port]; nd = E.navPtr[port + 1] - p0 - 1; }
                const int lim = (c < nd) ? c : nd;  // the reference bounds the search by the port's cell index (:194)
                for (int j = b0 + lane; j < b1; j += 32) {
                    const uint8_t v0 = sdec[j];
                    if (v0 & (T_ATANDIES | T_DEADNOW)) continue;          // dead before the last action
                    const int code0 = (v0 >> DEC_MOVE_SHIFT) & 7;
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
                        if (e < jumpCap) jumps[e] = JumpEntry{s + j, to, slot, (int)(v0 & 7)};
                        else atomicExch(&st->oversize, 1);  // the list is full: the step is redone on the generic path
                        sdec[j] = (uint8_t)((v0 & 7) | (DEC_DEAD << DEC_MOVE_SHIFT));
                        confL += (code0 != 0 ? 1 : 0) + navMoves - 1;  // the commit counts one move for a leaving agent
                    }
                }
                __syncwarp();
            }
        }

        // ---- pairing: RandomPair::findMates (actions/RandomPair.cpp:146-279) under the counter-mode law ------------------
        // fertile females and males are ranked by (random key, id), equal ranks mate.  With nF <= nM every fertile female has
        // a mate; otherwise the nM females with the smallest keys -- and only the birth candidates need to know.  Populations
        // with Genetics need the father of every birth: there both sexes are ranked in every cell that has a candidate.
        if (wantMasks) {
            __syncwarp();
            for (int ci = 0; ci < nc; ci++) {  // warp-uniform
                const int b0 = S.cs[ci], b1 = S.cs[ci + 1];
                if (b1 == b0) continue;
#if QHG_SMALL_CELL_PAIRING
                if (SB >= 8 && b1 - b0 <= 32) {  // (only in the kernels for sparse populations: at 150 agents per cell the extra code costs 3 %)
                    // a cell of at most 32 agents: one lane per agent, no lists.  The cell's bits of the three chunk masks are cut
                    // out by every lane alike (two broadcast loads and a funnel shift each: no reduction to learn the counts), both
                    // sexes get their keys from ONE Philox execution, a lane's rank is a walk over the set bits of its sex's mask
                    const int len = b1 - b0, w = b0 >> 5, sh = b0 & 31;
                    const unsigned cm = (len == 32) ? FULL : ((1u << len) - 1u);
                    const bool two = sh + len > 32;
                    const unsigned Fm = __funnelshift_r(S.mask[0][w], two ? S.mask[0][w + 1] : 0u, sh) & cm;
                    const unsigned Mm = __funnelshift_r(S.mask[1][w], two ? S.mask[1][w + 1] : 0u, sh) & cm;
                    const int nF = __popc(Fm), nMc = __popc(Mm);
                    if (!GEN && nF <= nMc) continue;  // every fertile female has a mate
                    const unsigned Cm = __funnelshift_r(S.mask[2][w], two ? S.mask[2][w + 1] : 0u, sh) & cm;
                    if (Cm == 0) continue;            // no birth candidate in the cell: nothing to settle
                    const int j = b0 + lane;
                    const bool isF = (Fm >> lane) & 1u, isM = GEN && ((Mm >> lane) & 1u), isC = (Cm >> lane) & 1u;
                    uint32_t *const keys = GEN ? S.u.g.keys : S.u.p.keys;
                    uint32_t key = 0;
                    if (isF || isM) key = agent_draws_rk(a.id[s + j], step, STREAM_PAIR, RK).x;
                    keys[lane] = key;
                    __syncwarp();
                    int r = 0;
                    bool tie = false;
                    const unsigned mine = isF ? Fm : (isM ? Mm : 0u);
                    for (unsigned m = mine; m; m &= m - 1u) {
                        const int t = __ffs(m) - 1;
                        const uint32_t kt = keys[t];
                        r += (kt < key) ? 1 : 0;
                        tie = tie || (kt == key && t != lane);
                    }
                    if (tie) {  // equal keys (about one pair in 10^8): the id decides
                        const int64_t myId = a.id[s + j];
                        for (unsigned m = mine; m; m &= m - 1u) {
                            const int t = __ffs(m) - 1;
                            if (t != lane && keys[t] == key && a.id[s + b0 + t] < myId) r++;
                        }
                    }
                    const int np = min(nF, nMc);  // couples
                    if (isC && r >= np) sdec[j] &= (uint8_t)~F_BORN;  // no mate: no birth
                    if constexpr (GEN) {
                        if (isM && r < np) S.u.g.maleOfRank[r] = (uint16_t)j;
                        __syncwarp();
                        if (isC && r < np) father[s + j] = s + S.u.g.maleOfRank[r];  // the mate of rank r is the father
                    }
                    __syncwarp();
                    continue;
                }
#endif
                // lane k looks at chunk k (of every group of 32 chunks: one group unless BIG) of the segment, restricted to the
                // positions [b0, b1) of this cell
                constexpr int NG = L::CAP / 1024;
                unsigned wF[NG], wM[NG], wC[NG], rmG[NG];
                int cF = 0, cM = 0;
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const int ch = 32 * g + lane;
                    const int lo = b0 - 32 * ch, hi = b1 - 32 * ch;
                    unsigned rm = 0;
                    if (hi > 0 && lo < 32) rm = ((hi >= 32) ? FULL : ((1u << hi) - 1u)) & ((lo <= 0) ? FULL : ~((1u << lo) - 1u));
                    rmG[g] = rm;
                    wF[g] = rm ? (S.mask[0][ch] & rm) : 0u;
                    wM[g] = rm ? (S.mask[1][ch] & rm) : 0u;
                    cF += __popc(wF[g]); cM += __popc(wM[g]);
                }
                const int nF = __reduce_add_sync(FULL, cF), nMc = __reduce_add_sync(FULL, cM);
                if (!GEN && nF <= nMc) continue;  // every fertile female has a mate
                int cC = 0;
#pragma unroll
                for (int g = 0; g < NG; g++) { wC[g] = rmG[g] ? (S.mask[2][32 * g + lane] & rmG[g]) : 0u; cC += __popc(wC[g]); }
                const int nCand = __reduce_add_sync(FULL, cC);
                if (nCand == 0) continue;  // no birth candidate in the cell: nothing to settle
                if (nF > (GEN ? L::MFG : L::MF) || (GEN && nMc > L::MFG)) {
                    if (lane == 0) atomicExch(&st->oversize, 1);
                    continue;
                }
                // the cell's fertile females in position order, and the candidates as indices into that list
                uint16_t *const ffJ = GEN ? S.u.g.ffJ : S.u.p.ffJ, *const candQ = GEN ? S.u.g.candQ : S.u.p.candQ;
                uint32_t *const keys = GEN ? S.u.g.keys : S.u.p.keys;
                int offF = 0, offC = 0, offM = 0;  // entries of the groups before this one
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const int gF = __popc(wF[g]), gC = __popc(wC[g]), gM = __popc(wM[g]);
                    int inF = gF, inC = gC, inM = gM;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int xF = __shfl_up_sync(FULL, inF, o), xC = __shfl_up_sync(FULL, inC, o);
                        if (lane >= o) { inF += xF; inC += xC; }
                        if constexpr (GEN) { const int xM = __shfl_up_sync(FULL, inM, o); if (lane >= o) inM += xM; }
                    }
                    const int baseF = offF + inF - gF;
                    unsigned w = wF[g];
                    int idx = baseF;
                    while (w) { const int t = __ffs(w) - 1; w &= w - 1; ffJ[idx++] = (uint16_t)(32 * (32 * g + lane) + t); }
                    w = wC[g];
                    idx = offC + inC - gC;
                    while (w) { const int t = __ffs(w) - 1; w &= w - 1; candQ[idx++] = (uint16_t)(baseF + __popc(wF[g] & ((1u << t) - 1u))); }
                    if constexpr (GEN) {
                        w = wM[g];
                        idx = offM + inM - gM;
                        while (w) { const int t = __ffs(w) - 1; w &= w - 1; S.u.g.mmJ[idx++] = (uint16_t)(32 * (32 * g + lane) + t); }
                    }
                    if constexpr (NG > 1) {
                        offF += __shfl_sync(FULL, inF, 31); offC += __shfl_sync(FULL, inC, 31);
                        if constexpr (GEN) offM += __shfl_sync(FULL, inM, 31);
                    }
                }
                __syncwarp();
                for (int q = lane; q < nF; q += 32) keys[q] = agent_draws_rk(a.id[s + ffJ[q]], step, STREAM_PAIR, RK).x;
                if constexpr (GEN) {
                    for (int m = lane; m < nMc; m += 32) S.u.g.mkeys[m] = agent_draws_rk(a.id[s + S.u.g.mmJ[m]], step, STREAM_PAIR, RK).x;
                }
                __syncwarp();
                const int np = min(nF, nMc);  // couples
                for (int i = lane; i < nCand; i += 32) {
                    const int q = candQ[i];
                    const uint32_t k = keys[q];
                    // rank = number of smaller keys; four keys per shared-memory load; "<=" counts reveal ties (the key itself is one)
                    int r = 0, le = 0;
                    const int nF4 = nF & ~3;
                    for (int e = 0; e < nF4; e += 4) {
                        const uint4 kk = *reinterpret_cast<const uint4 *>(&keys[e]);
                        r += (kk.x < k) + (kk.y < k) + (kk.z < k) + (kk.w < k);
                        le += (kk.x <= k) + (kk.y <= k) + (kk.z <= k) + (kk.w <= k);
                    }
                    for (int e = nF4; e < nF; e++) {
                        const uint32_t ke = keys[e];
                        r += (ke < k) ? 1 : 0;
                        le += (ke <= k) ? 1 : 0;
                    }
                    if ((le - r) > 1) {  // equal keys (about one pair in 10^8): the id decides
                        const int64_t myId = a.id[s + ffJ[q]];
                        for (int e = 0; e < nF; e++) {
                            if (e != q && keys[e] == k && a.id[s + ffJ[e]] < myId) r++;
                        }
                    }
                    if (r >= np) sdec[ffJ[q]] &= (uint8_t)~F_BORN;  // no mate: no birth
                    if constexpr (GEN) S.u.g.candR[i] = (uint16_t)r;
                }
                if constexpr (GEN) {
                    for (int m = lane; m < nMc; m += 32) {  // rank of every fertile male; the first np of them are mates
                        const uint32_t k = S.u.g.mkeys[m];
                        int r = 0, le = 0;
                        const int nM4 = nMc & ~3;
                        for (int e = 0; e < nM4; e += 4) {
                            const uint4 kk = *reinterpret_cast<const uint4 *>(&S.u.g.mkeys[e]);
                            r += (kk.x < k) + (kk.y < k) + (kk.z < k) + (kk.w < k);
                            le += (kk.x <= k) + (kk.y <= k) + (kk.z <= k) + (kk.w <= k);
                        }
                        for (int e = nM4; e < nMc; e++) {
                            const uint32_t ke = S.u.g.mkeys[e];
                            r += (ke < k) ? 1 : 0;
                            le += (ke <= k) ? 1 : 0;
                        }
                        if ((le - r) > 1) {
                            const int64_t myId = a.id[s + S.u.g.mmJ[m]];
                            for (int e = 0; e < nMc; e++) if (e != m && S.u.g.mkeys[e] == k && a.id[s + S.u.g.mmJ[e]] < myId) r++;
                        }
                        if (r < np) S.u.g.maleOfRank[r] = S.u.g.mmJ[m];
                    }
                    __syncwarp();
                    for (int i = lane; i < nCand; i += 32) {  // the mate of rank r is the father
                        const int r = S.u.g.candR[i];
                        if (r < np) father[s + ffJ[candQ[i]]] = s + S.u.g.maleOfRank[r];
                    }
                }
                __syncwarp();
            }
        }

        // ---- commit: final decision bytes, per-cell counts ---------------------------------------------------------------
        // four agents (one 32-bit word of decision bytes) per lane and round, all byte lanes in parallel
        {
            const uint32_t bornVoid = I.bornAfterAtan ? ONES : 0u, moveVoid = I.moveAfterAtan ? ONES : 0u;
            int stayL = 0, bornL = 0, moveL = 0, outL = 0, cellN = 0;
            if (myc < nc) {
                const int b0 = gOff + S.cs[myc], b1 = gOff + S.cs[myc + 1];  // byte range of my cell in S.dec
                cellN = b1 - b0;
                const uint32_t *sw = reinterpret_cast<const uint32_t *>(S.dec);
                uint32_t *gw32 = reinterpret_cast<uint32_t *>(dec + (s - gOff));
                for (int k = (b0 >> 2) + li; 4 * k < b1; k += LPC) {
                    const uint32_t w = sw[k];
                    uint32_t vm = ONES;  // the bytes of this word that belong to my cell
                    if (4 * k < b0 || 4 * k + 4 > b1) {
                        vm = 0;
#pragma unroll
                        for (int b = 0; b < 4; b++) if (4 * k + b >= b0 && 4 * k + b < b1) vm |= 1u << (8 * b);
                    }
                    const uint32_t at = (w >> 6) & ONES, dn = (w >> 7) & ONES, dd = at | dn;
                    const uint32_t code = (w >> DEC_MOVE_SHIFT) & 0x07070707u;
                    const uint32_t nz = ((code + 0x07070707u) >> 3) & ONES;       // move code != 0
                    const uint32_t born = (w >> 2) & ~(at & bornVoid) & ONES;
                    const uint32_t alive = ~dd & vm;
                    const uint32_t out = alive & nz;
                    const uint32_t d7 = (dd << 3) - dd;                            // 7 in every dead byte
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
                        for (int b = 0; b < 4; b++) if (out & (1u << (8 * b))) atomicAdd(&S.out[myc][((code >> (8 * b)) & 7) - 1], 1);
                    }
                }
            }
            // sums over the lanes of a cell (xor butterflies stay inside the aligned group)
            for (int o = LPC >> 1; o > 0; o >>= 1) {
                stayL += __shfl_xor_sync(FULL, stayL, o);
                bornL += __shfl_xor_sync(FULL, bornL, o);
                outL += __shfl_xor_sync(FULL, outL, o);
                moveL += __shfl_xor_sync(FULL, moveL, o);
            }
            if (myc < nc && li == 0) {  // a cell belongs to exactly one warp: plain stores
                stay[c0 + myc] = stayL;
                birthCount[c0 + myc] = bornL;
                if (bornL > L::MM) atomicExch(&st->oversize, 1);
                nBornL += bornL;
                nMoveL += moveL;
                nDeadL += cellN - stayL - outL;
            }
            nMoveL += confL;
        }
        __syncwarp();
        // the movers towards neighbour k take the slots [base, base+cnt) of that cell's arrivals: pass 2 places them without
        // atomics.  The returned slot is stored one sub-batch later (the atomic's latency stays off the critical path).
        {
#pragma unroll
            for (int r = 0; r < PEND; r++) {
                if (pendIdx[r] >= 0) moveBase[pendIdx[r]] = pendVal[r];
                pendIdx[r] = -1;
                const int q = lane + 32 * r;
                if (q < nc * 8 && (q & 7) < MAXN) {
                    const int ci = q >> 3, k = q & 7, cnt = S.out[ci][k];
                    pendVal[r] = cnt ? atomicAdd(&arrive[S.nbr[ci][k]], cnt) : 0;
                    pendIdx[r] = (c0 + ci) * MOVE_STRIDE + k;
                }
            }
        }
        __syncwarp();
        g0 = g1;
    }
    }
#pragma unroll
    for (int r = 0; r < PEND; r++) if (pendIdx[r] >= 0) moveBase[pendIdx[r]] = pendVal[r];
    nDeadL = __reduce_add_sync(FULL, nDeadL);
    nMoveL = __reduce_add_sync(FULL, nMoveL);
    nBornL = __reduce_add_sync(FULL, nBornL);
    if (lane == 0) {
        if (nDeadL) atomicAdd(&st->nDeaths, nDeadL);
        if (nMoveL) atomicAdd(&st->nMoves, nMoveL);
        if (nBornL) atomicAdd(&st->nBirths, nBornL);
    }
}

}  // namespace qhg
