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
#include "qhg_kernels.cuh"

namespace qhg {

constexpr int CW = 8;            // warps per CTA
constexpr int WCAP = 1024;       // largest cell (agents) the fast path handles; larger ones -> generic path
constexpr int MAXF = 512;        // most fertile females of one cell that can be ranked in shared memory
constexpr int QCAP = 64;         // work-queue entries per warp
constexpr int MAXMOTHERS = 128;  // most births of one cell per step on the fast path

// decision byte handed from pass 1 to pass 2: bit0 male, bit1 fertile (the agent's new flags), bit2 gave birth,
// bits 3-5 move code: 0 stays, 1..6 neighbour slot + 1, 7 dead
constexpr int DEC_MOVE_SHIFT = 3;
constexpr uint8_t DEC_DEAD = 7;
// transient bits while a cell is being worked on
constexpr uint8_t T_HASMATE = 4, T_ATANDIES = 0x40, T_DEADNOW = 0x80;

struct WarpSmem {
    double qaX[QCAP];          // ATanDeath queue: argument of the atan
    uint32_t qaU[QCAP];        //                  the agent's death draw
    uint32_t keys[MAXF];       // pairing keys of the cell's fertile females
    uint16_t keyJ[MAXF];       //   and their position in the cell
    uint16_t qaJ[QCAP];
    uint16_t qmJ[QCAP];        // WeightedMove queue: position in the cell
    uint8_t dec[WCAP];         // flags -> provisional decision of every agent of the cell
};

struct ProgramInfo {  // warp-uniform facts about the action program
    bool needAct0, hasFert, hasVerhulst;
    bool moveAfterAtan, bornAfterAtan;
};

__device__ __forceinline__ ProgramInfo program_info(const ActParams &P) {
    ProgramInfo I{false, false, false, false, false};
    int ka = -1;
    for (int k = 0; k < P.nOps; k++) {
        int op = prog_op(P, k);
        if (op == OP_ATANDEATH) { ka = k; I.needAct0 = true; }
        if (op == OP_WEIGHTEDMOVE) { I.needAct0 = true; if (ka >= 0) I.moveAfterAtan = true; }
        if (op == OP_VERHULST) { I.needAct0 = true; I.hasVerhulst = true; if (ka >= 0) I.bornAfterAtan = true; }
        if (op == OP_FERTILITY) I.hasFert = true;
    }
    return I;
}

// ---------------------------------------------------------------------------------------------
// pass 1
__global__ void __launch_bounds__(CW * 32)
k_cell_decide(DevStats *__restrict__ st, AgentArrays a, ActParams P, CellEnv E, int nCells, const int *__restrict__ cellStart,
              int doPair, int *__restrict__ stay, int *__restrict__ arrive, int *__restrict__ birthCount, uint8_t *__restrict__ dec) {
    __shared__ WarpSmem smem[CW];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpSmem &S = smem[wid];
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = lanemask_lt();
    const unsigned step = st->step;
    const ProgramInfo I = program_info(P);
    const int gw = blockIdx.x * CW + wid, nW = gridDim.x * CW;
    int nDead = 0, nMove = 0, nBorn = 0;  // warp-uniform tallies

    for (int c = gw; c < nCells; c += nW) {
        const int s = cellStart[c], n = cellStart[c + 1] - s;
        if (n == 0) continue;
        if (n > WCAP) {
            if (lane == 0) atomicExch(&st->oversize, 1);
            continue;
        }
        // ---- fertile counts; flags go to shared memory ----------------------------------------------------------
        int nF = 0, nM = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            uint8_t f = (j < n) ? a.flags[s + j] : 0;
            if (j < n) S.dec[j] = f;
            nF += __popc(__ballot_sync(FULL, (f & (F_FERTILE | F_MALE)) == F_FERTILE));
            nM += __popc(__ballot_sync(FULL, (f & (F_FERTILE | F_MALE)) == (F_FERTILE | F_MALE)));
        }
        // ---- pairing: RandomPair::findMates (actions/RandomPair.cpp:146-279) under the counter-mode law ----------
        // fertile females and fertile males are ranked by (random key, id); equal ranks mate.  Only "does this
        // female have a mate" matters to the actions: with nF <= nM every fertile female has one, otherwise the nM
        // females with the smallest keys.
        bool allPaired = false;
        if (doPair && nF > 0 && nM > 0) {
            if (nF <= nM) {
                allPaired = true;
            } else if (nF > MAXF) {
                if (lane == 0) atomicExch(&st->oversize, 1);
                continue;
            } else {
                int nf = 0;
                for (int j0 = 0; j0 < n; j0 += 32) {
                    const int j = j0 + lane;
                    const bool ff = (j < n) && ((S.dec[j] & (F_FERTILE | F_MALE)) == F_FERTILE);
                    const unsigned m = __ballot_sync(FULL, ff);
                    if (ff) {
                        const int pos = nf + __popc(m & lt);
                        S.keys[pos] = agent_draws(a.id[s + j], step, STREAM_PAIR, P.key).x;
                        S.keyJ[pos] = (uint16_t)j;
                    }
                    nf += __popc(m);
                }
                __syncwarp();
                for (int q0 = 0; q0 < nF; q0 += 32) {
                    const int q = q0 + lane;
                    if (q < nF) {
                        const uint32_t k = S.keys[q];
                        int r = 0;
                        for (int e = 0; e < nF; e++) {
                            const uint32_t ke = S.keys[e];
                            if (ke < k) r++;
                            else if (ke == k && e != q && a.id[s + S.keyJ[e]] < a.id[s + S.keyJ[q]]) r++;
                        }
                        if (r < nM) S.dec[S.keyJ[q]] |= T_HASMATE;
                    }
                }
            }
        }
        __syncwarp();

        // ---- actions, provisional decisions ------------------------------------------------------------------------
        int nqa = 0, nqm = 0;
        const int nreal = E.nNbr[c];
        const double *row = E.W + (size_t)c * WSTRIDE;
        const double bC = I.hasVerhulst ? E.B[c] : 0.0, dC = I.hasVerhulst ? E.D[c] : 0.0;
        auto flush_atan = [&]() {  // ATanDeath::execute, actions/ATanDeath.cpp:75-83, for the queued agents
            for (int e = lane; e < nqa; e += 32) {
                double p = __dadd_rn(0.5, __ddiv_rn(__dmul_rn(P.atanScale, atan_rn(S.qaX[e])), 3.141592653589793));
                if (u2d(S.qaU[e]) < p) S.dec[S.qaJ[e]] |= T_ATANDIES;
            }
            nqa = 0;
            __syncwarp();
        };
        auto flush_move = [&]() {  // WeightedMove::execute, actions/WeightedMove.cpp:56-98, for the queued agents
            for (int e = lane; e < nqm; e += 32) {
                const int j = S.qmJ[e];
                const uint32_t u = agent_draws(a.id[s + j], step, STREAM_ACT1, P.key).x;
                int pick = -1;
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
                    const int dst = E.nbr[(size_t)c * MAXN + pick - 1];
                    if (dst >= 0 && !(E.ice && E.ice[dst])) S.dec[j] |= (uint8_t)(pick << DEC_MOVE_SHIFT);
                }
            }
            nqm = 0;
            __syncwarp();
        };

        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            const bool valid = j < n;
            bool needAtan = false, needMove = false;
            double x = 0;
            uint32_t uDeath = 0;
            if (valid) {
                const int g = s + j;
                const uint8_t f0 = S.dec[j];
                uint8_t f = f0 & (F_MALE | F_FERTILE);
                const bool hasMate = allPaired ? (f == F_FERTILE) : ((f0 & T_HASMATE) != 0);
                const int64_t id = a.id[g];
                const float birth = a.birth[g];
                float age = P.storeAge ? a.age[g] : 0.0f;
                const float lastBirth = I.hasFert ? a.lastBirth[g] : 0.0f;
                uint4 r0 = make_uint4(0, 0, 0, 0);
                if (I.needAct0) r0 = agent_draws(id, step, STREAM_ACT0, P.key);
                bool alive = true, born = false;
#pragma unroll 1
                for (int k = 0; k < P.nOps && alive; k++) {
                    switch (prog_op(P, k)) {
                    case OP_GETOLD:  // actions/GetOld.cpp:37-48
                        age = __fsub_rn(P.t, birth);
                        break;
                    case OP_ATANDEATH: {  // actions/ATanDeath.cpp:66-90
                        age = __fsub_rn(P.t, birth);
                        x = __dmul_rn(P.atanSlope, __dadd_rn((double)age, -P.atanMaxAge));
                        if (x > P.atanXlo) {          // below: probability < 0, nobody dies
                            if (x < P.atanXhi) { needAtan = true; uDeath = r0.x; }  // decided when the queue is flushed
                            else alive = false;       // above: probability > 1
                        }
                        break;
                    }
                    case OP_OLDAGEDEATH: {  // actions/OldAgeDeath.cpp:48-67
                        age = __fsub_rn(P.t, birth);
                        const uint32_t u = agent_draws(id, step, STREAM_ACT1, P.key).w;
                        if ((double)age > __dadd_rn(P.oadMaxAge, u2range(u, P.oadLo, P.oadHi))) alive = false;
                        break;
                    }
                    case OP_WEIGHTEDMOVE:  // actions/WeightedMove.cpp:45-106; the neighbour is chosen at the flush
                        if (u2d(r0.y) < P.moveProb) needMove = true;
                        break;
                    case OP_FERTILITY: {  // actions/Fertility.cpp:49-74
                        bool fert;
                        if (!(f & F_MALE)) fert = (age > P.fertMinAge) && (age < P.fertMaxAge) && (__fsub_rn(P.t, lastBirth) > P.fertInterbirth);
                        else fert = age > P.fertMinAge;
                        f = (uint8_t)((f & F_MALE) | (fert ? F_FERTILE : 0));
                        break;
                    }
                    case OP_VERHULST: {  // actions/Verhulst.cpp:101-115 -> LinearBirth.cpp:122-168, LinearDeath.cpp:131-153
                        if (bC > 0) {
                            if (!(f & F_MALE) && hasMate && u2d(r0.z) < bC) born = true;
                        } else if (bC < 0) {
                            if (u2d(r0.z) < -bC) alive = false;
                        }
                        if (alive && u2d(r0.w) < dC) alive = false;
                        break;
                    }
                    case OP_DROWN:  // populations/tut_EnvironAltPop.cpp:100-116 (EVENT_ID_GEO)
                        if (E.alt[c] < 0 || (E.ice && E.ice[c])) alive = false;
                        break;
                    }
                }
                if (P.storeAge) a.age[g] = age;
                S.dec[j] = (uint8_t)(f | (born ? F_BORN : 0) | (alive ? 0 : T_DEADNOW));
            }
            // queue the rare expensive work
            unsigned ma = __ballot_sync(FULL, needAtan), mm = __ballot_sync(FULL, needMove);
            if (nqa + __popc(ma) > QCAP) flush_atan();
            if (nqm + __popc(mm) > QCAP) flush_move();
            if (needAtan) { const int e = nqa + __popc(ma & lt); S.qaX[e] = x; S.qaU[e] = uDeath; S.qaJ[e] = (uint16_t)j; }
            if (needMove) { const int e = nqm + __popc(mm & lt); S.qmJ[e] = (uint16_t)j; }
            nqa += __popc(ma);
            nqm += __popc(mm);
            __syncwarp();
        }
        flush_atan();
        flush_move();

        // ---- commit: final decision bytes, per-cell counts ---------------------------------------------------------
        int stayC = 0, bornC = 0;
        int outC[MAXN] = {0, 0, 0, 0, 0, 0};
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            const bool valid = j < n;
            const uint8_t v = valid ? S.dec[j] : (uint8_t)T_DEADNOW;
            const bool atanDies = (v & T_ATANDIES) != 0;
            const bool dead = atanDies || (v & T_DEADNOW);
            int code = (v >> DEC_MOVE_SHIFT) & 7;
            const bool born = (v & F_BORN) && !(atanDies && I.bornAfterAtan);
            const bool moveRegistered = valid && code != 0 && !(atanDies && I.moveAfterAtan);
            if (valid) dec[s + j] = (uint8_t)((v & (F_MALE | F_FERTILE)) | (born ? F_BORN : 0) | ((dead ? DEC_DEAD : code) << DEC_MOVE_SHIFT));
            if (dead) code = 0;
            stayC += __popc(__ballot_sync(FULL, valid && !dead && code == 0));
            bornC += __popc(__ballot_sync(FULL, valid && born));
            nDead += __popc(__ballot_sync(FULL, valid && dead));
            nMove += __popc(__ballot_sync(FULL, moveRegistered));
            const unsigned mv = __ballot_sync(FULL, valid && !dead && code != 0);
            if (mv) {
#pragma unroll
                for (int q = 0; q < MAXN; q++) outC[q] += __popc(__ballot_sync(FULL, valid && !dead && code == q + 1));
            }
        }
        nBorn += bornC;
        if (bornC > MAXMOTHERS && lane == 0) atomicExch(&st->oversize, 1);
        if (lane == 0) { stay[c] = stayC; birthCount[c] = bornC; }  // a cell belongs to exactly one warp: plain stores
        if (lane < MAXN) {
            int cnt = 0;
#pragma unroll
            for (int q = 0; q < MAXN; q++) if (lane == q) cnt = outC[q];
            if (cnt) atomicAdd(&arrive[E.nbr[(size_t)c * MAXN + lane]], cnt);
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (nDead) atomicAdd(&st->nDeaths, nDead);
        if (nMove) atomicAdd(&st->nMoves, nMove);
        if (nBorn) atomicAdd(&st->nBirths, nBorn);
    }
}

// ---------------------------------------------------------------------------------------------
// pass 2: counting-sort scatter (performMoves core/SPopulation.cpp:1058-1092) + newborns
// (makeOffspring / createAgentAtIndex :823-847,880-918; makePopSpecificOffspring populations/tut_EnvironAltPop.cpp:141-149)
struct WarpSmemB {
    int64_t motherId[MAXMOTHERS];
};

__global__ void __launch_bounds__(CW * 32)
k_cell_scatter(const DevStats *__restrict__ st, AgentArrays a, AgentArrays o, int nCells, const int *__restrict__ cellStart,
               const uint8_t *__restrict__ dec, const int *__restrict__ nbr, const int *__restrict__ newStart,
               const int *__restrict__ stay, const int *__restrict__ arrive, int *__restrict__ cursor,
               const int *__restrict__ birthBase, float t, int storeAge, RngKey key) {
    __shared__ WarpSmemB smem[CW];
    if (st->overflow || st->oversize) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpSmemB &S = smem[wid];
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = lanemask_lt();
    const unsigned step = st->step;
    const long long nextID = st->nextID;
    const int gw = blockIdx.x * CW + wid, nW = gridDim.x * CW;
    for (int c = gw; c < nCells; c += nW) {
        const int s = cellStart[c], n = cellStart[c + 1] - s;
        if (n == 0) continue;
        const int ns = newStart[c];
        int stayBase = 0, nMothers = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            const bool valid = j < n;
            const int g = s + j;
            const uint8_t v = valid ? dec[g] : (uint8_t)(DEC_DEAD << DEC_MOVE_SHIFT);
            const int code = v >> DEC_MOVE_SHIFT;
            const bool alive = code != DEC_DEAD, born = (v & F_BORN) != 0;
            const unsigned ms = __ballot_sync(FULL, alive && code == 0);
            const unsigned mb = __ballot_sync(FULL, born);
            int64_t id = 0;
            if (alive || born) id = a.id[g];
            if (alive) {
                int d = c, pos;
                if (code == 0) {
                    pos = ns + stayBase + __popc(ms & lt);
                } else {
                    d = nbr[(size_t)c * MAXN + code - 1];
                    pos = newStart[d] + stay[d] + atomicAdd(&cursor[d], 1);
                }
                o.id[pos] = id;
                o.birth[pos] = a.birth[g];
                o.lastBirth[pos] = a.lastBirth[g];
                o.cell[pos] = d;
                o.flags[pos] = (uint8_t)(v & (F_MALE | F_FERTILE));
                if (storeAge) o.age[pos] = a.age[g];
            }
            if (born) S.motherId[nMothers + __popc(mb & lt)] = id;
            stayBase += __popc(ms);
            nMothers += __popc(mb);
        }
        __syncwarp();
        // newborn id = nextID + rank of (cell, mother id) among this step's births; the same rank places the baby
        const int babyBase = ns + stayBase + arrive[c];
        for (int m = lane; m < nMothers; m += 32) {
            const int64_t mid = S.motherId[m];
            int r = 0;
            for (int e = 0; e < nMothers; e++) r += (S.motherId[e] < mid) ? 1 : 0;
            const int64_t cid = nextID + birthBase[c] + r;
            const uint32_t gnd = agent_draws(cid, step, STREAM_BABY, key).x >> 31;  // (uchar)(2*wrandd())
            const int pos = babyBase + r;
            o.id[pos] = cid;
            o.birth[pos] = t;
            o.lastBirth[pos] = 0.0f;
            o.cell[pos] = c;
            o.flags[pos] = (uint8_t)(gnd ? F_MALE : F_FERTILE);  // females are born FERTILE, core/SPopulation.cpp:895-898
            if (storeAge) o.age[pos] = 0.0f;
        }
        __syncwarp();
    }
}

}  // namespace qhg
