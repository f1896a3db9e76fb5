// qhg_tiles.cuh -- the tiled fast path: two passes over the agent state per step.
//
// Agents are binned by cell, so a contiguous range of agents is a contiguous range of whole cells.  A tile is
// about TILE_T agents cut at cell boundaries; one CTA owns one tile and keeps the per-cell work in shared memory:
//
//   pass 1  k_tile_decide   read id, birth, lastBirth, flags, cell (21 B/agent) -> pairing inside each cell
//                           (random-key bucket ranking in shared memory) -> all actions -> ONE decision byte per
//                           agent + per-cell stay/arrive/birth counts                                   (1 B/agent)
//   scan    k_scan_*        new cell starts from the counts
//   pass 2  k_tile_scatter  read decision byte + state (22 B/agent), write the survivors, movers and newborns
//                           into the other buffer at their new position                                  (21 B/agent)
//
// Nothing depends on the order of agents inside a cell: pairing ranks by (random key, id), newborn ids by
// (cell, mother id), random draws are keyed by agent id.  So atomics may hand out positions in any order and the
// result is still bit-identical to the oracle as a set of agents.
#pragma once
#include "qhg_kernels.cuh"

namespace qhg {

constexpr int TILE_T = 1024;    // nominal agents per tile
constexpr int TILE_CAP = 2048;  // shared-memory capacity of a tile (a tile ends at a cell boundary)
constexpr int TB = 256;         // threads per CTA
constexpr int ITEMS = TILE_CAP / TB;

// decision byte: bit0 male, bit1 fertile (the agent's new flags), bit2 gave birth, bits 3-5 move code
constexpr int DEC_MOVE_SHIFT = 3;
constexpr uint8_t DEC_DEAD = 7;  // move code 7 = dead, 0 = stays, 1..6 = neighbour slot + 1

struct TileSmem {
    int cnt[TILE_CAP];        // head flags -> bucket counts -> stay flags
    int off[TILE_CAP + 1];    // exclusive scans of cnt
    uint32_t key[TILE_CAP];   // pairing keys
    int nfm[TILE_CAP];        // per segment: fertile females (low 16 bits) and males (high 16 bits); later stay counts
    uint16_t seg[TILE_CAP];       // segment (= occupied cell) index of each agent of the tile
    uint16_t segStart[TILE_CAP + 2];
    uint16_t sorted[TILE_CAP];    // agents grouped by bucket
    uint16_t ranked[TILE_CAP];    // agents by (segment, sex, rank)
    uint8_t dec[TILE_CAP];
    int tmp[16];
    int a0, a1;
};

// first cell boundary at or after tile*TILE_T
__device__ __forceinline__ int tile_start(int tile, int n, const int *__restrict__ cell, const int *__restrict__ cellStart) {
    long long p = (long long)tile * TILE_T;
    if (p >= n) return n;
    if (p == 0) return 0;
    int c = cell[p];
    int s = cellStart[c];
    return (s == (int)p) ? (int)p : cellStart[c + 1];
}

// exclusive scan of in[0..n) into out[0..n], blocked over the CTA; returns the total.  All threads must call.
__device__ __forceinline__ int block_exscan(const int *in, int *out, int n, int *tmp) {
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int base = t * ITEMS;
    int v[ITEMS];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = base + k;
        int x = (idx < n) ? in[idx] : 0;
        v[k] = sum;
        sum += x;
    }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    __syncthreads();  // tmp may still be read from a previous call
    if (lane == 31) tmp[wid] = inc;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < TB / 32; w++) {
        int x = tmp[w];
        if (w < wid) wbase += x;
        total += x;
    }
    const int ex = wbase + inc - sum;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = base + k;
        if (idx < n) out[idx] = ex + v[k];
    }
    if (t == 0) out[n] = total;
    __syncthreads();
    return total;
}

// shared-memory counter add with one atomic per distinct key in the warp; returns nothing (counts only)
__device__ __forceinline__ void smem_agg_add(int *counter, int key, int amount, bool active) {
    unsigned act = __ballot_sync(0xffffffffu, active);
    if (active) {
        unsigned peers = __match_any_sync(act, key);
        if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&counter[key], amount * __popc(peers));
    }
}

// segment structure of a tile: seg[j] = index of agent j's cell among the occupied cells of the tile,
// segStart[s] = first agent of segment s.  cellv[k] holds the cell of item k of this thread.  Returns #segments.
__device__ __forceinline__ int build_segments(TileSmem &S, int nt, int a0, const int *__restrict__ cell, int (&cellv)[ITEMS]) {
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int j = threadIdx.x + k * TB;
        if (j < nt) {
            int c = cell[a0 + j];
            cellv[k] = c;
            S.cnt[j] = (j == 0 || cell[a0 + j - 1] != c) ? 1 : 0;
        }
    }
    __syncthreads();
    int nseg = block_exscan(S.cnt, S.off, nt, S.tmp);
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int j = threadIdx.x + k * TB;
        if (j < nt) {
            int head = S.cnt[j];
            int sg = S.off[j] + head - 1;
            S.seg[j] = (uint16_t)sg;
            if (head) S.segStart[sg] = (uint16_t)j;
        }
    }
    if (threadIdx.x == 0) S.segStart[nseg] = (uint16_t)nt;
    __syncthreads();
    return nseg;
}

// ---------------------------------------------------------------------------------------------
// pass 1
__global__ void __launch_bounds__(TB)
k_tile_decide(DevStats *__restrict__ st, AgentArrays a, ActParams P, CellEnv E, const int *__restrict__ cellStart,
              int doPair, int needMate, int *__restrict__ stay, int *__restrict__ arrive, int *__restrict__ birthCount,
              uint8_t *__restrict__ dec, int *__restrict__ mateOut) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem &S = *reinterpret_cast<TileSmem *>(smem_raw);
    const int n = st->nAgents;
    const unsigned step = st->step;
    const int tid = threadIdx.x;
    if (tid == 0) {
        S.a0 = tile_start(blockIdx.x, n, a.cell, cellStart);
        S.a1 = tile_start(blockIdx.x + 1, n, a.cell, cellStart);
    }
    __syncthreads();
    const int a0 = S.a0, nt = S.a1 - S.a0;
    if (nt <= 0) return;
    if (nt > TILE_CAP) {
        if (tid == 0) atomicExch(&st->oversize, 1);
        return;
    }
    int cellv[ITEMS];
    const int nseg = build_segments(S, nt, a0, a.cell, cellv);

    int64_t id[ITEMS];
    uint8_t f[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int j = tid + k * TB;
        if (j < nt) { id[k] = a.id[a0 + j]; f[k] = a.flags[a0 + j]; }
        else { id[k] = 0; f[k] = 0; }
    }

    // ---- pairing: RandomPair::findMates (actions/RandomPair.cpp:146-279) under the counter-mode law --------------
    // Inside a cell the fertile females and the fertile males are ranked by (random key, id); equal ranks mate.
    // Ranking is a bucket sort: bucket = floor(key * n / 2^32) (about one agent per bucket), exclusive scan over the
    // bucket counts, then an ordering inside the (tiny) bucket.
    bool hasMate[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) hasMate[k] = false;
    if (doPair) {
        for (int j = tid; j < nt; j += TB) { S.cnt[j] = 0; if (j < nseg) S.nfm[j] = 0; }
        __syncthreads();
        uint32_t key[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            int j = tid + k * TB;
            bool fert = (j < nt) && (f[k] & F_FERTILE);
            key[k] = 0;
            int sg = 0, male = 0;
            if (fert) {
                key[k] = agent_draws(id[k], step, STREAM_PAIR, P.key).x;
                S.key[j] = key[k];
                sg = S.seg[j];
                male = f[k] & F_MALE;
            }
            // one counter word per segment: females in the low half, males in the high half
            unsigned act = __ballot_sync(0xffffffffu, fert);
            if (fert) {
                unsigned peers = __match_any_sync(act, sg * 2 + male);
                if ((tid & 31) == __ffs(peers) - 1) atomicAdd(&S.nfm[sg], __popc(peers) << (male ? 16 : 0));
            }
        }
        __syncthreads();
        int slot[ITEMS], ord[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            int j = tid + k * TB;
            slot[k] = -1; ord[k] = 0;
            if (j < nt && (f[k] & F_FERTILE)) {
                int sg = S.seg[j];
                int w = S.nfm[sg];
                int nF = w & 0xffff, nM = w >> 16;
                if (nF > 0 && nM > 0) {  // a cell with one sex only pairs nobody (:186)
                    int male = f[k] & F_MALE;
                    int ns = male ? nM : nF;
                    int R = S.segStart[sg] + (male ? nF : 0);
                    slot[k] = R + (int)__umulhi(key[k], (uint32_t)ns);
                    ord[k] = atomicAdd(&S.cnt[slot[k]], 1);
                }
            }
        }
        __syncthreads();
        block_exscan(S.cnt, S.off, nt, S.tmp);
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            int j = tid + k * TB;
            if (slot[k] >= 0) S.sorted[S.off[slot[k]] + ord[k]] = (uint16_t)j;
        }
        __syncthreads();
        int rk[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            int j = tid + k * TB;
            rk[k] = -1;
            if (slot[k] >= 0) {
                int sg = S.seg[j];
                int w = S.nfm[sg];
                int nF = w & 0xffff;
                int male = f[k] & F_MALE;
                int R = S.segStart[sg] + (male ? nF : 0);
                int lo = S.off[slot[k]], hi = S.off[slot[k] + 1];
                int rib = 0;
                for (int e = lo; e < hi; e++) {
                    int je = S.sorted[e];
                    if (je == j) continue;
                    uint32_t ke = S.key[je];
                    if (ke < key[k] || (ke == key[k] && a.id[a0 + je] < id[k])) rib++;
                }
                int pbase = S.off[R];
                rk[k] = lo - pbase + rib;
                S.ranked[pbase + rk[k]] = (uint16_t)j;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            int j = tid + k * TB;
            int m = -3;
            if (slot[k] >= 0) {
                int sg = S.seg[j];
                int w = S.nfm[sg];
                int nF = w & 0xffff, nM = w >> 16;
                if (rk[k] < min(nF, nM)) {
                    int male = f[k] & F_MALE;
                    int Rother = S.segStart[sg] + (male ? 0 : nF);
                    m = a0 + S.ranked[S.off[Rother] + rk[k]];
                    hasMate[k] = true;
                }
            }
            if (mateOut && j < nt) mateOut[a0 + j] = m;
        }
        __syncthreads();
    }
    (void)needMate;

    // ---- actions -------------------------------------------------------------------------------------------------
    for (int j = tid; j < nseg; j += TB) S.nfm[j] = 0;  // now: stayers per segment
    __syncthreads();
    int nDead = 0, nMove = 0, nBorn = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int j = tid + k * TB;
        const bool valid = j < nt;
        bool stays = false, moves = false, born = false;
        int sg = 0, c = 0, to = 0;
        if (valid) {
            const int g = a0 + j;
            c = cellv[k];
            sg = S.seg[j];
            Decision d = run_actions(P, E, step, id[k], a.birth[g], P.storeAge ? a.age[g] : 0.0f, c, f[k], hasMate[k], a.lastBirth + g);
            if (P.storeAge && d.alive) a.age[g] = d.age;
            uint8_t code = d.alive ? (uint8_t)d.pick : DEC_DEAD;
            dec[g] = (uint8_t)(d.f | (d.born ? F_BORN : 0) | (code << DEC_MOVE_SHIFT));
            stays = d.alive && d.pick == 0;
            moves = d.alive && d.pick != 0;
            born = d.born;
            to = d.to;
            if (d.moving) nMove++;
            if (!d.alive) nDead++;
            if (born) nBorn++;
        }
        smem_agg_add(S.nfm, sg, 1, stays);
        warp_agg_inc(arrive, to, moves);
        warp_agg_inc(birthCount, c, born);
    }
    __syncthreads();
    // a cell lies in exactly one tile: plain stores of the stay counts
    for (int sg = tid; sg < nseg; sg += TB) stay[a.cell[a0 + S.segStart[sg]]] = S.nfm[sg];
    nDead = warp_sum(nDead); nMove = warp_sum(nMove); nBorn = warp_sum(nBorn);
    if ((tid & 31) == 0) {
        if (nDead) atomicAdd(&st->nDeaths, nDead);
        if (nMove) atomicAdd(&st->nMoves, nMove);
        if (nBorn) atomicAdd(&st->nBirths, nBorn);
    }
}

// ---------------------------------------------------------------------------------------------
// pass 2: counting-sort scatter (performMoves core/SPopulation.cpp:1058-1092) + newborns
// (makeOffspring / createAgentAtIndex :823-847,880-918; makePopSpecificOffspring populations/tut_EnvironAltPop.cpp:141-149)
__global__ void __launch_bounds__(TB)
k_tile_scatter(const DevStats *__restrict__ st, AgentArrays a, AgentArrays o, const int *__restrict__ cellStart,
               const uint8_t *__restrict__ dec, const int *__restrict__ nbr, const int *__restrict__ newStart,
               const int *__restrict__ stay, const int *__restrict__ arrive, int *__restrict__ cursor,
               const int *__restrict__ birthBase, float t, int storeAge, RngKey key) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem &S = *reinterpret_cast<TileSmem *>(smem_raw);
    if (st->overflow || st->oversize) return;
    const int n = st->nAgents;
    const unsigned step = st->step;
    const long long nextID = st->nextID;
    const int tid = threadIdx.x;
    if (tid == 0) {
        S.a0 = tile_start(blockIdx.x, n, a.cell, cellStart);
        S.a1 = tile_start(blockIdx.x + 1, n, a.cell, cellStart);
    }
    __syncthreads();
    const int a0 = S.a0, nt = S.a1 - S.a0;
    if (nt <= 0 || nt > TILE_CAP) return;
    int cellv[ITEMS];
    build_segments(S, nt, a0, a.cell, cellv);
    uint8_t dv[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int j = tid + k * TB;
        dv[k] = 0;
        if (j < nt) {
            dv[k] = dec[a0 + j];
            S.dec[j] = dv[k];
            S.cnt[j] = ((dv[k] >> DEC_MOVE_SHIFT) == 0) ? 1 : 0;
        }
    }
    __syncthreads();
    block_exscan(S.cnt, S.off, nt, S.tmp);  // stayers before agent j in the tile
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int j = tid + k * TB;
        if (j >= nt) continue;
        const int g = a0 + j;
        const int c = cellv[k];
        const int code = dv[k] >> DEC_MOVE_SHIFT;
        const int sg = S.seg[j];
        int64_t id = 0;
        if (code != DEC_DEAD || (dv[k] & F_BORN)) id = a.id[g];
        if (code != DEC_DEAD) {
            int d = c, pos;
            if (code == 0) {
                pos = newStart[c] + S.off[j] - S.off[S.segStart[sg]];
            } else {
                d = nbr[(size_t)c * MAXN + code - 1];
                pos = newStart[d] + stay[d] + atomicAdd(&cursor[d], 1);
            }
            o.id[pos] = id;
            o.birth[pos] = a.birth[g];
            o.lastBirth[pos] = a.lastBirth[g];
            o.cell[pos] = d;
            o.flags[pos] = (uint8_t)(dv[k] & (F_MALE | F_FERTILE));
            if (storeAge) o.age[pos] = a.age[g];
        }
        if (dv[k] & F_BORN) {
            // newborn id = nextID + rank of (cell, mother id) among this step's births
            int r = 0;
            const int s0 = S.segStart[sg], s1 = S.segStart[sg + 1];
            for (int e = s0; e < s1; e++) {
                if (e != j && (S.dec[e] & F_BORN) && a.id[a0 + e] < id) r++;
            }
            const int64_t cid = nextID + birthBase[c] + r;
            const uint32_t gnd = agent_draws(cid, step, STREAM_BABY, key).x >> 31;  // (uchar)(2*wrandd())
            const int pos = newStart[c] + stay[c] + arrive[c] + r;
            o.id[pos] = cid;
            o.birth[pos] = t;
            o.lastBirth[pos] = 0.0f;
            o.cell[pos] = c;
            o.flags[pos] = (uint8_t)(gnd ? F_MALE : F_FERTILE);  // females are born FERTILE, core/SPopulation.cpp:895-898
            if (storeAge) o.age[pos] = 0.0f;
        }
    }
}

}  // namespace qhg
