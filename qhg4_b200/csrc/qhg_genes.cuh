// qhg_genes.cuh -- genomes on the device: Genetics<.., BitGeneUtils> and Genetics<.., GeneUtils> (actions/Genetics.cpp,
// genes/BitGeneUtils.cpp, genes/GeneUtils.cpp).
//
// A genome is 2 strands x nBlocks 64-bit words of 1-bit or 2-bit nucleotides (GeneParams::bitsPerNuc; the 2-bit variant
// keeps nucleotides whole: breaks on even bits, doubled mask bits in free recombination, XOR with 01/10/11 as mutation).  Genomes live in a pool of fixed-size rows and are
// NOT moved when the agents are re-binned: an agent carries a 4-byte handle (row index).  Births take rows from a free
// stack (rows of last step's dead) or from the never-used tail; deaths push their rows after all births of the step
// have read their parents.
//
// Offspring law (counter mode, mirrored by oracle/qhg_oracle.cpp::makeGenome): every draw is a word of
// Philox(child id, step, stream) -- stream 4: strand choices (x, y) and mutation count (z); 0x01000000|parent<<20|block/2:
// free-recombination masks; 0x02000000|parent<<20|i/4: crossover break i; 0x03000000|i/4: position of mutation i.
#pragma once
#include "qhg_kernels.cuh"

namespace qhg {

constexpr int MAX_BINO = 64;    // entries of the mutation-count table (utils/BinomialDist.cpp:61-87)
constexpr int MAX_CROSS = 32;   // crossover breaks per parent handled by one warp

struct GeneParams {
    int genomeSize, nBlocks, numCrossOvers, nBino;
    double mutationRate;
    double bino[MAX_BINO];
    int bitsPerNuc;  // 1: BitGeneUtils, 2: GeneUtils
};

// genes/BitGeneUtils.cpp:86-106 with the break list in draw order (never sorted in the reference)
__device__ __forceinline__ unsigned long long make_multi_mask(const unsigned *br, int nbr) {
    unsigned k = 1u - (unsigned)(nbr & 1);
    unsigned long long out = 0;
    int i = nbr - 1;
    for (unsigned j = 0; j < 64; j++) {
        if (i >= 0 && j == 64u - br[i]) { i--; k = 1u - k; }
        out = (out << 1) + k;
    }
    return out;
}

// one warp per birth: Genetics::makeOffspring (actions/Genetics.cpp:285-337)
__global__ void __launch_bounds__(128)
k_make_offspring(const DevStats *__restrict__ st, GenomeCtl *__restrict__ ctl, const BirthEntry *__restrict__ births, GeneParams G,
                 RngKey key, const int *__restrict__ oldSlot, int *__restrict__ newSlot, unsigned long long *__restrict__ pool,
                 const int *__restrict__ freeStack) {
    __shared__ unsigned sbr[4][2][MAX_CROSS];
    if (st->overflow) return;
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    const int nb = G.nBlocks, row = 2 * nb;
    const unsigned step = st->step;
    const unsigned FULL = 0xffffffffu;
    // software pipeline: the next birth's record and its parents' row handles are fetched while this one is worked on
    // rows for the babies: birth e takes the e-th entry from the top of the free stack, or the (e - nFree)-th never-used row --
    // no counter is touched here (k_genome_ctl_reset books the rows after the kernel), so the births do not queue up on one
    // atomic; which row a genome lives in is not part of the result
    const int nBirths = ctl->nBirths, nFree0 = ctl->nFree, hwm0 = ctl->hwm;
    BirthEntry beN{};
    int smN = 0, sfN = 0;
    if (w < nBirths) { beN = births[w]; smN = oldSlot[beN.mother]; sfN = oldSlot[beN.father]; }
    for (int e = w; e < nBirths; e += nW) {
        const BirthEntry be = beN;
        const int sm = smN, sf = sfN;
        if (e + nW < nBirths) { beN = births[e + nW]; smN = oldSlot[beN.mother]; sfN = oldSlot[beN.father]; }
        int slot = 0;
        if (lane == 0) {  // a row for the baby
            slot = (e < nFree0) ? freeStack[nFree0 - 1 - e] : hwm0 + (e - nFree0);
            newSlot[be.babyPos] = slot;
        }
        slot = __shfl_sync(FULL, slot, 0);
        const unsigned long long *gm = pool + (size_t)sm * row;
        const unsigned long long *gf = pool + (size_t)sf * row;
        unsigned long long *gb = pool + (size_t)slot * row;
        const uint4 g0 = agent_draws(be.cid, step, 4u, key);
        const int i1 = (int)(g0.x >> 31), i2 = (int)(g0.y >> 31);  // (int)(2 * wrandd())
        // crossover breaks of both parents: lane i holds break i
        unsigned brM = 0, brF = 0;
        const int nc = G.numCrossOvers;
        if (nc > 0 && lane < nc) {
            const unsigned nBits = (unsigned)nb * 64u;
            const uint4 dm = agent_draws(be.cid, step, 0x02000000u | (0u << 20) | (unsigned)(lane / 4), key);
            const uint4 df = agent_draws(be.cid, step, 0x02000000u | (1u << 20) | (unsigned)(lane / 4), key);
            const unsigned wm = (lane & 3) == 0 ? dm.x : (lane & 3) == 1 ? dm.y : (lane & 3) == 2 ? dm.z : dm.w;
            const unsigned wf = (lane & 3) == 0 ? df.x : (lane & 3) == 1 ? df.y : (lane & 3) == 2 ? df.z : df.w;
            brM = u2int_s(wm, 0, nBits, (unsigned)G.bitsPerNuc);  // wrandi(0, nBits, BITSINNUC), genes/GeneUtils.cpp:191
            brF = u2int_s(wf, 0, nBits, (unsigned)G.bitsPerNuc);
            sbr[wl][0][lane] = brM;
            sbr[wl][1][lane] = brF;
        }
        __syncwarp();
        // mutations: count from the binomial table, lane i holds position i
        int nMut = 0;
        if (G.mutationRate > 0) {
            const double r = u2d(g0.z);
            while (nMut < G.nBino && r > G.bino[nMut]) nMut++;
        }
        // the words of the parents' rows are all requested before the first child word is stored (the rows live in ONE pool: a
        // store between the loads would order them one memory round trip after the other).  A lane takes TWO consecutive blocks
        // of one parent: one Philox call yields the 128 mask bits of both (blocks 2c and 2c+1 share counter c)
        constexpr int GU = 2;
        const int npair = (nb + 1) >> 1;  // pairs of blocks per parent (the last one may be half empty)
        for (int q0 = 0; q0 < 2 * npair; q0 += 32 * GU) {
            unsigned long long w0[GU][2], w1[GU][2];
#pragma unroll
            for (int u = 0; u < GU; u++) {
                const int q = q0 + u * 32 + lane;
                w0[u][0] = w0[u][1] = w1[u][0] = w1[u][1] = 0;
                if (q < 2 * npair) {
                    const int parent = (q < npair) ? 0 : 1;
                    const int b = 2 * (q - parent * npair);
                    const unsigned long long *P = parent ? gf : gm;
                    w0[u][0] = P[b]; w1[u][0] = P[nb + b];
                    if (b + 1 < nb) { w0[u][1] = P[b + 1]; w1[u][1] = P[nb + b + 1]; }
                }
            }
#pragma unroll
            for (int u = 0; u < GU; u++) {
                const int q = q0 + u * 32 + lane;
                if (q >= 2 * npair) continue;
                const int parent = (q < npair) ? 0 : 1;
                const int bb = 2 * (q - parent * npair);
                const int pick = parent ? i2 : i1;
                uint4 d = make_uint4(0, 0, 0, 0);
                if (nc == -1) d = agent_draws(be.cid, step, 0x01000000u | ((unsigned)parent << 20) | (unsigned)(bb / 2), key);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int b = bb + h;
                    if (b >= nb) continue;
                    const unsigned long long p0 = w0[u][h], p1 = w1[u][h];
                    unsigned long long t0 = p0, t1 = p1;
                    if (nc == -1) {  // free recombination, genes/BitGeneUtils.cpp:190-220
                        unsigned long long L = h ? (((unsigned long long)d.z << 32) + d.w) : (((unsigned long long)d.x << 32) + d.y);
                        if (G.bitsPerNuc == 2) { L &= 0x5555555555555555ull; L += L << 1; }  // makeFreeMask, genes/GeneUtils.cpp:322-340
                        t0 = (L & p0) | (~L & p1);
                        t1 = (L & p1) | (~L & p0);
                    } else if (nc > 0) {  // crossover, genes/BitGeneUtils.cpp:116-186
                        unsigned mine[MAX_CROSS];
                        int cnt = 0, below = 0;
                        for (int i = 0; i < nc; i++) {
                            const unsigned pos = sbr[wl][parent][i];
                            const int pb = (int)(pos >> 6);
                            if (pb < b) below++;
                            else if (pb == b) mine[cnt++] = pos & 63u;
                        }
                        const int cur = below & 1;
                        const unsigned long long c0 = cur ? p1 : p0, c1 = cur ? p0 : p1;
                        if (cnt == 0) { t0 = c0; t1 = c1; }
                        else {
                            const unsigned long long L = make_multi_mask(mine, cnt);
                            t0 = (L & c0) | (~L & c1);
                            t1 = (L & c1) | (~L & c0);
                        }
                    }
                    gb[parent * nb + b] = pick ? t1 : t0;
                }
            }
        }
        __syncwarp();
        // mutateNucs (genes/BitGeneUtils.cpp:57-75): flips commute, one atomic XOR per mutation
        for (int m = lane; m < nMut; m += 32) {
            if (G.bitsPerNuc == 2) {  // genes/GeneUtils.cpp:112-146: two draws per mutation (position on an even bit, mask 1..3)
                const uint4 d = agent_draws(be.cid, step, 0x03000000u | (unsigned)(m / 2), key);
                const unsigned wp = (m & 1) ? d.z : d.x, wk = (m & 1) ? d.w : d.y;
                const unsigned pos = u2int_s(wp, 0, 4u * (unsigned)G.genomeSize, 2u);
                atomicXor(&gb[pos >> 6], (unsigned long long)u2int(wk, 1, 4) << (pos & 63u));
            } else {
                const uint4 d = agent_draws(be.cid, step, 0x03000000u | (unsigned)(m / 4), key);
                const unsigned wv = (m & 3) == 0 ? d.x : (m & 3) == 1 ? d.y : (m & 3) == 2 ? d.z : d.w;
                const unsigned pos = u2int(wv, 0, 2u * (unsigned)G.genomeSize);
                atomicXor(&gb[pos >> 6], 1ull << (pos & 63u));
            }
        }
        __syncwarp();
    }
}

// rows of the agents that died this step go back to the free stack (after the births have read their parents)
__global__ void k_free_genomes(const DevStats *__restrict__ st, GenomeCtl *__restrict__ ctl, const int *__restrict__ dest,
                               const int *__restrict__ oldSlot, int *__restrict__ freeStack) {
    if (st->overflow) return;
    const int n = st->nAgents;
    const unsigned lt = lanemask_lt();
    for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {  // one push per warp, not per dead agent
        const int i = i0 + (int)threadIdx.x;
        const bool dead = i < n && dest[i] < 0;
        const unsigned m = __ballot_sync(0xffffffffu, dead);
        if (m) {
            const int leader = __ffs(m) - 1;
            int base = 0;
            if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(&ctl->nFree, __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (dead) freeStack[base + __popc(m & lt)] = oldSlot[i];
        }
    }
}

// the same on the fast path: an agent died this step (or left this rank) if its decision byte says so (move code 7).  Four
// decision bytes per thread, one push on the free stack per BLOCK and round (1024 agents), not per warp: every push is an
// atomic on the same counter.
__global__ void __launch_bounds__(256)
k_free_genomes_dec(const DevStats *__restrict__ st, GenomeCtl *__restrict__ ctl, const uint8_t *__restrict__ dec,
                   const int *__restrict__ oldSlot, int *__restrict__ freeStack) {
    __shared__ int wsum[8];
    __shared__ int blockBase;
    if (st->overflow || st->oversize || st->halt) return;
    const int n = st->nAgents;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    const uint32_t *dw = reinterpret_cast<const uint32_t *>(dec);  // the array has AGENT_SLACK bytes past the last agent
    for (int i0 = blockIdx.x * 1024; i0 < n; i0 += gridDim.x * 1024) {
        const int i = i0 + 4 * (int)threadIdx.x;
        uint32_t dead = 0;  // bit b: agent i + b is gone
        if (i < n) {
            const uint32_t w = dw[i >> 2];
#pragma unroll
            for (int b = 0; b < 4; b++) if (i + b < n && ((w >> (8 * b + 3)) & 31u) == 7u) dead |= 1u << b;
        }
        const int cnt = __popc(dead);
        // exclusive prefix of cnt over the block
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < 8; w++) { const int v = wsum[w]; wsum[w] = tot; tot += v; }
            blockBase = tot ? atomicAdd(&ctl->nFree, tot) : 0;
        }
        __syncthreads();
        int pos = blockBase + wsum[wid] + incl - cnt;
#pragma unroll
        for (int b = 0; b < 4; b++) if (dead & (1u << b)) freeStack[pos++] = oldSlot[i + b];
        __syncthreads();
    }
    (void)lt;
}

// sharded runs: the agents that arrived from other ranks took the rows after those of the births (k_place_migrants*), `st`
// then carries their number; the pool has `poolRows` rows
__global__ void k_genome_ctl_reset(GenomeCtl *ctl, int bookRows, int resetBirths, DevStats *st = nullptr, int arrivals = 0, int poolRows = 0) {
    if (bookRows && !(st && (st->overflow || st->oversize || st->halt))) {
        const int nb = ctl->nBirths + ((st && arrivals) ? st->nRecv : 0), take = min(nb, ctl->nFree);
        ctl->nFree -= take;
        ctl->hwm += nb - take;
        if (st && poolRows > 0 && ctl->hwm > poolRows) st->overflow = 1;
    }
    if (resetBirths) ctl->nBirths = 0;
}

__global__ void k_gather_genomes(const DevStats *__restrict__ st, int row, const int *__restrict__ slot,
                                 const unsigned long long *__restrict__ pool, unsigned long long *__restrict__ out) {
    const int n = st->nAgents;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)n * row; i += (size_t)gridDim.x * blockDim.x) {
        const size_t a = i / row, k = i % row;
        out[i] = pool[(size_t)slot[a] * row + k];
    }
}

}  // namespace qhg
